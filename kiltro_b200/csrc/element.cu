// K1 (seam form): batched element matrices written to HBM, one thread per element.
// Replaces the per-element Python call formulation.GetElementalMatrices (VariationalPrinciple.py:94-113 / :194-213)
// for all elements at once.  The production assembly never materialises these (assemble.cu, fused path); this entry
// point exists for the reference's operator seam (GetElementalMatrices / Assemble on explicit element data) and parity
// tests.
#include "element.cuh"

using namespace kbel;

template <bool HAS_U, bool MASS>
__global__ void __launch_bounds__(128) k_elements_quad4(const double2 *__restrict__ points,
                                                        const int4 *__restrict__ elements, int64_t nelem,
                                                        const double2 *__restrict__ velocity, ElemOpts o,
                                                        double *__restrict__ Ke, double *__restrict__ Fe,
                                                        double *__restrict__ Me) {
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < nelem; e += stride) {
        const int4 c = __ldg(elements + e);
        const int n[4] = {c.x, c.y, c.z, c.w};
        double X[4][2], U[4][2];
#pragma unroll
        for (int a = 0; a < 4; ++a) {
            const double2 p = __ldg(points + n[a]);
            X[a][0] = p.x; X[a][1] = p.y;
            if (HAS_U) { const double2 u = __ldg(velocity + n[a]); U[a][0] = u.x; U[a][1] = u.y; }
            else { U[a][0] = 0.0; U[a][1] = 0.0; }
        }
        double ke[16], fe[4], me[16];
        quad4(X, U, o, ke, fe, me);
        double2 *ko = reinterpret_cast<double2 *>(Ke + e * 16);
#pragma unroll
        for (int i = 0; i < 8; ++i) ko[i] = make_double2(ke[2 * i], ke[2 * i + 1]);
        double2 *fo = reinterpret_cast<double2 *>(Fe + e * 4);
        fo[0] = make_double2(fe[0], fe[1]); fo[1] = make_double2(fe[2], fe[3]);
        if (MASS) {
            double2 *mo = reinterpret_cast<double2 *>(Me + e * 16);
#pragma unroll
            for (int i = 0; i < 8; ++i) mo[i] = make_double2(me[2 * i], me[2 * i + 1]);
        }
    }
}

template <bool HAS_U, bool MASS>
__global__ void __launch_bounds__(128) k_elements_line2(const double *__restrict__ points,
                                                        const int2 *__restrict__ elements, int64_t nelem,
                                                        const double *__restrict__ velocity, ElemOpts o,
                                                        double *__restrict__ Ke, double *__restrict__ Fe,
                                                        double *__restrict__ Me) {
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < nelem; e += stride) {
        const int2 c = __ldg(elements + e);
        double X[2] = {__ldg(points + c.x), __ldg(points + c.y)};
        double U[2] = {0.0, 0.0};
        if (HAS_U) { U[0] = __ldg(velocity + c.x); U[1] = __ldg(velocity + c.y); }
        double ke[4], fe[2], me[4];
        line2(X, U, o, ke, fe, me);
        double2 *ko = reinterpret_cast<double2 *>(Ke + e * 4);
        ko[0] = make_double2(ke[0], ke[1]); ko[1] = make_double2(ke[2], ke[3]);
        reinterpret_cast<double2 *>(Fe + e * 2)[0] = make_double2(fe[0], fe[1]);
        if (MASS) {
            double2 *mo = reinterpret_cast<double2 *>(Me + e * 4);
            mo[0] = make_double2(me[0], me[1]); mo[1] = make_double2(me[2], me[3]);
        }
    }
}

ElemOpts kb_make_opts(const kb_material *m, int mode, int petrov, int want_mass) {
    ElemOpts o;
    o.k = m->k; o.rho_cv = m->rho * m->c_v; o.area = m->area; o.q = m->q;
    o.mode = mode; o.petrov = petrov; o.want_mass = want_mass;
    return o;
}

extern "C" int kb_element_matrices(int ndim, const double *d_points, const int32_t *d_elements, int64_t nelem,
                                   const double *d_velocity, const kb_material *material, int mode, int petrov,
                                   double *d_Ke, double *d_Fe, double *d_Me, void *stream) {
    if (!d_points || !d_elements || !material || !d_Ke || !d_Fe || nelem < 0) return KB_EINVAL;
    if (mode != KB_MODE_REFERENCE && mode != KB_MODE_CONSISTENT) return KB_EINVAL;
    if (ndim != 1 && ndim != 2) return KB_EINVAL;
    if (nelem == 0) return 0;
    cudaStream_t s = (cudaStream_t)stream;
    const ElemOpts o = kb_make_opts(material, mode, petrov, d_Me != nullptr);
    const int grid = kb_grid(nelem, 128, 8);
#define KB_EL(K, PT, ET, VT)                                                                                          \
    do {                                                                                                              \
        if (d_velocity && d_Me) K<true, true><<<grid, 128, 0, s>>>((const PT *)d_points, (const ET *)d_elements, nelem, (const VT *)d_velocity, o, d_Ke, d_Fe, d_Me);   \
        else if (d_velocity) K<true, false><<<grid, 128, 0, s>>>((const PT *)d_points, (const ET *)d_elements, nelem, (const VT *)d_velocity, o, d_Ke, d_Fe, d_Me);      \
        else if (d_Me) K<false, true><<<grid, 128, 0, s>>>((const PT *)d_points, (const ET *)d_elements, nelem, (const VT *)d_velocity, o, d_Ke, d_Fe, d_Me);            \
        else K<false, false><<<grid, 128, 0, s>>>((const PT *)d_points, (const ET *)d_elements, nelem, (const VT *)d_velocity, o, d_Ke, d_Fe, d_Me);                     \
    } while (0)
    if (ndim == 2) KB_EL(k_elements_quad4, double2, int4, double2);
    else KB_EL(k_elements_line2, double, int2, double);
#undef KB_EL
    KB_LAUNCH_CHECK();
    return 0;
}
