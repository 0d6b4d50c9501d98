// K10: element operators of the SURVEY 8(f) "next" rows, in seam form (batched element matrices to HBM, which
// kb_assemble_from_elements then gathers into CSR values on the cached pattern -- assemble_gather.cu):
//   KB_OP_ROBIN          replaces AssemblyRobinForces            (FiniteElements/Assembly.py:121-175)
//   KB_OP_CHARGALERKIN   replaces AssembleCharacteristicGalerkin (FiniteElements/Assembly.py:95-117) over
//                        GetElementalCharacteristicGalerkin      (VariationalPrinciple.py:264-308)
// plus kb_axpby, the vector combination the solver uses to fold these operators into K / F (same sparsity pattern).
#include "element.cuh"

using namespace kbel;

namespace {

template <int OP, bool HAS_U>
__global__ void __launch_bounds__(128) k_operator_quad4(const double2 *__restrict__ points, const int4 *__restrict__ elements,
                                                        int64_t nelem, const double2 *__restrict__ velocity, double p0,
                                                        double p1, double *__restrict__ Ke, double *__restrict__ Fe) {
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
        double ke[16], fe[4];
        if (OP == KB_OP_ROBIN) quad4_robin(X, p0, p1, ke, fe);
        else quad4_chargalerkin(X, U, p0, ke, fe);
        double2 *ko = reinterpret_cast<double2 *>(Ke + e * 16);
#pragma unroll
        for (int i = 0; i < 8; ++i) ko[i] = make_double2(ke[2 * i], ke[2 * i + 1]);
        double2 *fo = reinterpret_cast<double2 *>(Fe + e * 4);
        fo[0] = make_double2(fe[0], fe[1]); fo[1] = make_double2(fe[2], fe[3]);
    }
}

template <int OP, bool HAS_U>
__global__ void __launch_bounds__(128) k_operator_line2(const double *__restrict__ points, const int2 *__restrict__ elements,
                                                        int64_t nelem, const double *__restrict__ velocity, double p0,
                                                        double p1, double *__restrict__ Ke, double *__restrict__ Fe) {
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < nelem; e += stride) {
        const int2 c = __ldg(elements + e);
        const double X[2] = {__ldg(points + c.x), __ldg(points + c.y)};
        double U[2] = {0.0, 0.0};
        if (HAS_U) { U[0] = __ldg(velocity + c.x); U[1] = __ldg(velocity + c.y); }
        double ke[4], fe[2];
        if (OP == KB_OP_ROBIN) line2_robin(X, p0, p1, ke, fe);
        else line2_chargalerkin(X, U, p0, ke, fe);
        double2 *ko = reinterpret_cast<double2 *>(Ke + e * 4);
        ko[0] = make_double2(ke[0], ke[1]); ko[1] = make_double2(ke[2], ke[3]);
        reinterpret_cast<double2 *>(Fe + e * 2)[0] = make_double2(fe[0], fe[1]);
    }
}

__global__ void __launch_bounds__(256) k_axpby(int64_t n, double a, const double *x, double b, const double *y, double *out) {
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride)
        out[i] = a * x[i] + (y ? b * y[i] : 0.0);
}

}  // namespace

extern "C" int kb_element_operator(int ndim, int op, const double *d_points, const int32_t *d_elements, int64_t nelem,
                                   const double *d_velocity, double p0, double p1, double *d_Ke, double *d_Fe,
                                   void *stream) {
    if (!d_points || !d_elements || !d_Ke || !d_Fe || nelem < 0) return KB_EINVAL;
    if (ndim != 1 && ndim != 2) return KB_EINVAL;
    if (op != KB_OP_ROBIN && op != KB_OP_CHARGALERKIN) return KB_EINVAL;
    if (nelem == 0) return 0;
    cudaStream_t s = (cudaStream_t)stream;
    const int grid = kb_grid(nelem, 128, 8);
    const bool hu = d_velocity != nullptr;
#define KB_OPK(K, PT, ET, VT, OPV, HU) \
    K<OPV, HU><<<grid, 128, 0, s>>>((const PT *)d_points, (const ET *)d_elements, nelem, (const VT *)d_velocity, p0, p1, d_Ke, d_Fe)
    if (ndim == 2) {
        if (op == KB_OP_ROBIN) KB_OPK(k_operator_quad4, double2, int4, double2, KB_OP_ROBIN, false);
        else if (hu) KB_OPK(k_operator_quad4, double2, int4, double2, KB_OP_CHARGALERKIN, true);
        else KB_OPK(k_operator_quad4, double2, int4, double2, KB_OP_CHARGALERKIN, false);
    } else {
        if (op == KB_OP_ROBIN) KB_OPK(k_operator_line2, double, int2, double, KB_OP_ROBIN, false);
        else if (hu) KB_OPK(k_operator_line2, double, int2, double, KB_OP_CHARGALERKIN, true);
        else KB_OPK(k_operator_line2, double, int2, double, KB_OP_CHARGALERKIN, false);
    }
#undef KB_OPK
    KB_LAUNCH_CHECK();
    return 0;
}

// out = a x + b y   (y may be NULL: out = a x).  out may alias x or y.
extern "C" int kb_axpby(int64_t n, double a, const double *d_x, double b, const double *d_y, double *d_out, void *stream) {
    if (!d_x || !d_out || n < 0) return KB_EINVAL;
    if (n == 0) return 0;
    k_axpby<<<kb_grid(n, 256, 8), 256, 0, (cudaStream_t)stream>>>(n, a, d_x, b, d_y, d_out);
    KB_LAUNCH_CHECK();
    return 0;
}
