// K1+K3, grid form for LINE meshes numbered like Mesh.Line (kiltro/MeshGeneration/Mesh.py:14-30: node i at index i, element
// i = [i, i+1]): the 1-D counterpart of assemble_grid.cu.  Replaces the same call, Assemble (kiltro/FiniteElements/
// Assembly.py:8-92), with the element routine of VariationalPrinciple.py:42-69,117-161 (element.cuh::line2).
//
// One thread per node: row i of the tridiagonal matrix takes its <= 2 contributions per entry from the element matrices
// of elements i-1 and i, both evaluated in registers (each element is evaluated twice, by its two nodes: 2 x ~40 fp64
// instructions against 3 x 8 B written per row), added in ascending element order -- the reference's sequential += -- so the
// result equals the tile kernels' (kb_assemble_fused) up to the FMA contraction of the inlined element routine (<= 2 ulp;
// test_line_grid_assembly_matches_tile_kernels).  Neither the connectivity nor the column indices are
// read; consecutive threads read consecutive nodes and write consecutive CSR entries.
#include "common.cuh"
#include "element.cuh"

using namespace kbel;
ElemOpts kb_make_opts(const kb_material *m, int mode, int petrov, int want_mass);   // element.cu

namespace {

__global__ void __launch_bounds__(256) k_line_verify(const int2 *__restrict__ elements, int64_t nelem, int32_t *__restrict__ flag) {
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    bool bad = false;
    for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < nelem; e += stride) {
        const int2 c = __ldg(elements + e);
        bad |= (c.x != e) || (c.y != e + 1);
    }
    if (bad) *flag = 1;
}

template <bool HAS_U, bool MASS>
__global__ void __launch_bounds__(256) k_assemble_line(const double *__restrict__ points, const double *__restrict__ velocity,
                                                       ElemOpts o, int64_t nnode, const int32_t *__restrict__ indptr,
                                                       double *__restrict__ K, double *__restrict__ M,
                                                       double *__restrict__ Flux) {
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < nnode; i += stride) {
        const bool has_l = i > 0, has_r = i < nnode - 1;
        const double xc = points[i], uc = HAS_U ? velocity[i] : 0.0;
        double kl[4] = {0, 0, 0, 0}, fl[2] = {0, 0}, ml[4] = {0, 0, 0, 0};
        double kr[4] = {0, 0, 0, 0}, fr[2] = {0, 0}, mr[4] = {0, 0, 0, 0};
        if (has_l) {
            const double X[2] = {points[i - 1], xc}, U[2] = {HAS_U ? velocity[i - 1] : 0.0, uc};
            line2(X, U, o, kl, fl, ml);
        }
        if (has_r) {
            const double X[2] = {xc, points[i + 1]}, U[2] = {uc, HAS_U ? velocity[i + 1] : 0.0};
            line2(X, U, o, kr, fr, mr);
        }
        int p = indptr[i];
        if (has_l) { K[p] = kl[2]; if (MASS) M[p] = ml[2]; ++p; }                          // (i, i-1): row 1, col 0 of element i-1
        if (has_l && has_r) { K[p] = kl[3] + kr[0]; if (MASS) M[p] = ml[3] + mr[0]; }
        else if (has_l) { K[p] = kl[3]; if (MASS) M[p] = ml[3]; }
        else { K[p] = kr[0]; if (MASS) M[p] = mr[0]; }
        ++p;
        if (has_r) { K[p] = kr[1]; if (MASS) M[p] = mr[1]; }                               // (i, i+1): row 0, col 1 of element i
        Flux[i] = (has_l && has_r) ? fl[1] + fr[0] : (has_l ? fl[1] : fr[0]);
    }
}

}  // namespace

extern "C" int kb_line_verify(const int32_t *d_elements, int64_t nelem, int64_t nnode, int32_t *d_flag, void *stream) {
    if (!d_elements || !d_flag || nelem < 1 || nnode < 2) return KB_EINVAL;
    cudaStream_t s = (cudaStream_t)stream;
    KB_CUDA(cudaMemsetAsync(d_flag, 0, sizeof(int32_t), s));
    if (nelem != nnode - 1) {
        const int32_t one = 1;
        KB_CUDA(cudaMemcpyAsync(d_flag, &one, sizeof(one), cudaMemcpyHostToDevice, s));
        KB_CUDA(cudaStreamSynchronize(s));
        return 0;
    }
    k_line_verify<<<kb_grid(nelem, 256, 8), 256, 0, s>>>((const int2 *)d_elements, nelem, d_flag);
    KB_LAUNCH_CHECK();
    return 0;
}

extern "C" int kb_assemble_line(const double *d_points, const double *d_velocity, const kb_material *material, int mode,
                                int petrov, int64_t nnode, const int32_t *d_indptr, double *d_K, double *d_M, double *d_Flux,
                                void *stream) {
    if (!d_points || !material || !d_indptr || !d_K || !d_Flux || nnode < 2) return KB_EINVAL;
    if (nnode >= INT32_MAX / 4) return KB_ETOOBIG;
    if (mode != KB_MODE_REFERENCE && mode != KB_MODE_CONSISTENT) return KB_EINVAL;
    cudaStream_t s = (cudaStream_t)stream;
    const ElemOpts o = kb_make_opts(material, mode, petrov, d_M != nullptr);
    const int g = kb_grid(nnode, 256, 8);
    if (d_velocity) {
        if (d_M) k_assemble_line<true, true><<<g, 256, 0, s>>>(d_points, d_velocity, o, nnode, d_indptr, d_K, d_M, d_Flux);
        else k_assemble_line<true, false><<<g, 256, 0, s>>>(d_points, d_velocity, o, nnode, d_indptr, d_K, d_M, d_Flux);
    } else {
        if (d_M) k_assemble_line<false, true><<<g, 256, 0, s>>>(d_points, d_velocity, o, nnode, d_indptr, d_K, d_M, d_Flux);
        else k_assemble_line<false, false><<<g, 256, 0, s>>>(d_points, d_velocity, o, nnode, d_indptr, d_K, d_M, d_Flux);
    }
    KB_LAUNCH_CHECK();
    return 0;
}
