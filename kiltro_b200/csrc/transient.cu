// K7 helpers: time-step rule and the matrices of the theta-method step.
// Replaces the (unrunnable) FEMSolver.TransientSolver, kiltro/FEMSolver.py:253-297, and the time-step rule
// of FEMSolver.py:138-145.  The reference's intent is forward Euler with a consistent mass matrix and the reduced block
// of the FULL inverse, T_in^{n+1} = T_in^n - dt (M^-1)[in,in] (K[in,in] T_in^n + F_in); here every step solves
//     (M + theta dt Kff) y = Kff T^n + Fm   on the full system (Kff: rows and columns of fixed DOFs zeroed),
//     T^{n+1} = T^n - dt y at free DOFs, prescribed values at fixed DOFs,
// which for theta = 0 is exactly that scheme (solving with the full M and keeping y[in] IS (M^-1)[in,in] applied to the
// free residual) and for theta > 0 is the incremental theta-method (new capability, no reference counterpart).
// The time loop itself (kb_transient_run) lives in krylov.cu next to the solvers it drives.
#include "common.cuh"

namespace {

// dt = factor * min_e rho c_v |X1 - X0|^2 / (2k): integer atomicMin on the bit pattern (positive doubles order like
// unsigned integers), so the result does not depend on the order of arrival.
template <int NDIM>
__global__ void __launch_bounds__(256) k_dt_rule(const double *__restrict__ points, const int32_t *__restrict__ elements,
                                                 int64_t nelem, double c, unsigned long long *__restrict__ out) {
    constexpr int NPE = NDIM == 2 ? 4 : 2;
    double best = INFINITY;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < nelem; e += stride) {
        const int n0 = elements[e * NPE], n1 = elements[e * NPE + 1];
        double l2 = 0.0;
#pragma unroll
        for (int d = 0; d < NDIM; ++d) { const double dx = points[(int64_t)n1 * NDIM + d] - points[(int64_t)n0 * NDIM + d]; l2 += dx * dx; }
        const double len = sqrt(l2);                     // np.linalg.norm then **2 (FEMSolver.py:142-143)
        const double v = c * (len * len);
        best = fmin(best, v);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) best = fmin(best, __shfl_xor_sync(KB_FULL, best, o));
    if ((threadIdx.x & 31) == 0 && best < INFINITY) atomicMin(out, (unsigned long long)__double_as_longlong(best));
}

__global__ void k_dt_finish(const unsigned long long *in, double factor, double *dt) {
    *dt = factor * __longlong_as_double((long long)*in);
}

__global__ void __launch_bounds__(256) k_transient_matrix(int64_t ndof, const int32_t *__restrict__ indptr,
                                                          const int32_t *__restrict__ indices,
                                                          const double *__restrict__ K, const double *__restrict__ M,
                                                          const int32_t *__restrict__ renum, double c,
                                                          double *__restrict__ Kff, double *__restrict__ A) {
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < ndof; i += stride) {
        const bool rfree = renum[i] >= 0;
        for (int p = indptr[i], pe = indptr[i + 1]; p < pe; ++p) {
            const double kff = (rfree && __ldg(renum + indices[p]) >= 0) ? K[p] : 0.0;
            Kff[p] = kff;
            A[p] = M[p] + c * kff;
        }
    }
}

// out = c_k*Kff + c_m*M (+ identity on the rows of fixed DOFs when asked): the operators of the partitioned solves
__global__ void __launch_bounds__(256) k_masked_matrix(int64_t ndof, const int32_t *__restrict__ indptr,
                                                       const int32_t *__restrict__ indices,
                                                       const double *__restrict__ K, const double *__restrict__ M,
                                                       const int32_t *__restrict__ renum, double c_k, double c_m,
                                                       int identity_fixed, double *__restrict__ out) {
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < ndof; i += stride) {
        const bool rfree = renum[i] >= 0;
        for (int p = indptr[i], pe = indptr[i + 1]; p < pe; ++p) {
            const int col = indices[p];
            const bool ff = rfree && __ldg(renum + col) >= 0;
            double v = ff ? c_k * K[p] : 0.0;
            if (M != nullptr) v += c_m * M[p];                 // the mass matrix stays whole: (M^-1)[in,in] convention
            if (identity_fixed && !rfree && col == (int)i) v = 1.0;
            out[p] = v;
        }
    }
}

__global__ void __launch_bounds__(256) k_mask_free(int64_t ndof, const int32_t *__restrict__ renum,
                                                   const double *__restrict__ in, double *__restrict__ out) {
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < ndof; i += stride)
        out[i] = renum[i] >= 0 ? in[i] : 0.0;
}

}  // namespace

extern "C" int kb_time_step_rule(int ndim, const double *d_points, const int32_t *d_elements, int64_t nelem,
                                 const kb_material *material, double factor, double *d_dt, void *stream) {
    if (!d_points || !d_elements || !material || !d_dt || nelem < 1 || (ndim != 1 && ndim != 2)) return KB_EINVAL;
    cudaStream_t s = (cudaStream_t)stream;
    unsigned long long *tmp = reinterpret_cast<unsigned long long *>(d_dt);    // the 8 bytes of the result double as the
    KB_CUDA(cudaMemsetAsync(tmp, 0xff, sizeof(unsigned long long), s));          // atomic-min accumulator: nothing is allocated
    const double c = material->rho * material->c_v / (2.0 * material->k);
    const int g = kb_grid(nelem, 256, 8);
    if (ndim == 2) k_dt_rule<2><<<g, 256, 0, s>>>(d_points, d_elements, nelem, c, tmp);
    else k_dt_rule<1><<<g, 256, 0, s>>>(d_points, d_elements, nelem, c, tmp);
    KB_LAUNCH_CHECK();
    k_dt_finish<<<1, 1, 0, s>>>(tmp, factor, d_dt);
    KB_LAUNCH_CHECK();
    return 0;
}

extern "C" int kb_transient_matrix(int64_t ndof, int64_t nnz, const int32_t *d_indptr, const int32_t *d_indices,
                                   const double *d_K, const double *d_M, const int32_t *d_renum, double theta_dt,
                                   double *d_Kff, double *d_A, void *stream) {
    if (!d_indptr || !d_indices || !d_K || !d_M || !d_renum || !d_Kff || !d_A || ndof < 0 || nnz < 0) return KB_EINVAL;
    if (ndof == 0) return 0;
    k_transient_matrix<<<kb_grid(ndof, 256, 8), 256, 0, (cudaStream_t)stream>>>(ndof, d_indptr, d_indices, d_K, d_M,
                                                                                d_renum, theta_dt, d_Kff, d_A);
    KB_LAUNCH_CHECK();
    return 0;
}

extern "C" int kb_mask_free(int64_t ndof, const int32_t *d_renum, const double *d_in, double *d_out, void *stream) {
    if (!d_renum || !d_in || !d_out || ndof < 0) return KB_EINVAL;
    if (ndof == 0) return 0;
    k_mask_free<<<kb_grid(ndof, 256, 8), 256, 0, (cudaStream_t)stream>>>(ndof, d_renum, d_in, d_out);
    KB_LAUNCH_CHECK();
    return 0;
}

extern "C" int kb_masked_matrix(int64_t ndof, int64_t nnz, const int32_t *d_indptr, const int32_t *d_indices,
                                const double *d_K, const double *d_M, const int32_t *d_renum, double c_k, double c_m,
                                int identity_fixed, double *d_out, void *stream) {
    if (!d_indptr || !d_indices || !d_K || !d_renum || !d_out || ndof < 0 || nnz < 0) return KB_EINVAL;
    if (ndof == 0) return 0;
    k_masked_matrix<<<kb_grid(ndof, 256, 8), 256, 0, (cudaStream_t)stream>>>(ndof, d_indptr, d_indices, d_K, d_M, d_renum,
                                                                             c_k, c_m, identity_fixed, d_out);
    KB_LAUNCH_CHECK();
    return 0;
}
