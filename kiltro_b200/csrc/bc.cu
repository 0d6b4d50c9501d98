// K4: boundary conditions on device.  Replaces kiltro/BoundaryCondition.py:
//   :100-122 GetDirichletBoundaryConditions  -> kb_dirichlet_split   (NaN flag = free DOF; stream compaction by scan)
//   :125-140 ComputeNeumannFluxes            -> kb_neumann_fluxes    (nodal point loads)
//   :169-200 ApplyDirichletGetReducedMatrices-> kb_apply_dirichlet_reduce (RHS elimination fused with the physical
//                                               compaction K_b = K[in,:][:,in]; scipy's fancy indexing makes several
//                                               full copies of K, here it is one counting pass + one fill pass)
//   :203-241 UpdateFixDoFs/UpdateFreeDoFs/TotalComponentSol -> kb_total_component_sol
#include "common.cuh"
#include "scan.cuh"

namespace {

struct FreeIn {
    const double *flags;
    __device__ int operator()(int64_t i) const { return isnan(flags[i]) ? 1 : 0; }
};
struct FreeOut {
    const double *flags;
    int32_t *renum;
    int64_t *cols_in, *cols_out;
    double *applied;
    __device__ void operator()(int64_t i, int ex, int is_free) const {
        if (is_free) {
            renum[i] = ex;
            cols_in[ex] = i;
        } else {
            const int64_t k = i - ex;                    // fixed DOFs before i
            renum[i] = (int32_t)(-1 - k);
            cols_out[k] = i;
            applied[k] = flags[i];
        }
    }
};

__global__ void __launch_bounds__(256) k_neumann(const double *__restrict__ flags, int64_t n, double *__restrict__ F) {
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        const double f = flags[i];
        F[i] = isnan(f) ? 0.0 : f;
    }
}

// reduced-row lengths: cnt[renum[i]] = #free columns in free row i ; cnt[n_free] = 0
__global__ void __launch_bounds__(256) k_reduced_count(int64_t ndof, const int32_t *__restrict__ indptr,
                                                       const int32_t *__restrict__ indices,
                                                       const int32_t *__restrict__ renum, int64_t n_free,
                                                       int32_t *__restrict__ cnt) {
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < ndof; i += stride) {
        const int r = renum[i];
        if (r < 0) continue;
        int c = 0;
        for (int p = indptr[i], pe = indptr[i + 1]; p < pe; ++p) c += (__ldg(renum + __ldg(indices + p)) >= 0);
        cnt[r] = c;
    }
    if (blockIdx.x == 0 && threadIdx.x == 0) cnt[n_free] = 0;
}

struct CntIn {
    const int32_t *c;
    __device__ int operator()(int64_t i) const { return c[i]; }
};
struct CntOut {
    int32_t *c;
    __device__ void operator()(int64_t i, int ex, int) const { c[i] = ex; }
};

template <bool MASS>
__global__ void __launch_bounds__(256) k_reduced_fill(int64_t ndof, const int32_t *__restrict__ indptr,
                                                      const int32_t *__restrict__ indices,
                                                      const double *__restrict__ K, const double *__restrict__ M,
                                                      const int32_t *__restrict__ renum,
                                                      const double *__restrict__ applied, double load_factor,
                                                      double *__restrict__ F, const int32_t *__restrict__ kb_indptr,
                                                      int32_t *__restrict__ kb_indices, double *__restrict__ kb_data,
                                                      double *__restrict__ mb_data, double *__restrict__ Fb) {
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < ndof; i += stride) {
        const int r = renum[i];
        if (r < 0) continue;
        int q = kb_indptr[r];
        double s = 0.0;
        for (int p = indptr[i], pe = indptr[i + 1]; p < pe; ++p) {
            const int col = __ldg(indices + p);
            const int rc = __ldg(renum + col);
            const double v = K[p];
            if (rc >= 0) {
                kb_indices[q] = rc;
                kb_data[q] = v;
                if (MASS) mb_data[q] = M[p];
                ++q;
            } else {
                const double td = __ldg(applied + (-1 - rc));
                if (!(fabs(td) <= 1.0e-8))                                  // ~np.isclose(T_D, 0)        :182
                    s = __dadd_rn(s, __dmul_rn(v, td));                     // csr mat-vec, ascending column
            }
        }
        const double f = __dadd_rn(F[i], -__dmul_rn(s, load_factor));       // F[in] - (K[in,out] T_D)*LoadFactor  :183-184
        F[i] = f;
        Fb[r] = f;                                                          // :190
    }
}

__global__ void __launch_bounds__(256) k_total_sol(int64_t ndof, const int32_t *__restrict__ renum,
                                                   const double *__restrict__ sol, const double *__restrict__ applied,
                                                   double *__restrict__ T) {
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < ndof; i += stride) {
        const int r = renum[i];
        double v;
        if (r >= 0) v = sol ? sol[r] : 0.0;
        else v = applied ? applied[-1 - r] : 0.0;
        T[i] = v;
    }
}

}  // namespace

extern "C" int kb_dirichlet_split(const double *d_flags, int64_t ndof, int32_t *d_renum, int64_t *d_columns_in,
                                  int64_t *d_columns_out, double *d_applied, int64_t *h_counts, void *stream) {
    if (!d_flags || !d_renum || !d_columns_in || !d_columns_out || !d_applied || !h_counts || ndof < 0) return KB_EINVAL;
    if (ndof >= INT32_MAX) return KB_ETOOBIG;
    cudaStream_t s = (cudaStream_t)stream;
    if (ndof == 0) { h_counts[0] = h_counts[1] = 0; return 0; }
    int *ws;
    const int64_t nws = kbscan::workspace_ints(ndof);
    KB_CUDA(cudaMallocAsync((void **)&ws, sizeof(int) * nws, s));
    int rc = kbscan::exclusive_scan(FreeIn{d_flags}, FreeOut{d_flags, d_renum, d_columns_in, d_columns_out, d_applied},
                                    ndof, ws, s);
    if (rc) return rc;
    int nfree = 0;
    KB_CUDA(cudaMemcpyAsync(&nfree, ws + kbscan::num_tiles(ndof), sizeof(int), cudaMemcpyDeviceToHost, s));
    KB_CUDA(cudaFreeAsync(ws, s));
    KB_CUDA(cudaStreamSynchronize(s));
    h_counts[0] = nfree;
    h_counts[1] = ndof - nfree;
    return 0;
}

extern "C" int kb_neumann_fluxes(const double *d_flags, int64_t ndof, double *d_F, void *stream) {
    if (!d_flags || !d_F || ndof < 0) return KB_EINVAL;
    if (ndof == 0) return 0;
    k_neumann<<<kb_grid(ndof, 256, 8), 256, 0, (cudaStream_t)stream>>>(d_flags, ndof, d_F);
    KB_LAUNCH_CHECK();
    return 0;
}

extern "C" int kb_apply_dirichlet_reduce(int64_t ndof, const int32_t *d_indptr, const int32_t *d_indices,
                                         const double *d_K, const double *d_M, const int32_t *d_renum,
                                         const double *d_applied, double load_factor, double *d_F, int64_t n_free,
                                         int32_t *d_Kb_indptr, int32_t *d_Kb_indices, double *d_Kb_data,
                                         double *d_Mb_data, double *d_Fb, int64_t *h_nnz_b, void *stream) {
    if (!d_indptr || !d_indices || !d_K || !d_renum || !d_F || !d_Kb_indptr || !h_nnz_b || ndof < 0 || n_free < 0 ||
        n_free > ndof)
        return KB_EINVAL;
    if (n_free > 0 && (!d_Kb_indices || !d_Kb_data || !d_Fb)) return KB_EINVAL;      // empty reduced system: nothing to fill
    if (n_free < ndof && !d_applied) return KB_EINVAL;
    if ((d_M == nullptr) != (d_Mb_data == nullptr)) return KB_EINVAL;
    cudaStream_t s = (cudaStream_t)stream;
    const int g = kb_grid(ndof > 0 ? ndof : 1, 256, 8);
    k_reduced_count<<<g, 256, 0, s>>>(ndof, d_indptr, d_indices, d_renum, n_free, d_Kb_indptr);
    KB_LAUNCH_CHECK();
    int *ws;
    const int64_t nws = kbscan::workspace_ints(n_free + 1);
    KB_CUDA(cudaMallocAsync((void **)&ws, sizeof(int) * nws, s));
    int rc = kbscan::exclusive_scan(CntIn{d_Kb_indptr}, CntOut{d_Kb_indptr}, n_free + 1, ws, s);
    if (rc) return rc;
    int nnzb = 0;
    KB_CUDA(cudaMemcpyAsync(&nnzb, ws + kbscan::num_tiles(n_free + 1), sizeof(int), cudaMemcpyDeviceToHost, s));
    KB_CUDA(cudaFreeAsync(ws, s));
    if (ndof > 0) {
        if (d_M) k_reduced_fill<true><<<g, 256, 0, s>>>(ndof, d_indptr, d_indices, d_K, d_M, d_renum, d_applied, load_factor,
                                                        d_F, d_Kb_indptr, d_Kb_indices, d_Kb_data, d_Mb_data, d_Fb);
        else k_reduced_fill<false><<<g, 256, 0, s>>>(ndof, d_indptr, d_indices, d_K, d_M, d_renum, d_applied, load_factor,
                                                     d_F, d_Kb_indptr, d_Kb_indices, d_Kb_data, d_Mb_data, d_Fb);
        KB_LAUNCH_CHECK();
    }
    KB_CUDA(cudaStreamSynchronize(s));
    *h_nnz_b = nnzb;
    return 0;
}

extern "C" int kb_total_component_sol(int64_t ndof, const int32_t *d_renum, const double *d_sol,
                                      const double *d_applied, double *d_T, void *stream) {
    if (!d_renum || !d_T || ndof < 0) return KB_EINVAL;
    if (ndof == 0) return 0;
    k_total_sol<<<kb_grid(ndof, 256, 8), 256, 0, (cudaStream_t)stream>>>(ndof, d_renum, d_sol, d_applied, d_T);
    KB_LAUNCH_CHECK();
    return 0;
}
