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
    const int32_t *prev;                                 // renumbering of the previous call (may be null)
    int *changed;                                        // set to 1 when the new renumbering differs from `prev`
    __device__ void operator()(int64_t i, int ex, int is_free) const {
        int32_t v;
        if (is_free) {
            v = ex;
            cols_in[ex] = i;
        } else {
            const int64_t k = i - ex;                    // fixed DOFs before i
            v = (int32_t)(-1 - k);
            cols_out[k] = i;
            applied[k] = flags[i];
        }
        renum[i] = v;
        if (prev != nullptr && prev[i] != v) *changed = 1;       // benign race: every writer stores 1
    }
};

__global__ void __launch_bounds__(256) k_neumann(const double *__restrict__ flags, int64_t n, double *__restrict__ F) {
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        const double f = flags[i];
        F[i] = isnan(f) ? 0.0 : f;
    }
}

// reduced-row lengths: cnt[renum[i]] = #free columns in free row i ; cnt[n_free] = 0.  row_mask (optional): bit k of
// row_mask[i] = the k-th entry of row i survives (free row, free column); *nomask = 1 when some free row has more than
// 32 entries (then the masks cannot drive the numeric phase).
__global__ void __launch_bounds__(256) k_reduced_count(int64_t ndof, const int32_t *__restrict__ indptr,
                                                       const int32_t *__restrict__ indices,
                                                       const int32_t *__restrict__ renum, int64_t n_free,
                                                       int32_t *__restrict__ cnt, uint32_t *__restrict__ row_mask,
                                                       int *__restrict__ nomask) {
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    bool wide = false;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < ndof; i += stride) {
        const int r = renum[i];
        if (r < 0) { if (row_mask) row_mask[i] = 0u; continue; }
        int c = 0;
        uint32_t m = 0u;
        const int p0 = indptr[i], pe = indptr[i + 1];
        if (pe - p0 > 32) wide = true;
        for (int p = p0; p < pe; ++p) {
            const int keep = (__ldg(renum + __ldg(indices + p)) >= 0);
            c += keep;
            if (keep && p - p0 < 32) m |= 1u << (p - p0);
        }
        cnt[r] = c;
        if (row_mask) row_mask[i] = m;
    }
    if (blockIdx.x == 0 && threadIdx.x == 0) cnt[n_free] = 0;
    if (wide && nomask) *nomask = 1;
}

// reduced column indices (symbolic phase, once per boundary-condition set-up): one thread per free row
__global__ void __launch_bounds__(256) k_reduced_indices(int64_t ndof, const int32_t *__restrict__ indptr,
                                                         const int32_t *__restrict__ indices,
                                                         const int32_t *__restrict__ renum,
                                                         const int32_t *__restrict__ kb_indptr,
                                                         int32_t *__restrict__ kb_indices) {
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < ndof; i += stride) {
        const int r = renum[i];
        if (r < 0) continue;
        int o = kb_indptr[r];
        for (int p = indptr[i], pe = indptr[i + 1]; p < pe; ++p) {
            const int rc = __ldg(renum + __ldg(indices + p));
            if (rc >= 0) kb_indices[o++] = rc;
        }
    }
}

// Numeric phase driven by the cached row masks: no column index is read except in the (few) rows that lose an entry.
// A CTA takes NR_ROWS consecutive rows; consecutive threads walk consecutive entries of K (coalesced loads), each entry
// finds its row through a byte map in shared memory, tests its bit of the row's mask and -- when it survives -- goes to
// kb_indptr[renum[row]] + popcount(mask below it): destinations of consecutive survivors are consecutive, so the stores
// coalesce without a compaction pass.  HBM traffic: K in, K_b out, 4 + 4 + 4 bytes per row (+ F / F_b).
constexpr int NR_ROWS = 256, NR_CAP = 4096;
template <bool MASS>
__global__ void __launch_bounds__(NR_ROWS) k_reduced_numeric(int64_t ndof, const int32_t *__restrict__ indptr,
                                                            const int32_t *__restrict__ indices,
                                                            const double *__restrict__ K, const double *__restrict__ M,
                                                            const int32_t *__restrict__ renum,
                                                            const double *__restrict__ applied, double load_factor,
                                                            double *__restrict__ F, const uint32_t *__restrict__ row_mask,
                                                            const int32_t *__restrict__ kb_indptr,
                                                            double *__restrict__ kb_data, double *__restrict__ mb_data,
                                                            double *__restrict__ Fb) {
    __shared__ int s_ip[NR_ROWS + 1];
    __shared__ int s_q[NR_ROWS];
    __shared__ uint32_t s_mask[NR_ROWS];
    __shared__ uint8_t s_row[NR_CAP];
    const int tid = threadIdx.x;
    const int64_t nblocks = (ndof + NR_ROWS - 1) / NR_ROWS;
    for (int64_t blk = blockIdx.x; blk < nblocks; blk += gridDim.x) {
        const int64_t r0 = blk * NR_ROWS;
        const int nr = (int)((ndof - r0) < NR_ROWS ? (ndof - r0) : NR_ROWS);
        for (int i = tid; i <= nr; i += NR_ROWS) s_ip[i] = indptr[r0 + i];
        int r = -1;
        uint32_t m = 0u;
        if (tid < nr) {
            r = renum[r0 + tid];
            m = r >= 0 ? row_mask[r0 + tid] : 0u;
            s_mask[tid] = m;
            s_q[tid] = r >= 0 ? kb_indptr[r] : 0;
        }
        __syncthreads();
        const int p0 = s_ip[0], np = s_ip[nr] - p0;
        if (kb_data != nullptr) {
            if (np <= NR_CAP) {
                if (tid < nr)
                    for (int i = s_ip[tid] - p0, ie = s_ip[tid + 1] - p0; i < ie; ++i) s_row[i] = (uint8_t)tid;
                __syncthreads();
                // four entries per thread in flight: the loads are issued unconditionally (every i < np is a valid entry of
                // K) before the mask tests, which the compiler would otherwise serialise load -> store -> load
                for (int i0 = tid; i0 < np; i0 += 4 * NR_ROWS) {
                    double kv[4], mv[4];
#pragma unroll
                    for (int u = 0; u < 4; ++u) {
                        const int i = i0 + u * NR_ROWS;
                        kv[u] = i < np ? ld_stream(K + p0 + i) : 0.0;
                        mv[u] = (MASS && i < np) ? ld_stream(M + p0 + i) : 0.0;
                    }
#pragma unroll
                    for (int u = 0; u < 4; ++u) {
                        const int i = i0 + u * NR_ROWS;
                        if (i < np) {
                            const int row = s_row[i];
                            const int k = i - (s_ip[row] - p0);
                            const uint32_t mm = s_mask[row];
                            if ((mm >> k) & 1u) {
                                const int dest = s_q[row] + __popc(mm & ((1u << k) - 1u));
                                kb_data[dest] = kv[u];
                                if (MASS) mb_data[dest] = mv[u];
                            }
                        }
                    }
                }
            } else if (r >= 0) {                                           // oversized block: per-row walk
                int o = s_q[tid];
                for (int p = s_ip[tid], k = 0; p < s_ip[tid + 1]; ++p, ++k)
                    if ((m >> k) & 1u) { kb_data[o] = K[p]; if (MASS) mb_data[o] = M[p]; ++o; }
            }
        }
        // right-hand side: F[in] -= K[in, out] (T_D LoadFactor)   (BoundaryCondition.py:182-184), F_b = F[in]   (:190)
        if (r >= 0) {
            const int pa = s_ip[tid], len = s_ip[tid + 1] - pa;
            const uint32_t full = len >= 32 ? 0xffffffffu : ((1u << len) - 1u);
            double f = F[r0 + tid];
            if (m != full) {
                double s = 0.0;
                for (int k = 0; k < len; ++k)
                    if (!((m >> k) & 1u)) {
                        const int rc = __ldg(renum + __ldg(indices + pa + k));
                        const double td = __ldg(applied + (-1 - rc));
                        if (!(fabs(td) <= 1.0e-8))                                  // ~np.isclose(T_D, 0)        :182
                            s = __dadd_rn(s, __dmul_rn(K[pa + k], td));             // csr mat-vec, ascending column
                    }
                f = __dadd_rn(f, -__dmul_rn(s, load_factor));
                F[r0 + tid] = f;
            }
            Fb[r] = f;
        }
        __syncthreads();
    }
}

struct CntIn {
    const int32_t *c;
    __device__ int operator()(int64_t i) const { return c[i]; }
};
struct CntOut {
    int32_t *c;
    __device__ void operator()(int64_t i, int ex, int) const { c[i] = ex; }
};

// Fill pass of the reduction, coalesced: a CTA takes FB_ROWS consecutive rows.  Phase 1 streams the rows' entries (K,
// indices, the renumbering of every column) into shared memory with consecutive threads on consecutive entries; phase 2 is
// one thread per row -- compaction in ascending column order, RHS elimination as a sequential sum over the fixed columns
// (csr mat-vec order) -- entirely in shared memory; phase 3 streams the compacted block, which is contiguous in K_b, back
// out.  Blocks whose entries do not fit FB_CAP walk their rows in global memory instead.  (A plain one-thread-per-row
// kernel reads 72-byte row chunks with 72-byte lane strides: 2.4 ms at 16 M rows for ~0.6 ms worth of traffic.)
constexpr int FB_ROWS = 256;
constexpr int FB_CAP = 2560;
template <bool MASS>
__global__ void __launch_bounds__(FB_ROWS) k_reduced_fill_flat(int64_t ndof, const int32_t *__restrict__ indptr,
                                                             const int32_t *__restrict__ indices,
                                                             const double *__restrict__ K, const double *__restrict__ M,
                                                             const int32_t *__restrict__ renum,
                                                             const double *__restrict__ applied, double load_factor,
                                                             double *__restrict__ F, const int32_t *__restrict__ kb_indptr,
                                                             int32_t *__restrict__ kb_indices, double *__restrict__ kb_data,
                                                             double *__restrict__ mb_data, double *__restrict__ Fb) {
    extern __shared__ __align__(16) unsigned char fb_raw[];
    double *s_v = reinterpret_cast<double *>(fb_raw);                      // [FB_CAP] staged K values
    double *s_ov = s_v + FB_CAP;                                           // [FB_CAP] compacted K values
    double *s_m = s_ov + FB_CAP;                                           // [FB_CAP] staged M values   (MASS)
    double *s_om = s_m + (MASS ? FB_CAP : 0);                              // [FB_CAP] compacted M values (MASS)
    int *s_rc = reinterpret_cast<int *>(s_om + (MASS ? FB_CAP : 0));       // [FB_CAP] renum of the column
    int *s_oc = s_rc + FB_CAP;                                             // [FB_CAP] compacted reduced column
    __shared__ int s_ip[FB_ROWS + 1];
    __shared__ int s_qbase, s_qend;
    const int tid = threadIdx.x;
    const int64_t nblocks = (ndof + FB_ROWS - 1) / FB_ROWS;
    for (int64_t blk = blockIdx.x; blk < nblocks; blk += gridDim.x) {
        const int64_t r0 = blk * FB_ROWS;
        const int nr = (int)((ndof - r0) < FB_ROWS ? (ndof - r0) : FB_ROWS);
        for (int i = tid; i <= nr; i += FB_ROWS) s_ip[i] = indptr[r0 + i];
        if (tid == 0) { s_qbase = -1; s_qend = -1; }
        __syncthreads();
        const int p0 = s_ip[0], np = s_ip[nr] - p0;
        const int r = tid < nr ? renum[r0 + tid] : -1;                     // reduced index of this thread's row (or fixed)
        const int q = r >= 0 ? kb_indptr[r] : 0;
        if (r >= 0) {                                                      // first / last free row of the block: output extent
            atomicMin(reinterpret_cast<unsigned int *>(&s_qbase), (unsigned int)q);      // -1 = 0xffffffff: any q is smaller
            atomicMax(&s_qend, kb_indptr[r + 1]);
        }
        if (np <= FB_CAP) {
            for (int i = tid; i < np; i += FB_ROWS) {
                s_v[i] = K[p0 + i];
                if (MASS) s_m[i] = M[p0 + i];
                s_rc[i] = __ldg(renum + __ldg(indices + p0 + i));
            }
            __syncthreads();
            if (r >= 0) {
                int o = q - s_qbase;
                double s = 0.0;
                for (int i = s_ip[tid] - p0, ie = s_ip[tid + 1] - p0; i < ie; ++i) {
                    const int rc = s_rc[i];
                    const double v = s_v[i];
                    if (rc >= 0) {
                        s_oc[o] = rc; s_ov[o] = v;
                        if (MASS) s_om[o] = s_m[i];
                        ++o;
                    } else {
                        const double td = __ldg(applied + (-1 - rc));
                        if (!(fabs(td) <= 1.0e-8))                                  // ~np.isclose(T_D, 0)        :182
                            s = __dadd_rn(s, __dmul_rn(v, td));                     // csr mat-vec, ascending column
                    }
                }
                const double f = __dadd_rn(F[r0 + tid], -__dmul_rn(s, load_factor));   // :183-184
                F[r0 + tid] = f;
                Fb[r] = f;                                                          // :190
            }
            __syncthreads();
            if (s_qend > s_qbase) {
                const int qb = s_qbase, no = s_qend - qb;
                for (int i = tid; i < no; i += FB_ROWS) {
                    kb_indices[qb + i] = s_oc[i];
                    kb_data[qb + i] = s_ov[i];
                    if (MASS) mb_data[qb + i] = s_om[i];
                }
            }
        } else if (r >= 0) {                                                // oversized block: per-row walk in global memory
            int o = q;
            double s = 0.0;
            for (int p = s_ip[tid], pe = s_ip[tid + 1]; p < pe; ++p) {
                const int rc = __ldg(renum + __ldg(indices + p));
                const double v = K[p];
                if (rc >= 0) {
                    kb_indices[o] = rc; kb_data[o] = v;
                    if (MASS) mb_data[o] = M[p];
                    ++o;
                } else {
                    const double td = __ldg(applied + (-1 - rc));
                    if (!(fabs(td) <= 1.0e-8)) s = __dadd_rn(s, __dmul_rn(v, td));
                }
            }
            const double f = __dadd_rn(F[r0 + tid], -__dmul_rn(s, load_factor));
            F[r0 + tid] = f;
            Fb[r] = f;
        }
        __syncthreads();
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

extern "C" size_t kb_dirichlet_split_workspace_bytes(int64_t ndof) {
    return sizeof(int) * (size_t)(kbscan::workspace_ints(ndof > 0 ? ndof : 1) + 4);
}

extern "C" int kb_dirichlet_split(const double *d_flags, int64_t ndof, const int32_t *d_prev_renum, int32_t *d_renum,
                                  int64_t *d_columns_in, int64_t *d_columns_out, double *d_applied, void *d_workspace,
                                  size_t workspace_bytes, int64_t *h_counts, void *stream) {
    if (!d_flags || !d_renum || !d_columns_in || !d_columns_out || !d_applied || !h_counts || !d_workspace || ndof < 0 ||
        d_prev_renum == d_renum)
        return KB_EINVAL;
    if (ndof >= INT32_MAX) return KB_ETOOBIG;
    if (workspace_bytes < kb_dirichlet_split_workspace_bytes(ndof)) return KB_EWORKSPACE;
    cudaStream_t s = (cudaStream_t)stream;
    if (ndof == 0) { h_counts[0] = h_counts[1] = 0; h_counts[2] = d_prev_renum ? 1 : 0; return 0; }
    int *ws = (int *)d_workspace;
    const int64_t nt = kbscan::num_tiles(ndof);
    int *changed = ws + nt + 1;
    KB_CUDA(cudaMemsetAsync(changed, 0, sizeof(int), s));
    int rc = kbscan::exclusive_scan(FreeIn{d_flags}, FreeOut{d_flags, d_renum, d_columns_in, d_columns_out, d_applied,
                                                            d_prev_renum, changed}, ndof, ws, s);
    if (rc) return rc;
    int h[2] = {0, 0};                                   // {n_free, changed}: adjacent words, one copy
    KB_CUDA(cudaMemcpyAsync(h, ws + nt, sizeof(int) * 2, cudaMemcpyDeviceToHost, s));
    KB_CUDA(cudaStreamSynchronize(s));
    h_counts[0] = h[0];
    h_counts[1] = ndof - h[0];
    h_counts[2] = (d_prev_renum != nullptr && h[1] == 0) ? 1 : 0;
    return 0;
}

extern "C" int kb_neumann_fluxes(const double *d_flags, int64_t ndof, double *d_F, void *stream) {
    if (!d_flags || !d_F || ndof < 0) return KB_EINVAL;
    if (ndof == 0) return 0;
    k_neumann<<<kb_grid(ndof, 256, 8), 256, 0, (cudaStream_t)stream>>>(d_flags, ndof, d_F);
    KB_LAUNCH_CHECK();
    return 0;
}

// ---- ApplyDirichletGetReducedMatrices in two phases.  Symbolic (once per pattern + Dirichlet set): reduced row pointers,
// reduced column indices, one 32-bit survival mask per row.  Numeric (every step): values, right-hand side.
extern "C" size_t kb_dirichlet_reduce_workspace_bytes(int64_t ndof, int64_t n_free) {
    (void)ndof;
    return sizeof(int) * (size_t)(kbscan::workspace_ints(n_free + 1) + 4);
}

extern "C" int kb_dirichlet_reduce_symbolic(int64_t ndof, const int32_t *d_indptr, const int32_t *d_indices,
                                            const int32_t *d_renum, int64_t n_free, uint32_t *d_row_mask,
                                            int32_t *d_Kb_indptr, int32_t *d_Kb_indices, void *d_workspace,
                                            size_t workspace_bytes, int64_t *h_info, void *stream) {
    if (!d_indptr || !d_indices || !d_renum || !d_row_mask || !d_Kb_indptr || !d_workspace || !h_info || ndof < 0 ||
        n_free < 0 || n_free > ndof)
        return KB_EINVAL;
    if (n_free > 0 && !d_Kb_indices) return KB_EINVAL;
    if (workspace_bytes < kb_dirichlet_reduce_workspace_bytes(ndof, n_free)) return KB_EWORKSPACE;
    cudaStream_t s = (cudaStream_t)stream;
    int *ws = (int *)d_workspace;
    const int64_t nt = kbscan::num_tiles(n_free + 1);
    int *nomask = ws + nt + 1;
    KB_CUDA(cudaMemsetAsync(nomask, 0, sizeof(int), s));
    const int g = kb_grid(ndof > 0 ? ndof : 1, 256, 8);
    k_reduced_count<<<g, 256, 0, s>>>(ndof, d_indptr, d_indices, d_renum, n_free, d_Kb_indptr, d_row_mask, nomask);
    KB_LAUNCH_CHECK();
    int rc = kbscan::exclusive_scan(CntIn{d_Kb_indptr}, CntOut{d_Kb_indptr}, n_free + 1, ws, s);
    if (rc) return rc;
    if (ndof > 0 && n_free > 0) {
        k_reduced_indices<<<g, 256, 0, s>>>(ndof, d_indptr, d_indices, d_renum, d_Kb_indptr, d_Kb_indices);
        KB_LAUNCH_CHECK();
    }
    int h[2] = {0, 0};                                   // {nnz_b, nomask}
    KB_CUDA(cudaMemcpyAsync(h, ws + nt, sizeof(int) * 2, cudaMemcpyDeviceToHost, s));
    KB_CUDA(cudaStreamSynchronize(s));
    h_info[0] = h[0];
    h_info[1] = h[1] ? 0 : 1;                            // 1: the masks can drive kb_dirichlet_reduce_numeric
    return 0;
}

// d_Kb_data == NULL: right-hand side only (F, F_b); no host synchronisation, nothing allocated
extern "C" int kb_dirichlet_reduce_numeric(int64_t ndof, const int32_t *d_indptr, const int32_t *d_indices, const double *d_K,
                                           const double *d_M, const int32_t *d_renum, const double *d_applied,
                                           double load_factor, double *d_F, int64_t n_free, const uint32_t *d_row_mask,
                                           const int32_t *d_Kb_indptr, double *d_Kb_data, double *d_Mb_data, double *d_Fb,
                                           void *stream) {
    if (!d_indptr || !d_indices || !d_K || !d_renum || !d_F || !d_row_mask || !d_Kb_indptr || ndof < 0 || n_free < 0 ||
        n_free > ndof)
        return KB_EINVAL;
    if (n_free > 0 && !d_Fb) return KB_EINVAL;
    if (n_free < ndof && !d_applied) return KB_EINVAL;
    if ((d_M == nullptr) != (d_Mb_data == nullptr) || (d_Mb_data != nullptr && d_Kb_data == nullptr)) return KB_EINVAL;
    if (ndof == 0) return 0;
    cudaStream_t s = (cudaStream_t)stream;
    const int gf = kb_grid(kb_cdiv(ndof, NR_ROWS) * NR_ROWS, NR_ROWS, 8);
    if (d_M) k_reduced_numeric<true><<<gf, NR_ROWS, 0, s>>>(ndof, d_indptr, d_indices, d_K, d_M, d_renum, d_applied, load_factor,
                                                          d_F, d_row_mask, d_Kb_indptr, d_Kb_data, d_Mb_data, d_Fb);
    else k_reduced_numeric<false><<<gf, NR_ROWS, 0, s>>>(ndof, d_indptr, d_indices, d_K, d_M, d_renum, d_applied, load_factor,
                                                       d_F, d_row_mask, d_Kb_indptr, d_Kb_data, d_Mb_data, d_Fb);
    KB_LAUNCH_CHECK();
    return 0;
}

// One-shot form (any row length; recomputes the structure on every call): count + scan + block-staged fill.
extern "C" size_t kb_apply_dirichlet_reduce_workspace_bytes(int64_t ndof, int64_t n_free) {
    return kb_dirichlet_reduce_workspace_bytes(ndof, n_free);
}
extern "C" int kb_apply_dirichlet_reduce(int64_t ndof, const int32_t *d_indptr, const int32_t *d_indices,
                                         const double *d_K, const double *d_M, const int32_t *d_renum,
                                         const double *d_applied, double load_factor, double *d_F, int64_t n_free,
                                         int32_t *d_Kb_indptr, int32_t *d_Kb_indices, double *d_Kb_data,
                                         double *d_Mb_data, double *d_Fb, void *d_workspace, size_t workspace_bytes,
                                         int64_t *h_nnz_b, void *stream) {
    if (!d_indptr || !d_indices || !d_K || !d_renum || !d_F || !d_Kb_indptr || !h_nnz_b || !d_workspace || ndof < 0 ||
        n_free < 0 || n_free > ndof)
        return KB_EINVAL;
    if (n_free > 0 && (!d_Kb_indices || !d_Kb_data || !d_Fb)) return KB_EINVAL;      // empty reduced system: nothing to fill
    if (n_free < ndof && !d_applied) return KB_EINVAL;
    if ((d_M == nullptr) != (d_Mb_data == nullptr)) return KB_EINVAL;
    if (workspace_bytes < kb_apply_dirichlet_reduce_workspace_bytes(ndof, n_free)) return KB_EWORKSPACE;
    cudaStream_t s = (cudaStream_t)stream;
    const int g = kb_grid(ndof > 0 ? ndof : 1, 256, 8);
    k_reduced_count<<<g, 256, 0, s>>>(ndof, d_indptr, d_indices, d_renum, n_free, d_Kb_indptr, nullptr, nullptr);
    KB_LAUNCH_CHECK();
    int *ws = (int *)d_workspace;
    int rc = kbscan::exclusive_scan(CntIn{d_Kb_indptr}, CntOut{d_Kb_indptr}, n_free + 1, ws, s);
    if (rc) return rc;
    int nnzb = 0;
    KB_CUDA(cudaMemcpyAsync(&nnzb, ws + kbscan::num_tiles(n_free + 1), sizeof(int), cudaMemcpyDeviceToHost, s));
    if (ndof > 0) {
        static KbOncePerDevice configured;
        constexpr size_t bytes_m = (size_t)FB_CAP * (4 * 8 + 2 * 4), bytes_s = (size_t)FB_CAP * (2 * 8 + 2 * 4);
        if (configured.need()) {
            KB_CUDA(cudaFuncSetAttribute(k_reduced_fill_flat<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes_m));
            KB_CUDA(cudaFuncSetAttribute(k_reduced_fill_flat<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes_s));
        }
        const int gf = kb_grid(kb_cdiv(ndof, FB_ROWS) * FB_ROWS, FB_ROWS, 4);
        if (d_M) k_reduced_fill_flat<true><<<gf, FB_ROWS, bytes_m, s>>>(ndof, d_indptr, d_indices, d_K, d_M, d_renum, d_applied,
                                                                       load_factor, d_F, d_Kb_indptr, d_Kb_indices, d_Kb_data,
                                                                       d_Mb_data, d_Fb);
        else k_reduced_fill_flat<false><<<gf, FB_ROWS, bytes_s, s>>>(ndof, d_indptr, d_indices, d_K, d_M, d_renum, d_applied,
                                                                    load_factor, d_F, d_Kb_indptr, d_Kb_indices, d_Kb_data,
                                                                    d_Mb_data, d_Fb);
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
