// K3 (seam form): CSR values from materialised element matrices through the row-owner gather map.
// Replaces the scatter-add of FiniteElements/Assembly.py:60-61 (convection), :65-66 (mass), :69-72 (Flux).
// One thread per CSR entry sums its contributors in ascending element order -- the order of the reference's
// sequential `+=` -- so the result is deterministic and free of atomics.
#include "common.cuh"

template <bool MASS>
__global__ void __launch_bounds__(256) k_gather_values(int64_t nnz, const int32_t *__restrict__ seg_ptr,
                                                       const int32_t *__restrict__ gather_src,
                                                       const double *__restrict__ Ke, const double *__restrict__ Me,
                                                       double *__restrict__ K, double *__restrict__ M) {
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t u = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; u < nnz; u += stride) {
        const int p0 = __ldg(seg_ptr + u), p1 = __ldg(seg_ptr + u + 1);
        double k = 0.0, m = 0.0;
        for (int p = p0; p < p1; ++p) {
            const int src = __ldg(gather_src + p);
            k += __ldg(Ke + src);
            if (MASS) m += __ldg(Me + src);
        }
        K[u] = k;
        if (MASS) M[u] = m;
    }
}

// Flux[node] = sum over incident elements (ascending) of Fe[e][local node]: the contributors of the diagonal entry
// (node, node) are exactly the incident elements, with src = e*npe^2 + a*npe + a.
__global__ void __launch_bounds__(256) k_gather_flux(int64_t nnode, int npe, const int32_t *__restrict__ seg_ptr,
                                                     const int32_t *__restrict__ gather_src,
                                                     const int32_t *__restrict__ diag_pos,
                                                     const double *__restrict__ Fe, double *__restrict__ Flux) {
    const int cap = npe * npe;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < nnode; i += stride) {
        const int u = __ldg(diag_pos + i);
        double f = 0.0;
        if (u >= 0) {
            const int p0 = __ldg(seg_ptr + u), p1 = __ldg(seg_ptr + u + 1);
            for (int p = p0; p < p1; ++p) {
                const int src = __ldg(gather_src + p);
                const int e = src / cap, a = (src - e * cap) / (npe + 1);
                f += __ldg(Fe + (int64_t)e * npe + a);
            }
        }
        Flux[i] = f;
    }
}

extern "C" int kb_assemble_from_elements(int npe, int64_t nnode, int64_t nnz, const int32_t *d_seg_ptr,
                                         const int32_t *d_gather_src, const int32_t *d_diag_pos, const double *d_Ke,
                                         const double *d_Me, const double *d_Fe, double *d_K, double *d_M,
                                         double *d_Flux, void *stream) {
    if (!d_seg_ptr || !d_gather_src || !d_Ke || !d_K || (npe != 2 && npe != 4)) return KB_EINVAL;
    if ((d_Me == nullptr) != (d_M == nullptr)) return KB_EINVAL;
    cudaStream_t s = (cudaStream_t)stream;
    if (nnz > 0) {
        const int g = kb_grid(nnz, 256, 8);
        if (d_M) k_gather_values<true><<<g, 256, 0, s>>>(nnz, d_seg_ptr, d_gather_src, d_Ke, d_Me, d_K, d_M);
        else k_gather_values<false><<<g, 256, 0, s>>>(nnz, d_seg_ptr, d_gather_src, d_Ke, d_Me, d_K, d_M);
        KB_LAUNCH_CHECK();
    }
    if (d_Flux) {
        if (!d_Fe || !d_diag_pos) return KB_EINVAL;
        k_gather_flux<<<kb_grid(nnode, 256, 8), 256, 0, s>>>(nnode, npe, d_seg_ptr, d_gather_src, d_diag_pos, d_Fe, d_Flux);
        KB_LAUNCH_CHECK();
    }
    return 0;
}
