// K5: CSR SpMV y = A x (fp64 values, int32 indices) with optional fused dot products.
//
// "CSR-stream" formulation for short FE rows (3 nnz/row for line-2, 9 for quad-4): a CTA owns a contiguous block of rows
// holding ~SPMV_NNZ non-zeros (kb_spmv_partition), streams values and column indices fully coalesced (nnz-parallel),
// stages the products val*x[col] in shared memory, then one thread per row adds its products left to right (fixed
// order -> bitwise reproducible).  x is gathered through L1/L2: a row block touches three short contiguous bands of x.
// Row blocks whose nnz exceed the staging buffer (a very long row) fall back to a CTA-cooperative row loop.
//
// Fused epilogue (the Krylov solvers' dot products, krylov.cu): per-CTA partial sums of <w,y> and <y,y> are written to
// a partials array and the last CTA to finish (integer ticket) adds them in a fixed order and calls a finalisation
// functor -- no floating-point atomics anywhere, so dots are deterministic.
#pragma once
#include "common.cuh"

namespace kbspmv {

constexpr int THREADS = 256;
constexpr int NNZ_TARGET = 2304;          // non-zeros per CTA aimed for by the partition (256 quad-4 rows)
constexpr int NNZ_SMEM = 3072;            // staging capacity (24 KB): target + longest row that still takes the fast path

static inline int64_t num_blocks(int64_t nnz) { return nnz > 0 ? (nnz + NNZ_TARGET - 1) / NNZ_TARGET : 1; }

// rowblk[b] = last row r with indptr[r] - indptr[0] <= b*NNZ_TARGET (b = 1..nblk-1), rowblk[0] = 0, rowblk[nblk] = nrows
// (indptr may be a row-range view of a larger CSR: offsets are taken relative to indptr[0])
static __global__ void __launch_bounds__(256) k_partition(int64_t nrows, int64_t nblk, const int32_t *__restrict__ indptr,
                                                   int32_t *__restrict__ rowblk) {
    const int64_t b = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (b > nblk) return;
    if (b == 0) { rowblk[0] = 0; return; }
    if (b == nblk) { rowblk[nblk] = (int32_t)nrows; return; }
    const int64_t target = (int64_t)indptr[0] + b * NNZ_TARGET;
    int64_t lo = 0, hi = nrows;              // invariant: indptr[lo] <= target ; answer in [lo, hi]
    while (lo < hi) {
        const int64_t mid = (lo + hi + 1) >> 1;
        if ((int64_t)indptr[mid] <= target) lo = mid; else hi = mid - 1;
    }
    rowblk[b] = (int32_t)lo;
}

struct NoFinal {
    __device__ void operator()(double, double, double = 0.0, double = 0.0) const {}
};

// DOTS: 0 none, 1 = <w,y> and <y,y>.  Final(d0, d1) runs in exactly one thread after all CTAs contributed.
// `done` (may be null): device flag; when *done != 0 the kernel exits at once (solver already converged).
template <int DOTS, class Final>
__global__ void __launch_bounds__(THREADS) k_spmv(int64_t nrows, const int32_t *__restrict__ indptr,
                                                  const int32_t *__restrict__ indices,
                                                  const double *__restrict__ vals, const int32_t *__restrict__ rowblk,
                                                  const double *__restrict__ x, double *__restrict__ y,
                                                  const double *__restrict__ w, double *__restrict__ partials,
                                                  unsigned int *__restrict__ ticket, const int *__restrict__ done,
                                                  Final fin, const double *__restrict__ w2 = nullptr) {
    if (done != nullptr && *done != 0) return;
    __shared__ double prod[NNZ_SMEM];
    __shared__ double red[32];
    __shared__ int s_last;
    const int r0 = rowblk[blockIdx.x], r1 = rowblk[blockIdx.x + 1];
    const int p0 = indptr[r0], p1 = indptr[r1];
    const int cnt = p1 - p0;
    double d0 = 0.0, d1 = 0.0, d2 = 0.0, d3 = 0.0;
    if (cnt <= NNZ_SMEM) {
        // phase 1: nnz-parallel products, 4 independent loads in flight per thread
        int k = threadIdx.x;
        for (; k + 3 * THREADS < cnt; k += 4 * THREADS) {
            const int c0 = __ldcs(indices + p0 + k), c1 = __ldcs(indices + p0 + k + THREADS);
            const int c2 = __ldcs(indices + p0 + k + 2 * THREADS), c3 = __ldcs(indices + p0 + k + 3 * THREADS);
            const double v0 = __ldcs(vals + p0 + k), v1 = __ldcs(vals + p0 + k + THREADS);
            const double v2 = __ldcs(vals + p0 + k + 2 * THREADS), v3 = __ldcs(vals + p0 + k + 3 * THREADS);
            const double x0 = __ldg(x + c0), x1 = __ldg(x + c1), x2 = __ldg(x + c2), x3 = __ldg(x + c3);
            prod[k] = v0 * x0; prod[k + THREADS] = v1 * x1;
            prod[k + 2 * THREADS] = v2 * x2; prod[k + 3 * THREADS] = v3 * x3;
        }
        for (; k < cnt; k += THREADS) prod[k] = __ldcs(vals + p0 + k) * __ldg(x + __ldcs(indices + p0 + k));
        __syncthreads();
        // phase 2: one thread per row, left-to-right sum
        for (int r = r0 + threadIdx.x; r < r1; r += THREADS) {
            const int a = indptr[r] - p0, b = indptr[r + 1] - p0;
            double s = 0.0;
            for (int q = a; q < b; ++q) s += prod[q];
            y[r] = s;
            if (DOTS) { d0 += __ldg(w + r) * s; d1 += s * s; }
            if (DOTS == 2) { d2 += __ldg(w2 + r) * s; d3 += __ldg(w2 + r) * __ldg(w + r); }
        }
    } else {
        // slow path: the CTA walks its rows one at a time
        for (int r = r0; r < r1; ++r) {
            const int a = indptr[r], b = indptr[r + 1];
            double s = 0.0;
            for (int q = a + threadIdx.x; q < b; q += THREADS) s += vals[q] * __ldg(x + indices[q]);
            s = block_sum(s, red);
            if (threadIdx.x == 0) {
                y[r] = s;
                if (DOTS) { d0 += __ldg(w + r) * s; d1 += s * s; }
                if (DOTS == 2) { d2 += __ldg(w2 + r) * s; d3 += __ldg(w2 + r) * __ldg(w + r); }
            }
        }
    }
    if (DOTS) {
        d0 = block_sum(d0, red);
        d1 = block_sum(d1, red);
        if (DOTS == 2) { d2 = block_sum(d2, red); d3 = block_sum(d3, red); }
        if (threadIdx.x == 0) {
            partials[blockIdx.x] = d0;
            partials[gridDim.x + blockIdx.x] = d1;
            if (DOTS == 2) { partials[2 * gridDim.x + blockIdx.x] = d2; partials[3 * gridDim.x + blockIdx.x] = d3; }
            __threadfence();
            const unsigned int t = atomicInc(ticket, gridDim.x - 1);     // wraps back to 0: self-resetting
            s_last = (t == gridDim.x - 1);
        }
        __syncthreads();
        if (s_last) {
            __threadfence();
            double a0 = 0.0, a1 = 0.0, a2 = 0.0, a3 = 0.0;
            for (int i = threadIdx.x; i < (int)gridDim.x; i += THREADS) {
                a0 += __ldcg(partials + i);
                a1 += __ldcg(partials + gridDim.x + i);
                if (DOTS == 2) { a2 += __ldcg(partials + 2 * gridDim.x + i); a3 += __ldcg(partials + 3 * gridDim.x + i); }
            }
            a0 = block_sum(a0, red);
            a1 = block_sum(a1, red);
            if (DOTS == 2) { a2 = block_sum(a2, red); a3 = block_sum(a3, red); }
            if (threadIdx.x == 0) fin(a0, a1, a2, a3);
        }
    }
}

}  // namespace kbspmv
