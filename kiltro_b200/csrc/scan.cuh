// Device-wide exclusive scan of int32 values (reduce / spine / down-sweep), templated on an input functor
// in(i) -> int and an output functor out(i, exclusive_prefix, value).  Totals must fit int32 (checked by callers).
// Tile = 256 threads x 8 items, each thread owns 8 consecutive items.
#pragma once
#include "common.cuh"

namespace kbscan {

constexpr int THREADS = 256;
constexpr int ITEMS = 8;
constexpr int TILE = THREADS * ITEMS;

static inline int64_t num_tiles(int64_t n) { return n > 0 ? (n + TILE - 1) / TILE : 1; }
// workspace ints needed: one per tile + 1
static inline int64_t workspace_ints(int64_t n) { return num_tiles(n) + 1; }

template <class In>
__global__ void __launch_bounds__(THREADS) k_reduce(In in, int64_t n, int *__restrict__ tile_sum) {
    __shared__ int sm[33];
    const int64_t base = (int64_t)blockIdx.x * TILE + (int64_t)threadIdx.x * ITEMS;
    int s = 0;
#pragma unroll
    for (int k = 0; k < ITEMS; ++k) { const int64_t i = base + k; if (i < n) s += in(i); }
    int total;
    block_scan_excl(s, sm, &total);
    if (threadIdx.x == 0) tile_sum[blockIdx.x] = total;
}

// single block: exclusive scan of tile_sum[0..nt) in place; tile_sum[nt] = grand total
static __global__ void __launch_bounds__(1024) k_spine(int *__restrict__ tile_sum, int64_t nt) {
    __shared__ int sm[33];
    int carry = 0;
    for (int64_t b = 0; b < nt; b += 1024) {
        const int64_t i = b + threadIdx.x;
        const int v = i < nt ? tile_sum[i] : 0;
        int total;
        const int ex = block_scan_excl(v, sm, &total);
        if (i < nt) tile_sum[i] = carry + ex;
        carry += total;
    }
    if (threadIdx.x == 0) tile_sum[nt] = carry;
}

template <class In, class Out>
__global__ void __launch_bounds__(THREADS) k_down(In in, int64_t n, const int *__restrict__ tile_sum, Out out) {
    __shared__ int sm[33];
    const int64_t base = (int64_t)blockIdx.x * TILE + (int64_t)threadIdx.x * ITEMS;
    int v[ITEMS], s = 0;
#pragma unroll
    for (int k = 0; k < ITEMS; ++k) { const int64_t i = base + k; v[k] = i < n ? in(i) : 0; s += v[k]; }
    int total;
    int ex = block_scan_excl(s, sm, &total) + tile_sum[blockIdx.x];
#pragma unroll
    for (int k = 0; k < ITEMS; ++k) { const int64_t i = base + k; if (i < n) out(i, ex, v[k]); ex += v[k]; }
}

// Runs the three kernels.  d_ws: workspace_ints(n) ints; after completion d_ws[num_tiles(n)] holds the total.
template <class In, class Out>
static inline int exclusive_scan(In in, Out out, int64_t n, int *d_ws, cudaStream_t s) {
    const int64_t nt = num_tiles(n);
    k_reduce<<<(unsigned)nt, THREADS, 0, s>>>(in, n, d_ws);
    KB_LAUNCH_CHECK();
    k_spine<<<1, 1024, 0, s>>>(d_ws, nt);
    KB_LAUNCH_CHECK();
    k_down<<<(unsigned)nt, THREADS, 0, s>>>(in, n, d_ws, out);
    KB_LAUNCH_CHECK();
    return 0;
}

}  // namespace kbscan
