// Stable LSD radix sort of (uint64 key, uint32 payload) pairs, 8 bits per pass (histogram / scan / scatter).
// Used only by the symbolic phase (pattern.cu, tileplan.cu): one-off per mesh, not on the numeric hot path.
#pragma once
#include "common.cuh"
#include "scan.cuh"

namespace kbradix {

// ----------------------------------------------------------------------------------------------- LSD radix sort
constexpr int RS_THREADS = 256;
constexpr int RS_WARPS = RS_THREADS / 32;
constexpr int RS_ITEMS = 16;
constexpr int RS_TILE = RS_THREADS * RS_ITEMS;

// per-tile digit histogram -> hist[digit * ntiles + tile]
static __global__ void __launch_bounds__(RS_THREADS) k_rs_hist(const uint64_t *__restrict__ keys, int64_t n, int shift,
                                                        int64_t ntiles, int *__restrict__ hist) {
    __shared__ int h[256];
    h[threadIdx.x] = 0;
    __syncthreads();
    const int64_t base = (int64_t)blockIdx.x * RS_TILE;
#pragma unroll
    for (int k = 0; k < RS_ITEMS; ++k) {
        const int64_t i = base + (int64_t)k * RS_THREADS + threadIdx.x;
        if (i < n) atomicAdd(&h[(int)((keys[i] >> shift) & 0xff)], 1);     // integer, shared memory: order-free
    }
    __syncthreads();
    hist[(int64_t)threadIdx.x * ntiles + blockIdx.x] = h[threadIdx.x];
}

// stable scatter.  Warp w owns the contiguous items [w*512, (w+1)*512) of the tile, visited in 16 chunks of 32.
static __global__ void __launch_bounds__(RS_THREADS) k_rs_scatter(const uint64_t *__restrict__ keys_in,
                                                           const uint32_t *__restrict__ vals_in, int64_t n, int shift,
                                                           int64_t ntiles, const int *__restrict__ hist_scanned,
                                                           uint64_t *__restrict__ keys_out,
                                                           uint32_t *__restrict__ vals_out) {
    __shared__ int cnt[RS_WARPS][256];
    __shared__ int dbase[256];
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    for (int i = threadIdx.x; i < RS_WARPS * 256; i += RS_THREADS) (&cnt[0][0])[i] = 0;
    dbase[threadIdx.x] = hist_scanned[(int64_t)threadIdx.x * ntiles + blockIdx.x];
    __syncthreads();
    const int64_t base = (int64_t)blockIdx.x * RS_TILE + (int64_t)w * (32 * RS_ITEMS) + lane;
    uint64_t key[RS_ITEMS];
    uint32_t val[RS_ITEMS];
    int rank[RS_ITEMS];
    const uint32_t lt = (1u << lane) - 1u;
#pragma unroll
    for (int c = 0; c < RS_ITEMS; ++c) {
        const int64_t i = base + c * 32;
        const bool ok = i < n;
        key[c] = ok ? keys_in[i] : ~0ull;
        val[c] = ok ? vals_in[i] : 0u;
    }
#pragma unroll
    for (int c = 0; c < RS_ITEMS; ++c) {
        const int d = (int)((key[c] >> shift) & 0xff);
        const uint32_t peers = __match_any_sync(KB_FULL, d);
        const int before = __popc(peers & lt);
        const int old = cnt[w][d];                  // same value for every peer (read before the leader's update)
        __syncwarp();
        if (before == 0) cnt[w][d] = old + __popc(peers);
        __syncwarp();
        rank[c] = old + before;
    }
    __syncthreads();
    {   // exclusive scan over warps for digit = threadIdx.x
        int run = 0;
#pragma unroll
        for (int ww = 0; ww < RS_WARPS; ++ww) { const int t = cnt[ww][threadIdx.x]; cnt[ww][threadIdx.x] = run; run += t; }
    }
    __syncthreads();
#pragma unroll
    for (int c = 0; c < RS_ITEMS; ++c) {
        const int64_t i = base + c * 32;
        if (i < n) {
            const int d = (int)((key[c] >> shift) & 0xff);
            const int64_t pos = (int64_t)dbase[d] + cnt[w][d] + rank[c];
            keys_out[pos] = key[c];
            vals_out[pos] = val[c];
        }
    }
}

struct HistIn {
    const int *h;
    __device__ int operator()(int64_t i) const { return h[i]; }
};
struct HistOut {
    int *h;
    __device__ void operator()(int64_t i, int ex, int) const { h[i] = ex; }
};


// scratch sizes (ints) for n items
static inline int64_t hist_ints(int64_t n) { return 256 * kb_cdiv(n > 0 ? n : 1, RS_TILE); }
static inline int64_t scan_ints(int64_t n) { return kbscan::workspace_ints(hist_ints(n)); }

// Sorts by the low `keybits` bits.  keys/vals are double buffers; returns (through *result) the index of the buffer
// holding the sorted data.  hist: hist_ints(n) ints, scanws: scan_ints(n) ints.
static inline int sort_pairs(uint64_t *keys[2], uint32_t *vals[2], int64_t n, int keybits, int *hist, int *scanws,
                             cudaStream_t s, int *result) {
    int cur = 0;
    if (n > 0) {
        const int64_t ntiles = kb_cdiv(n, RS_TILE);
        const int64_t nhist = 256 * ntiles;
        for (int shift = 0; shift < keybits; shift += 8) {
            k_rs_hist<<<(unsigned)ntiles, RS_THREADS, 0, s>>>(keys[cur], n, shift, ntiles, hist);
            KB_LAUNCH_CHECK();
            int rc = kbscan::exclusive_scan(HistIn{hist}, HistOut{hist}, nhist, scanws, s);
            if (rc) return rc;
            k_rs_scatter<<<(unsigned)ntiles, RS_THREADS, 0, s>>>(keys[cur], vals[cur], n, shift, ntiles, hist,
                                                                 keys[cur ^ 1], vals[cur ^ 1]);
            KB_LAUNCH_CHECK();
            cur ^= 1;
        }
    }
    *result = cur;
    return 0;
}

static inline int bits_for(uint64_t maxval) {     // number of bits needed to represent values in [0, maxval]
    int b = 1;
    while (b < 63 && (maxval >> b) != 0) ++b;
    return b;
}

}  // namespace kbradix
