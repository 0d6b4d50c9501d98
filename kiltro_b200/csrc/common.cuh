// Shared device/host helpers for libkiltro_b200 (sm_100a).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include "../../include/kiltro_b200.h"

#define KB_WARP 32
#define KB_FULL 0xffffffffu

extern int64_t g_kb_launches;          // api.cu
#define KB_COUNT_LAUNCH() (++g_kb_launches)

#define KB_CUDA(expr)                                                    \
    do {                                                                 \
        cudaError_t _e = (expr);                                         \
        if (_e != cudaSuccess) return (int)_e;                           \
    } while (0)

#define KB_LAUNCH_CHECK()                                                \
    do {                                                                 \
        KB_COUNT_LAUNCH();                                               \
        cudaError_t _e = cudaGetLastError();                             \
        if (_e != cudaSuccess) return (int)_e;                           \
    } while (0)

static inline int64_t kb_cdiv(int64_t a, int64_t b) { return (a + b - 1) / b; }

// Function attributes (cudaFuncSetAttribute) are per device: `static KbOncePerDevice once; if (once.need()) { ... }`
// runs the block once for every device ordinal a process uses (a process-wide flag would skip the second device).
struct KbOncePerDevice {
    bool done[64] = {};
    bool need() {
        int dev = 0;
        if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return true;
        if (done[dev]) return false;
        done[dev] = true;
        return true;
    }
};

// Programmatic dependent launch (sm_90+): a kernel launched with kb_launch_pdl may become resident while the previous
// kernel of the stream is still draining; everything it does before kb_pdl_wait() (shared-memory set-up, loads of data
// the previous kernel does not write) overlaps that tail and the launch latency.  kb_pdl_wait() returns when the
// previous kernel has completed and its writes are visible; without the launch attribute it is a no-op.  A kernel
// must execute it before its first access to anything a predecessor reads or writes, and before any early return
// (its own successor is released only when every CTA has passed the wait or exited).
#ifdef __CUDACC__
__device__ __forceinline__ void kb_pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }

extern int g_kb_pdl;                   // api.cu: 1 (default) = the Krylov loops launch their kernels with the attribute
template <class... KArgs, class... Args>
static inline cudaError_t kb_launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t s,
                                        bool pdl, Args... args) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = s;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr; cfg.numAttrs = (pdl && g_kb_pdl) ? 1 : 0;
    return cudaLaunchKernelEx(&cfg, kernel, KArgs(args)...);
}
#endif

// grid size for a grid-stride kernel: a multiple of the SM count, capped by the work available
int kb_num_sms();                      // api.cu (cached)
static inline int kb_grid(int64_t work_items, int threads, int ctas_per_sm) {
    int64_t need = kb_cdiv(work_items, threads);
    int64_t cap = (int64_t)kb_num_sms() * ctas_per_sm;
    if (need < 1) need = 1;
    return (int)(need < cap ? need : cap);
}

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(KB_FULL, v, o);
    return v;
}
__device__ __forceinline__ int warp_sum_i(int v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(KB_FULL, v, o);
    return v;
}
// inclusive warp scan
__device__ __forceinline__ int warp_scan_incl(int v) {
    const int lane = threadIdx.x & 31;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        int t = __shfl_up_sync(KB_FULL, v, o);
        if (lane >= o) v += t;
    }
    return v;
}

// Block-wide exclusive scan of one int per thread (blockDim.x <= 1024, multiple of 32).
// `smem` must hold 33 ints.  Returns the exclusive prefix; *total gets the block sum.
__device__ __forceinline__ int block_scan_excl(int v, int *smem, int *total) {
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = blockDim.x >> 5;
    int inc = warp_scan_incl(v);
    if (lane == 31) smem[w] = inc;
    __syncthreads();
    if (w == 0) {
        int s = lane < nw ? smem[lane] : 0;
        int si = warp_scan_incl(s);
        smem[lane] = si - s;
        if (lane == 31) smem[32] = si;
    }
    __syncthreads();
    int r = smem[w] + inc - v;
    *total = smem[32];
    __syncthreads();
    return r;
}

// Block-wide fixed-order sum of one double per thread; result valid in thread 0.  `smem` holds 32 doubles.
__device__ __forceinline__ double block_sum(double v, double *smem) {
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = blockDim.x >> 5;
    v = warp_sum(v);
    if (lane == 0) smem[w] = v;
    __syncthreads();
    double r = 0.0;
    if (w == 0) {
        r = lane < nw ? smem[lane] : 0.0;
        r = warp_sum(r);
    }
    __syncthreads();
    return r;
}

// streaming loads/stores that do not pollute L1 (data touched once)
__device__ __forceinline__ double ld_stream(const double *p) { return __ldcs(p); }
__device__ __forceinline__ int ld_stream(const int *p) { return __ldcs(p); }
__device__ __forceinline__ void st_stream(double *p, double v) { __stcs(p, v); }
