// K5/K6: SpMV entry points and the Jacobi-preconditioned Krylov solvers (CG for symmetric systems, BiCGSTAB otherwise).
//
// Replaces LinearSolver.Solve -> scipy.sparse.linalg.spsolve (SuperLU), kiltro/FEMSolver.py:350-364, with iterative
// solvers that run entirely on device: the host loop below only enqueues kernels; all scalars (rho, alpha, omega, beta)
// live in device memory and are produced by the CTA that finishes a fused dot product last, so there is no host
// round-trip per iteration.  Convergence is tested on device every iteration (a `done` flag freezes the iterate: later
// kernels of the batch exit immediately) and read back by the host every `check_every` iterations.
//
// Jacobi preconditioning is applied once, as a diagonal scaling of a private copy of the values:
//   CG:        Ahat = D^-1/2 A D^-1/2,  bhat = D^-1/2 b,  x = D^-1/2 xhat      (== Jacobi-PCG in exact arithmetic)
//   BiCGSTAB:  Ahat = A D^-1,           xhat = D x                              (== right-Jacobi BiCGSTAB; the residual
//              of the scaled system IS the true residual b - A x)
// which removes two vector reads and one write per iteration compared with applying D^-1 inside the loop.
#include <math.h>
#include <string.h>
#include <cstdio>
#include <cstdlib>
#include "common.cuh"
#include "spmv.cuh"
#include "spmv_tma.cuh"
#include "spmv_pat.cuh"
#include "peer.cuh"

using namespace kbspmv;

namespace {

// ---- optional live timing of the SpMV launches inside the solver loop (bench.py's roofline leg): every `every`-th
// iteration one SpMV launch is bracketed by two events on the launching stream; elapsed times are collected at the
// solver's own synchronisation points.  Batches in which the solver converged are discarded (their later launches exit
// early and would bias the mean downwards).
struct SpmvProfile {
    bool on = false;
    int every = 8;
    static constexpr int CAP = 64;
    cudaEvent_t ev[2 * CAP];
    bool have_events = false;
    int used = 0;
    int64_t samples = 0;
    double total_ms = 0.0;
} g_prof;

void prof_collect(bool keep) {
    if (keep)
        for (int i = 0; i < g_prof.used; ++i) {
            float ms = 0.f;
            if (cudaEventElapsedTime(&ms, g_prof.ev[2 * i], g_prof.ev[2 * i + 1]) == cudaSuccess) {
                g_prof.total_ms += ms;
                g_prof.samples += 1;
            }
        }
    g_prof.used = 0;
}

// device scalar slots
enum { S_RHO = 0, S_ALPHA, S_OMEGA, S_BETA, S_RES2, S_TOL2, S_AUX0, S_AUX1, S_R0N2, S_WHY, S_URES2, S_UB2, S_COUNT = 16 };   // S_R0N2: <r,r> at the (re)start = |rhat|^2
// S_URES2 / S_UB2 (CG only): ||b - A x||^2 and ||b||^2 of the UNSCALED system (the CG recurrence runs on D^-1/2 A D^-1/2)

struct Work {
    double *vals_hat, *scale, *v0, *v1, *v2, *v3, *v4, *v5, *partials, *sc;
    int16_t *cols16;    // 16-bit column offsets (col - row) of the scaled copy, valid when *bandflag == 0
    int *bandflag;      // device: set to 1 by the setup pass when some |col - row| does not fit 16 bits
    bool use16 = false; // host copy of the decision (re-read from bandflag after setup)
    // row-pattern form (spmv_pat.cuh): one byte per row + a dictionary of <= 256 column-offset patterns
    uint8_t *row_pat;   // (n rounded up to 16) + 16 bytes, zero-filled tail
    unsigned long long *pat_keys;   // P_SLOTS hash keys (0 = empty slot)
    pat_off_t *pat_tab; // P_SLOTS x P_MAXLEN offsets; valid when bandflag[1] == 0
    bool usepat = false;
    int32_t *rowblk;
    unsigned int *ticket;
    int *done;          // done[0]: 0 running, 1 converged, 2 breakdown ; done[1]: iteration at which it was set
    int *iter;          // device iteration counter (incremented by the last CTA of the closing kernel)
};

size_t align256(size_t x) { return (x + 255) & ~(size_t)255; }

size_t carve(Work *w, char *base, int64_t n, int64_t nnz) {
    const int64_t nblk = num_blocks(nnz);
    const int64_t nvec_blocks = 4 * (int64_t)kb_num_sms() * 8 + 4 * nblk + 64;
    size_t off = 0;
#define TAKE(field, type, count)                                  \
    do {                                                          \
        if (w) w->field = (type *)(base + off);                   \
        off += align256(sizeof(type) * (size_t)(count));          \
    } while (0)
    TAKE(vals_hat, double, nnz > 0 ? nnz : 1);
    TAKE(scale, double, n + 1);
    TAKE(v0, double, n + 1); TAKE(v1, double, n + 1); TAKE(v2, double, n + 1);
    TAKE(v3, double, n + 1); TAKE(v4, double, n + 1); TAKE(v5, double, n + 1);
    TAKE(cols16, int16_t, nnz > 0 ? nnz + 8 : 8);
    TAKE(bandflag, int, 4);
    TAKE(row_pat, uint8_t, n + 48);
    TAKE(pat_keys, unsigned long long, P_SLOTS);
    TAKE(pat_tab, pat_off_t, P_SLOTS * P_MAXLEN);
    TAKE(partials, double, nvec_blocks);
    TAKE(sc, double, S_COUNT);
    TAKE(rowblk, int32_t, nblk + 1);
    TAKE(ticket, unsigned int, 4);
    TAKE(done, int, 4);
    TAKE(iter, int, 4);
#undef TAKE
    return off;
}

// ------------------------------------------------------------------------------------------------ setup kernels
// scale[i] from the diagonal: MODE 0 (BiCGSTAB) 1/d ; MODE 1 (CG) 1/sqrt(d) ; d == 0 (or d < 0 for CG) -> 1
template <int MODE>
__global__ void __launch_bounds__(256) k_diag_scale(int64_t n, const int32_t *__restrict__ indptr,
                                                    const int32_t *__restrict__ indices,
                                                    const double *__restrict__ vals, double *__restrict__ scale) {
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        double d = 0.0;
        for (int p = indptr[i], pe = indptr[i + 1]; p < pe; ++p)
            if (indices[p] == (int)i) d += vals[p];
        double s;
        if (MODE == 0) s = (d != 0.0 && isfinite(d)) ? 1.0 / d : 1.0;
        else s = (d > 0.0 && isfinite(d)) ? 1.0 / sqrt(d) : 1.0;
        scale[i] = s;
    }
}

// vals_hat[p] = vals[p] * scale[col]  (MODE 0)   or   scale[row] * vals[p] * scale[col]  (MODE 1)
// Coalesced: a CTA takes 256 consecutive rows, the row of every entry is spread through shared memory by one thread per
// row, then consecutive threads walk consecutive entries (a one-thread-per-row walk reads 72-byte chunks at 72-byte lane
// strides: 1.7 ms at 16 M rows).  Blocks with more than SM_CAP entries walk their rows directly.
constexpr int SM_ROWS = 256, SM_CAP = 4096;
template <int MODE>
__global__ void __launch_bounds__(SM_ROWS) k_scale_matrix(int64_t n, const int32_t *__restrict__ indptr,
                                                          const int32_t *__restrict__ indices,
                                                          const double *__restrict__ vals,
                                                          const double *__restrict__ scale, double *__restrict__ out,
                                                          int16_t *__restrict__ cols16 = nullptr,
                                                          int *__restrict__ bandflag = nullptr) {
    __shared__ int s_ip[SM_ROWS + 1];
    __shared__ uint8_t s_row[SM_CAP];                      // row (relative to the block) of every staged entry
    const int tid = threadIdx.x;
    bool wide = false;
    const int64_t nblocks = (n + SM_ROWS - 1) / SM_ROWS;
    for (int64_t blk = blockIdx.x; blk < nblocks; blk += gridDim.x) {
        const int64_t r0 = blk * SM_ROWS;
        const int nr = (int)((n - r0) < SM_ROWS ? (n - r0) : SM_ROWS);
        for (int i = tid; i <= nr; i += SM_ROWS) s_ip[i] = indptr[r0 + i];
        __syncthreads();
        const int p0 = s_ip[0], np = s_ip[nr] - p0;
        if (np <= SM_CAP) {
            if (tid < nr)
                for (int i = s_ip[tid] - p0, ie = s_ip[tid + 1] - p0; i < ie; ++i) s_row[i] = (uint8_t)tid;
            __syncthreads();
            for (int i = tid; i < np; i += SM_ROWS) {
                const int64_t row = r0 + s_row[i];
                const int col = indices[p0 + i];
                const double si = MODE == 1 ? __ldg(scale + row) : 1.0;
                out[p0 + i] = si * vals[p0 + i] * __ldg(scale + col);
                if (cols16 != nullptr) {
                    const int d = col - (int)row;
                    if (d < -32768 || d > 32767) wide = true;
                    cols16[p0 + i] = (int16_t)d;
                }
            }
        } else if (tid < nr) {
            const int64_t i = r0 + tid;
            const double si = MODE == 1 ? scale[i] : 1.0;
            for (int p = s_ip[tid], pe = s_ip[tid + 1]; p < pe; ++p) {
                const int col = indices[p];
                out[p] = si * vals[p] * __ldg(scale + col);
                if (cols16 != nullptr) {
                    const int d = col - (int)i;
                    if (d < -32768 || d > 32767) wide = true;
                    cols16[p] = (int16_t)d;
                }
            }
        }
        __syncthreads();
    }
    if (wide) *bandflag = 1;          // benign race: every writer stores the same value
}

// generic fused vector kernel: body(i) per element, optional 2 dot accumulators finalised by the last CTA
template <int DOTS, class Body, class Final>
__global__ void __launch_bounds__(256) k_vec(int64_t n, Body body, double *__restrict__ partials,
                                             unsigned int *__restrict__ ticket, const int *__restrict__ done,
                                             Final fin) {
    if (done != nullptr && *done != 0) return;
    __shared__ double red[32];
    __shared__ int s_last;
    double d0 = 0.0, d1 = 0.0;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
#pragma unroll 4
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) body(i, d0, d1);
    if (DOTS) {
        d0 = block_sum(d0, red);
        d1 = block_sum(d1, red);
        if (threadIdx.x == 0) {
            partials[blockIdx.x] = d0;
            partials[gridDim.x + blockIdx.x] = d1;
            __threadfence();
            const unsigned int t = atomicInc(ticket, gridDim.x - 1);
            s_last = (t == gridDim.x - 1);
        }
        __syncthreads();
        if (s_last) {
            __threadfence();
            double a0 = 0.0, a1 = 0.0;
            for (int i = threadIdx.x; i < (int)gridDim.x; i += 256) {
                a0 += __ldcg(partials + i);
                a1 += __ldcg(partials + gridDim.x + i);
            }
            a0 = block_sum(a0, red);
            a1 = block_sum(a1, red);
            if (threadIdx.x == 0) fin(a0, a1);
        }
    }
}

struct NoFin { __device__ void operator()(double, double, double = 0.0, double = 0.0) const {} };

template <int DOTS, class Body, class Final>
int launch_vec(int64_t n, Body body, Final fin, const Work &w, const int *done, cudaStream_t s) {
    const int grid = kb_grid(n, 256, 8);
    k_vec<DOTS, Body, Final><<<grid, 256, 0, s>>>(n, body, w.partials, w.ticket, done, fin);
    KB_LAUNCH_CHECK();
    return 0;
}

// Hot vector kernels: the body is split into load / apply so that four grid-stride iterations have all their loads in
// flight before the first store (explicit memory-level parallelism; the compiler cannot hoist loads over the stores of
// the previous iteration because the streams may alias as far as it knows).  Scalars are read once per thread.
template <int DOTS, class Body, class Final>
__global__ void __launch_bounds__(256) k_vec4(int64_t n, Body body, double *__restrict__ partials,
                                              unsigned int *__restrict__ ticket, const int *__restrict__ done,
                                              Final fin) {
    kb_pdl_wait();                                          // scalars, vectors and `done` come from the previous kernel
    if (done != nullptr && *done != 0) return;
    __shared__ double red[32];
    __shared__ int s_last;
    double d0 = 0.0, d1 = 0.0;
    const typename Body::S sc = body.scalars();
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    for (; i + 3 * stride < n; i += 4 * stride) {
        typename Body::L l0, l1, l2, l3;
        body.load(i, l0); body.load(i + stride, l1); body.load(i + 2 * stride, l2); body.load(i + 3 * stride, l3);
        body.apply(i, l0, sc, d0, d1); body.apply(i + stride, l1, sc, d0, d1);
        body.apply(i + 2 * stride, l2, sc, d0, d1); body.apply(i + 3 * stride, l3, sc, d0, d1);
    }
    for (; i < n; i += stride) {
        typename Body::L l0;
        body.load(i, l0);
        body.apply(i, l0, sc, d0, d1);
    }
    if (DOTS) {
        d0 = block_sum(d0, red);
        d1 = block_sum(d1, red);
        if (threadIdx.x == 0) {
            partials[blockIdx.x] = d0;
            partials[gridDim.x + blockIdx.x] = d1;
            __threadfence();
            const unsigned int t = atomicInc(ticket, gridDim.x - 1);
            s_last = (t == gridDim.x - 1);
        }
        __syncthreads();
        if (s_last) {
            __threadfence();
            double a0 = 0.0, a1 = 0.0;
            for (int k = threadIdx.x; k < (int)gridDim.x; k += 256) {
                a0 += __ldcg(partials + k);
                a1 += __ldcg(partials + gridDim.x + k);
            }
            a0 = block_sum(a0, red);
            a1 = block_sum(a1, red);
            if (threadIdx.x == 0) fin(a0, a1);
        }
    }
}

template <int DOTS, class Body, class Final>
int launch_vec4(int64_t n, Body body, Final fin, double *partials, unsigned int *ticket, const int *done, cudaStream_t s,
                bool pdl = false) {
    if (n <= 0) return 0;
    const int grid = kb_grid(kb_cdiv(n, 4), 256, 8);
    KB_CUDA(kb_launch_pdl(k_vec4<DOTS, Body, Final>, dim3(grid), dim3(256), 0, s, pdl, n, body, partials, ticket, done, fin));
    KB_LAUNCH_CHECK();
    return 0;
}

// ---- hot bodies (scalars by device pointer: the single-GPU solver passes addresses inside its scalar block, the
// partitioned solver (K9) its all-reduced values)
struct HBiS {               // r -= alpha v
    const double *alpha, *v; double *r;
    struct S { double a; }; struct L { double r, v; };
    __device__ S scalars() const { return S{alpha[0]}; }
    __device__ void load(int64_t i, L &l) const { l.r = r[i]; l.v = __ldcs(v + i); }
    __device__ void apply(int64_t i, const L &l, const S &s, double &, double &) const { r[i] = l.r - s.a * l.v; }
};
struct HBiUpdate {          // x += alpha p + omega s ; r = s - omega t ; d0 = <rhat,r> ; d1 = <r,r>
    const double *alpha, *omega, *p, *t, *rhat; double *x, *r;
    struct S { double a, om; }; struct L { double x, p, s, t, rh; };
    __device__ S scalars() const { return S{alpha[0], omega[0]}; }
    __device__ void load(int64_t i, L &l) const { l.x = x[i]; l.p = p[i]; l.s = r[i]; l.t = __ldcs(t + i); l.rh = rhat[i]; }
    __device__ void apply(int64_t i, const L &l, const S &s, double &d0, double &d1) const {
        x[i] = l.x + (s.a * l.p + s.om * l.s);
        const double ri = l.s - s.om * l.t;
        r[i] = ri;
        d0 += l.rh * ri; d1 += ri * ri;
    }
};
struct HBiUpdateP {         // x += alpha p + omega s ; r = s - omega t ; p = r + beta (p - omega v) ; d0 = <r,r>
    const double *alpha, *omega, *beta, *t, *v; double *x, *r, *p;
    struct S { double a, om, b; }; struct L { double x, p, s, t, v; };
    __device__ S scalars() const { return S{alpha[0], omega[0], beta[0]}; }
    __device__ void load(int64_t i, L &l) const { l.x = x[i]; l.p = p[i]; l.s = r[i]; l.t = __ldcs(t + i); l.v = __ldcs(v + i); }
    __device__ void apply(int64_t i, const L &l, const S &s, double &d0, double &) const {
        x[i] = l.x + (s.a * l.p + s.om * l.s);
        const double ri = l.s - s.om * l.t;
        r[i] = ri;
        p[i] = ri + s.b * (l.p - s.om * l.v);
        d0 += ri * ri;
    }
};
struct HBiP {               // p = r + beta (p - omega v)
    const double *beta, *omega, *r, *v; double *p;
    struct S { double b, om; }; struct L { double r, p, v; };
    __device__ S scalars() const { return S{beta[0], omega[0]}; }
    __device__ void load(int64_t i, L &l) const { l.r = r[i]; l.p = p[i]; l.v = __ldcs(v + i); }
    __device__ void apply(int64_t i, const L &l, const S &s, double &, double &) const { p[i] = l.r + s.b * (l.p - s.om * l.v); }
};
struct HCgUpdate {          // x += alpha p ; r -= alpha q ; d0 = <r,r>
    const double *alpha, *p, *q; double *x, *r;
    struct S { double a; }; struct L { double x, p, r, q; };
    __device__ S scalars() const { return S{alpha[0]}; }
    __device__ void load(int64_t i, L &l) const { l.x = x[i]; l.p = p[i]; l.r = r[i]; l.q = __ldcs(q + i); }
    __device__ void apply(int64_t i, const L &l, const S &s, double &d0, double &) const {
        x[i] = l.x + s.a * l.p;
        const double ri = l.r - s.a * l.q;
        r[i] = ri;
        d0 += ri * ri;
    }
};
struct HCgP {               // p = r + beta p
    const double *beta, *r; double *p;
    struct S { double b; }; struct L { double r, p; };
    __device__ S scalars() const { return S{beta[0]}; }
    __device__ void load(int64_t i, L &l) const { l.r = r[i]; l.p = p[i]; }
    __device__ void apply(int64_t i, const L &l, const S &s, double &, double &) const { p[i] = l.r + s.b * l.p; }
};

// SpMV launcher: persistent TMA-fed kernel when the CSR arrays are 16-byte aligned, one-shot CSR-stream kernel otherwise
int g_force_simple_spmv = 0;
int g_no_patterns = 0;   // 1 disables the row-pattern form of the SpMV (kb_spmv_force_no_patterns)
int g_force_idx32 = 0;   // 1 keeps 32-bit column indices even when the 16-bit offset form is available (kb_spmv_force_idx32)

// ---- row-pattern dictionary (spmv_pat.cuh): built on device from the CSR pattern, once per solver setup
__device__ __forceinline__ unsigned long long pat_mix(unsigned long long x) {
    x ^= x >> 33; x *= 0xff51afd7ed558ccdull; x ^= x >> 33; x *= 0xc4ceb9fe1a85ec53ull; x ^= x >> 33;
    return x;
}
// Every row hashes its (length, col - row ...) tuple and claims / finds a slot of a 256-entry open-addressing table; the
// slot number is the row's pattern id.  flag[1] is raised when a row is longer than P_MAXLEN or the table overflows.
__global__ void __launch_bounds__(256) k_pattern_insert(int64_t n, const int32_t *__restrict__ indptr,
                                                        const int32_t *__restrict__ indices,
                                                        unsigned long long *__restrict__ keys, pat_off_t *__restrict__ tab,
                                                        uint8_t *__restrict__ row_pat, int *__restrict__ flag) {
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        const int p0 = indptr[i], len = indptr[i + 1] - p0;
        if (len > P_MAXLEN) { flag[1] = 1; row_pat[i] = 0; continue; }
        unsigned long long h = pat_mix(0x9e3779b97f4a7c15ull + (unsigned long long)len);
        bool wide = false;
        for (int k = 0; k < len; ++k) {
            const int d = indices[p0 + k] - (int)i;
            wide |= (d < -32768 || d > 32767);
            h = pat_mix(h ^ (unsigned long long)(unsigned int)d);
        }
        if (wide) { flag[1] = 1; row_pat[i] = 0; continue; }
        h |= 1ull;                                              // 0 is the empty marker
        int slot = (int)(h % P_SLOTS), found = -1;
        for (int probe = 0; probe < P_SLOTS; ++probe) {
            // test before test-and-set: after the first few rows every key is already there, and 16 M rows hammering
            // the same handful of slots with atomics cost 10.8 ms at S16
            unsigned long long prev = *reinterpret_cast<volatile unsigned long long *>(keys + slot);
            if (prev == 0ull) prev = atomicCAS(keys + slot, 0ull, h);
            if (prev == 0ull) {                                 // claimed: this row defines the pattern
                for (int k = 0; k < P_MAXLEN; ++k) tab[slot * P_MAXLEN + k] = (pat_off_t)(k < len ? indices[p0 + k] - (int)i : 0);
                found = slot; break;
            }
            if (prev == h) { found = slot; break; }
            slot = (slot + 1) % P_SLOTS;
        }
        if (found < 0) { flag[1] = 1; found = 0; }
        row_pat[i] = (uint8_t)found;
    }
}
// Second pass (after the table is complete): every row checks its offsets against the slot it was given -- a hash
// collision between different patterns raises flag[1] instead of producing a wrong matrix.
__global__ void __launch_bounds__(256) k_pattern_verify(int64_t n, const int32_t *__restrict__ indptr,
                                                        const int32_t *__restrict__ indices, const pat_off_t *__restrict__ tab,
                                                        const uint8_t *__restrict__ row_pat, int *__restrict__ flag) {
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    bool bad = false;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        const int p0 = indptr[i], len = indptr[i + 1] - p0;
        if (len > P_MAXLEN) { bad = true; continue; }
        const pat_off_t *po = tab + (int)row_pat[i] * P_MAXLEN;
        for (int k = 0; k < len; ++k) bad |= ((int)__ldg(po + k) != indices[p0 + k] - (int)i);
    }
    if (bad) flag[1] = 1;
}

template <int DOTS, class Final>
int launch_spmv_raw(int64_t n, int64_t nnz, const int32_t *indptr, const int32_t *indices, const double *vals,
                    const int32_t *rowblk, double *partials, unsigned int *ticket, const double *x, double *y,
                    const double *dotw, const int *done, Final fin, cudaStream_t s, const int16_t *indices16 = nullptr,
                    const double *dotw2 = nullptr, kbpeer::Wait pwait = kbpeer::Wait{nullptr, nullptr, 0ull, nullptr, 0},
                    const uint8_t *row_pat = nullptr, const pat_off_t *pat_tab = nullptr, bool pdl = false) {
    const int64_t nblk = num_blocks(nnz);
    if (row_pat != nullptr) {                           // row-pattern form: 8 bytes per non-zero cross HBM
        // like the 16-bit form it addresses x relative to the first row of the view: never fall back silently
        if (pat_tab == nullptr || g_force_simple_spmv || !tma_eligible(indptr, row_pat, vals)) return KB_EINVAL;
        static KbOncePerDevice configured;                 // one flag per template instantiation
        constexpr size_t bytes = pat_smem_bytes();
        if (configured.need()) {
            KB_CUDA(cudaFuncSetAttribute(k_spmv_pat<DOTS, Final>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes));
        }
        KB_CUDA(kb_launch_pdl(k_spmv_pat<DOTS, Final>, dim3(tma_grid(nblk)), dim3(P_THREADS), bytes, s, pdl, nblk, indptr,
                              row_pat, pat_tab, vals, rowblk, x, y, dotw, partials, ticket, done, fin, dotw2, pwait));
        KB_LAUNCH_CHECK();
        return 0;
    }
    // 16-bit offsets change the meaning of x (relative to the first row of the view): never fall back silently
    if (indices16 != nullptr && (g_force_simple_spmv || !tma_eligible(indptr, indices16, vals))) return KB_EINVAL;
    if (indices16 != nullptr) {
        static KbOncePerDevice configured16;
        constexpr size_t bytes = tma_smem_bytes<short>();
        if (configured16.need()) {
            KB_CUDA(cudaFuncSetAttribute(k_spmv_tma<DOTS, Final, short>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes));
        }
        k_spmv_tma<DOTS, Final, short><<<tma_grid(nblk), T_THREADS, bytes, s>>>(nblk, indptr, (const short *)indices16, vals,
                                                                               rowblk, x, y, dotw, partials, ticket, done, fin,
                                                                               dotw2, pwait);
    } else if (!g_force_simple_spmv && tma_eligible(indptr, indices, vals)) {
        static KbOncePerDevice configured;          // one flag per template instantiation
        constexpr size_t bytes = tma_smem_bytes<int>();
        if (configured.need()) {
            KB_CUDA(cudaFuncSetAttribute(k_spmv_tma<DOTS, Final, int>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes));
        }
        k_spmv_tma<DOTS, Final, int><<<tma_grid(nblk), T_THREADS, bytes, s>>>(nblk, indptr, indices, vals, rowblk, x, y, dotw,
                                                                             partials, ticket, done, fin, dotw2, pwait);
    } else {
        if (pwait.f0 != nullptr || pwait.f1 != nullptr) return KB_EINVAL;     // the peer path needs 16-byte aligned CSR arrays
        k_spmv<DOTS, Final><<<(unsigned)nblk, THREADS, 0, s>>>(n, indptr, indices, vals, rowblk, x, y, dotw, partials,
                                                              ticket, done, fin, dotw2);
    }
    KB_LAUNCH_CHECK();
    return 0;
}

template <int DOTS, class Final>
int launch_spmv(int64_t n, int64_t nnz, const int32_t *indptr, const int32_t *indices, const double *vals,
                const Work &w, const double *x, double *y, const double *dotw, const int *done, Final fin,
                cudaStream_t s, const double *dotw2 = nullptr, bool pdl = false) {
    return launch_spmv_raw<DOTS, Final>(n, nnz, indptr, indices, vals, w.rowblk, w.partials, w.ticket, x, y, dotw, done,
                                        fin, s, w.use16 ? w.cols16 : nullptr, dotw2,
                                        kbpeer::Wait{nullptr, nullptr, 0ull, nullptr, 0}, w.usepat ? w.row_pat : nullptr, w.pat_tab,
                                        pdl);
}

// ------------------------------------------------------------------------------------------------ functors
struct BodyScaleIn {        // xh = x / scale ; bh = b * bscale   (bscale may be null -> copy)
    const double *x, *b, *scale; double *xh, *bh; int scale_b;
    __device__ void operator()(int64_t i, double &, double &) const {
        const double s = scale[i];
        xh[i] = x[i] / s;
        bh[i] = scale_b ? b[i] * s : b[i];
    }
};
struct BodyResidual {       // r = b - q ; p = r ; rhat = r ; d0 = <r,r> ; d1 = <b,b>
    const double *b, *q; double *r, *p, *rhat;
    __device__ void operator()(int64_t i, double &d0, double &d1) const {
        const double bi = b[i], ri = bi - q[i];
        r[i] = ri; p[i] = ri;
        if (rhat) rhat[i] = ri;
        d0 += ri * ri; d1 += bi * bi;
    }
};
struct FinInit {            // rho = <r,r> ; tol2 = max(rtol^2 <b,b>, tiny) ; done if already converged
    double *sc; int *done; double rtol2;
    __device__ void operator()(double rr, double bb, double = 0.0, double = 0.0) const {
        sc[S_RHO] = rr; sc[S_RES2] = rr; sc[S_AUX0] = bb; sc[S_R0N2] = rr;
        const double tol2 = rtol2 * bb;
        sc[S_TOL2] = tol2;
        sc[S_ALPHA] = 0.0; sc[S_OMEGA] = 1.0; sc[S_BETA] = 0.0;
        if (!(rr > tol2)) { done[0] = isfinite(rr) ? 1 : 2; }
    }
};
// CG iterates on the symmetrically scaled system; the verdict (and kb_solve_info) use the residual of the system the
// caller handed over: r_unscaled = r / scale, b_unscaled = bh / scale  (scale = 1/sqrt(diag)).
struct BodyUnscaledNorms {  // d0 = ||r / scale||^2 ; d1 = ||bh / scale||^2
    const double *r, *bh, *scale;
    __device__ void operator()(int64_t i, double &d0, double &d1) const {
        const double s = scale[i], ru = r[i] / s, bu = bh[i] / s;
        d0 += ru * ru; d1 += bu * bu;
    }
};
struct FinInitUnscaled {    // runs right after FinInit: the unscaled residual decides; the recurrence's (scaled) target follows it
    double *sc; int *done; double rtol2, slack;
    __device__ void operator()(double ru2, double bu2, double = 0.0, double = 0.0) const {
        sc[S_URES2] = ru2; sc[S_UB2] = bu2;
        const double tolu2 = rtol2 * bu2;
        if (!isfinite(ru2)) { done[0] = 2; return; }
        if (!(ru2 > tolu2)) { done[0] = 1; return; }
        if (done[0] == 1) done[0] = 0;                         // converged in the scaled norm only: keep iterating
        // scaled residual the recurrence has to reach for the unscaled one to meet rtol (ratio of the two norms taken
        // from the current residual; `slack` < 1 after a cycle that ended short of it)
        const double t = slack * sc[S_RES2] * (tolu2 / ru2);
        if (t < sc[S_TOL2]) sc[S_TOL2] = t;
    }
};
struct BodyUnscale {        // x = xh * scale
    const double *xh, *scale; double *x;
    __device__ void operator()(int64_t i, double &, double &) const { x[i] = xh[i] * scale[i]; }
};

// ---- CG
struct FinCgAlpha {         // after q = A p, d0 = <p,q>: alpha = rho / <p,q>
    double *sc; int *done;
    __device__ void operator()(double pq, double, double = 0.0, double = 0.0) const {
        const double a = sc[S_RHO] / pq;
        sc[S_ALPHA] = a;
        if (!isfinite(a) || pq == 0.0) done[0] = 2;
    }
};
struct FinCgBeta {          // beta = rho_new / rho ; rho = rho_new ; convergence test ; iteration count
    double *sc; int *done; int *iter;
    __device__ void operator()(double rr, double, double = 0.0, double = 0.0) const {
        sc[S_BETA] = rr / sc[S_RHO];
        sc[S_RHO] = rr; sc[S_RES2] = rr;
        const int it = ++iter[0];
        if (!isfinite(rr)) { done[0] = 2; done[1] = it; }
        else if (!(rr > sc[S_TOL2])) { done[0] = 1; done[1] = it; }
    }
};

// ---- BiCGSTAB
// Breakdown guards (done = 2 -> the host restarts from the true residual of the current iterate, rhat = r):
//   * exact breakdown: <rhat,v> = 0, rho = 0, omega = 0 or a non-finite scalar;
//   * spike: the step alpha v about to be applied is 1e5 times longer than the residual it is meant to reduce.  In
//     fp64 BiCGSTAB <rhat,v> and rho sink to rounding noise on transport-like problems (the residual front and the
//     shadow Krylov vectors travel in opposite directions: seen on the uniform-velocity closed-form case at 16 M DOF);
//     the iteration usually survives on noise, but one near-zero <rhat,v> produces an alpha of size 1/eps that wrecks
//     the iterate beyond recovery.  The flag is raised BEFORE the update (the following kernels see done != 0), so the
//     iterate the restart begins from is still good;
//   * divergence: the recurrence residual has grown 1e5-fold over the residual the cycle started from.
// (A scale-free test on cos(rhat, r) does not work: rho legitimately decays relative to |rhat||r|.)
constexpr double KB_GROWTH2_MAX = 1.0e10;
struct FinBiAlpha {         // after v = A p, d = {<rhat,v>, <v,v>}: alpha = rho / <rhat,v>
    double *sc; int *done;
    __device__ void operator()(double rv, double vv, double = 0.0, double = 0.0) const {
        const double a = sc[S_RHO] / rv;
        sc[S_ALPHA] = a;
        if (!isfinite(a) || rv == 0.0) { done[0] = 2; sc[S_WHY] = 1.0; }
        else if (a * a * vv > KB_GROWTH2_MAX * sc[S_RES2]) { done[0] = 2; sc[S_WHY] = 2.0; }
    }
};
// Merged update (one vector kernel instead of two): with <rhat,t> and <rhat,s> taken in the epilogue of t = A s,
// rho_new = <rhat, s - omega t> = <rhat,s> - omega <rhat,t> is known BEFORE x and r are updated, so beta is too and
// p = r + beta (p - omega v) rides along with the x/r update (64 n bytes instead of 56 n + 32 n).
struct FinBiOmega2 {        // d = {<s,t>, <t,t>, <rhat,t>, <rhat,s>}
    double *sc; int *done; int *iter;
    __device__ void operator()(double ts, double tt, double rt, double rs) const {
        const double om = tt > 0.0 ? ts / tt : 0.0;
        const double rho_new = rs - om * rt;
        const double beta = (rho_new / sc[S_RHO]) * (sc[S_ALPHA] / om);
        sc[S_OMEGA] = om; sc[S_BETA] = beta; sc[S_RHO] = rho_new;
        sc[S_AUX1] = (isfinite(beta) && rho_new != 0.0 && om != 0.0) ? 0.0 : 1.0;     // breakdown unless converged below
        if (sc[S_AUX1] != 0.0) sc[S_WHY] = 3.0;
    }
};
struct FinBiRes {           // after the merged update: d0 = <r,r>
    double *sc; int *done; int *iter;
    __device__ void operator()(double rr, double, double = 0.0, double = 0.0) const {
        sc[S_RES2] = rr;
        const int it = ++iter[0];
        if (!(rr > sc[S_TOL2]) && isfinite(rr)) { done[0] = 1; done[1] = it; }
        else if (!isfinite(rr) || sc[S_AUX1] != 0.0 || rr > KB_GROWTH2_MAX * sc[S_R0N2]) {
            done[0] = 2; done[1] = it;
            if (sc[S_AUX1] == 0.0) sc[S_WHY] = isfinite(rr) ? 5.0 : 6.0;
        }
    }
};

__global__ void k_reset_state(unsigned int *ticket, int *done, int *iter, double *sc) {
    if (threadIdx.x < 4) { ticket[threadIdx.x] = 0; done[threadIdx.x] = 0; iter[threadIdx.x] = 0; }
    if (threadIdx.x < S_COUNT) sc[threadIdx.x] = 0.0;
}
__global__ void k_clear_done(int *done) { done[0] = 0; done[1] = 0; }

struct HostState { double sc[S_COUNT]; int done[4]; int iter[4]; };

int read_state(const Work &w, HostState *h, cudaStream_t s) {
    KB_CUDA(cudaMemcpyAsync(h->sc, w.sc, sizeof(double) * S_COUNT, cudaMemcpyDeviceToHost, s));
    KB_CUDA(cudaMemcpyAsync(h->done, w.done, sizeof(int) * 4, cudaMemcpyDeviceToHost, s));
    KB_CUDA(cudaMemcpyAsync(h->iter, w.iter, sizeof(int) * 4, cudaMemcpyDeviceToHost, s));
    KB_CUDA(cudaStreamSynchronize(s));
    return 0;
}

// common driver: METHOD 0 = CG, 1 = BiCGSTAB
template <int METHOD>
int krylov_solve(int64_t n, int64_t nnz, const int32_t *indptr, const int32_t *indices, const double *vals,
                 const double *b, double *x, double rtol, int maxiter, int check_every, void *workspace,
                 size_t workspace_bytes, kb_solve_info *info, cudaStream_t s, bool reuse_setup = false) {
    if (!indptr || !indices || (!vals && nnz > 0) || !b || !x || !workspace || !info || n < 0 || nnz < 0) return KB_EINVAL;
    if (n >= INT32_MAX || nnz >= INT32_MAX) return KB_ETOOBIG;
    if (workspace_bytes < carve(nullptr, nullptr, n, nnz)) return KB_EWORKSPACE;
    memset(info, 0, sizeof(*info));
    if (n == 0) return 0;
    if (check_every < 1) check_every = 1;
    if (maxiter < 1) maxiter = 1;
    const int64_t launches0 = g_kb_launches;
    Work w;
    carve(&w, (char *)workspace, n, nnz);
    double *xh = w.v0, *r = w.v1, *p = w.v2, *q = w.v3, *bh = w.v4, *rhat = w.v5;   // BiCGSTAB: q = v, bh reused as t

    k_reset_state<<<1, 32, 0, s>>>(w.ticket, w.done, w.iter, w.sc);
    KB_LAUNCH_CHECK();
    if (!reuse_setup) {      // row partition + Jacobi-scaled copy of the values (kept in the workspace between calls)
        const int64_t nblk = num_blocks(nnz);
        k_partition<<<(unsigned)kb_cdiv(nblk + 1, 256), 256, 0, s>>>(n, nblk, indptr, w.rowblk);
        KB_LAUNCH_CHECK();
        const int g = kb_grid(n, 256, 8);
        k_diag_scale<METHOD == 0 ? 1 : 0><<<g, 256, 0, s>>>(n, indptr, indices, vals, w.scale);
        KB_LAUNCH_CHECK();
        KB_CUDA(cudaMemsetAsync(w.bandflag, 0, sizeof(int) * 4, s));
        k_scale_matrix<METHOD == 0 ? 1 : 0><<<kb_grid(n, SM_ROWS, 8), SM_ROWS, 0, s>>>(n, indptr, indices, vals, w.scale, w.vals_hat,
                                                             g_force_idx32 ? nullptr : w.cols16, w.bandflag);
        KB_LAUNCH_CHECK();
        if (!g_force_idx32 && !g_no_patterns) {                // row-pattern dictionary (bandflag[1] = not expressible)
            KB_CUDA(cudaMemsetAsync(w.pat_keys, 0, sizeof(unsigned long long) * P_SLOTS, s));
            KB_CUDA(cudaMemsetAsync(w.row_pat, 0, (size_t)n + 48, s));
            k_pattern_insert<<<g, 256, 0, s>>>(n, indptr, indices, w.pat_keys, w.pat_tab, w.row_pat, w.bandflag);
            KB_LAUNCH_CHECK();
            k_pattern_verify<<<g, 256, 0, s>>>(n, indptr, indices, w.pat_tab, w.row_pat, w.bandflag);
            KB_LAUNCH_CHECK();
        }
    }
    if (!g_force_idx32) {   // index form of the SpMV, decided once per setup (one small readback): row patterns when the
        int flag[2] = {1, 1};   // matrix has <= 256 of them, else 16-bit column offsets when the band fits, else 32-bit
        KB_CUDA(cudaMemcpyAsync(flag, w.bandflag, sizeof(int) * 2, cudaMemcpyDeviceToHost, s));
        KB_CUDA(cudaStreamSynchronize(s));
        w.use16 = (flag[0] == 0) && !g_force_simple_spmv && tma_eligible(indptr, w.cols16, w.vals_hat);
        w.usepat = (flag[1] == 0) && !g_no_patterns && !g_force_simple_spmv && tma_eligible(indptr, w.row_pat, w.vals_hat);
    }
    int rc;
    // xh = x / scale ; bh = scaled rhs
    if ((rc = launch_vec<0>(n, BodyScaleIn{x, b, w.scale, xh, bh, METHOD == 0 ? 1 : 0}, NoFin{}, w, nullptr, s))) return rc;

    HostState hs;
    double *t = METHOD == 1 ? x : nullptr;   // BiCGSTAB's extra vector borrows the caller's x (rewritten from xh at the end)
    int total_iters = 0, status = KB_ENOTCONV, restarts = 0;
    const double rtol2 = rtol * rtol;
    double prev_true = INFINITY;
    bool recurrence_converged = false, attainable = false;
    // residual / rhs norms (squared) the verdict is taken on: the caller's system.  BiCGSTAB iterates on A D^-1, whose
    // residual IS b - A x; CG iterates on D^-1/2 A D^-1/2 and measures the unscaled residual separately.
    auto res2_of = [&](const HostState &h) { return METHOD == 0 ? h.sc[S_URES2] : h.sc[S_RES2]; };
    auto b2_of = [&](const HostState &h) { return METHOD == 0 ? h.sc[S_UB2] : h.sc[S_AUX0]; };
    for (int restart = 0; restart < 16 && total_iters < maxiter; ++restart) {
        restarts = restart;
        // (re)start from the TRUE residual: r = bh - Ahat xh ; p = r ; rhat = r ; rho = <r,r>
        if ((rc = launch_spmv<0>(n, nnz, indptr, indices, w.vals_hat, w, xh, q, nullptr, nullptr, NoFinal{}, s))) return rc;
        if ((rc = launch_vec<1>(n, BodyResidual{bh, q, r, p, METHOD == 1 ? rhat : nullptr},
                                FinInit{w.sc, w.done, rtol2}, w, nullptr, s))) return rc;
        if (METHOD == 0)
            if ((rc = launch_vec<1>(n, BodyUnscaledNorms{r, bh, w.scale},
                                    FinInitUnscaled{w.sc, w.done, rtol2, restart == 0 ? 1.0 : 0.25}, w, nullptr, s))) return rc;
        if ((rc = read_state(w, &hs, s))) return rc;
        const double true_rr = res2_of(hs), tol2 = rtol2 * b2_of(hs);
        if (hs.done[0] == 1) { status = 0; break; }                       // true residual meets the tolerance
        if (hs.done[0] == 2) { status = KB_EBREAKDOWN; break; }           // non-finite residual
        // attainable accuracy: a whole restart cycle (converged recurrence or breakdown at the rounding floor) did not
        // bring the TRUE residual down 2x in norm while it already sits within 10x of the tolerance -- fp64 rounding
        // floors ||b - A x||; reported through info->attainable_accuracy, never silently
        if (restart > 0 && !(true_rr < 0.25 * prev_true) && true_rr <= 100.0 * tol2) { attainable = true; break; }
        // a recurrence that converged while the true residual stalls far above the tolerance: stop, report failure
        if (recurrence_converged && !(true_rr < 0.25 * prev_true) && true_rr <= 1.0e6 * tol2) { attainable = true; break; }
        prev_true = true_rr;
        recurrence_converged = false;
        bool stop = false;
        while (!stop && total_iters < maxiter) {
            const int batch = (maxiter - total_iters) < check_every ? (maxiter - total_iters) : check_every;
            for (int it = 0; it < batch; ++it) {
                const bool sample = g_prof.on && (it % g_prof.every == 0) && g_prof.used < SpmvProfile::CAP;
                if (sample) cudaEventRecord(g_prof.ev[2 * g_prof.used], s);
                // every kernel of the loop is a programmatic dependent launch (kb_launch_pdl): its launch and set-up
                // overlap the tail of the kernel before it
                if (METHOD == 0) {
                    if ((rc = launch_spmv<1>(n, nnz, indptr, indices, w.vals_hat, w, p, q, p, w.done, FinCgAlpha{w.sc, w.done}, s, nullptr, true))) return rc;
                    if (sample) cudaEventRecord(g_prof.ev[2 * g_prof.used++ + 1], s);
                    if ((rc = launch_vec4<1>(n, HCgUpdate{w.sc + S_ALPHA, p, q, xh, r}, FinCgBeta{w.sc, w.done, w.iter}, w.partials, w.ticket, w.done, s, true))) return rc;
                    if ((rc = launch_vec4<0>(n, HCgP{w.sc + S_BETA, r, p}, NoFin{}, w.partials, w.ticket, w.done, s, true))) return rc;
                } else {
                    if ((rc = launch_spmv<1>(n, nnz, indptr, indices, w.vals_hat, w, p, q, rhat, w.done, FinBiAlpha{w.sc, w.done}, s, nullptr, true))) return rc;
                    if (sample) cudaEventRecord(g_prof.ev[2 * g_prof.used++ + 1], s);
                    if ((rc = launch_vec4<0>(n, HBiS{w.sc + S_ALPHA, q, r}, NoFin{}, w.partials, w.ticket, w.done, s, true))) return rc;
                    if ((rc = launch_spmv<2>(n, nnz, indptr, indices, w.vals_hat, w, r, t, r, w.done, FinBiOmega2{w.sc, w.done, w.iter}, s, rhat, true))) return rc;
                    if ((rc = launch_vec4<1>(n, HBiUpdateP{w.sc + S_ALPHA, w.sc + S_OMEGA, w.sc + S_BETA, t, q, xh, r, p}, FinBiRes{w.sc, w.done, w.iter}, w.partials, w.ticket, w.done, s, true))) return rc;
                }
            }
            if ((rc = read_state(w, &hs, s))) return rc;
            total_iters = hs.iter[0];
            if (g_prof.on) prof_collect(hs.done[0] == 0);
            if (hs.done[0] != 0) stop = true;
        }
        if (getenv("KB_DEBUG_SOLVER"))
            fprintf(stderr, "[kb solver] cycle %d: iters %d done %d why %.0f rho %.3e alpha %.3e omega %.3e beta %.3e rr %.3e r0n2 %.3e tol2 %.3e\n",
                    restart, total_iters, hs.done[0], hs.sc[S_WHY], hs.sc[S_RHO], hs.sc[S_ALPHA], hs.sc[S_OMEGA], hs.sc[S_BETA],
                    hs.sc[S_RES2], hs.sc[S_R0N2], hs.sc[S_TOL2]);
        if (hs.done[0] == 1) recurrence_converged = true;                 // verify against the true residual above
        else if (hs.done[0] == 2) status = KB_EBREAKDOWN;                 // restart from the current iterate
        else break;                                                       // maxiter exhausted
        k_clear_done<<<1, 1, 0, s>>>(w.done);
        KB_LAUNCH_CHECK();
    }
    // true residual with the final iterate (for BiCGSTAB the scaled system's residual is b - A x itself; CG adds the
    // unscaled norms)
    if ((rc = launch_spmv<0>(n, nnz, indptr, indices, w.vals_hat, w, xh, q, nullptr, nullptr, NoFinal{}, s))) return rc;
    if ((rc = launch_vec<1>(n, BodyResidual{bh, q, r, p, nullptr}, FinInit{w.sc, w.iter + 2, rtol2}, w, nullptr, s))) return rc;
    if (METHOD == 0)
        if ((rc = launch_vec<1>(n, BodyUnscaledNorms{r, bh, w.scale}, FinInitUnscaled{w.sc, w.iter + 2, rtol2, 1.0}, w, nullptr, s))) return rc;
    if ((rc = launch_vec<0>(n, BodyUnscale{xh, w.scale, x}, NoFin{}, w, nullptr, s))) return rc;
    if ((rc = read_state(w, &hs, s))) return rc;
    const double res2 = res2_of(hs), b2 = b2_of(hs);
    info->iterations = total_iters;
    info->residual = sqrt(res2);
    info->rhs_norm = sqrt(b2);
    info->restarts = restarts;
    // Verdict, on the TRUE residual of the caller's system only:
    //   0              ||b - A x|| <= rtol ||b||, or <= 10 rtol ||b|| with attainable_accuracy = 1 (fp64 floor, see above)
    //   KB_ESTAGNATED  the iteration stalled above that
    //   KB_EBREAKDOWN  non-finite residual / breakdown that the restarts did not cure
    //   KB_ENOTCONV    maxiter (or the restart budget) exhausted
    if (b2 == 0.0 && res2 == 0.0) status = 0;                                              // b = 0 -> x = 0
    else if (!isfinite(res2)) status = KB_EBREAKDOWN;
    else if (res2 <= rtol2 * b2) status = 0;
    else if (attainable) {
        info->attainable_accuracy = 1;
        status = (res2 <= 100.0 * rtol2 * b2) ? 0 : KB_ESTAGNATED;
    } else if (status == 0) status = KB_ENOTCONV;
    info->status = status;
    info->kernel_launches = g_kb_launches - launches0;
    return 0;
}

}  // namespace

// ------------------------------------------------------------------------------------------------ K7 transient loop
namespace {
struct BodyRhs {            // rhs = (Kff T) + Fm   (both already zero at fixed DOFs)
    const double *kt, *fm; double *rhs;
    __device__ void operator()(int64_t i, double &, double &) const { rhs[i] = kt[i] + fm[i]; }
};
struct BodyStep {           // T = fixed ? td : T - dt*y
    const int32_t *renum; const double *td, *y; double *T; double dt;
    __device__ void operator()(int64_t i, double &, double &) const {
        const int r = renum[i];
        T[i] = r >= 0 ? T[i] - dt * y[i] : td[-1 - r];
    }
};
}  // namespace

// T = fixed ? applied : T - dt*y   (one theta-step update; also used by the partitioned transient loop)
extern "C" int kb_transient_update(int64_t ndof, const int32_t *d_renum, const double *d_applied, const double *d_y,
                                   double dt, double *d_T, void *stream) {
    if (!d_renum || !d_y || !d_T || ndof < 0) return KB_EINVAL;      // d_applied may be NULL when nothing is prescribed
    if (ndof == 0) return 0;
    const int grid = kb_grid(ndof, 256, 8);
    k_vec<0, BodyStep, NoFin><<<grid, 256, 0, (cudaStream_t)stream>>>(ndof, BodyStep{d_renum, d_applied, d_y, d_T, dt}, nullptr,
                                                                    nullptr, nullptr, NoFin{});
    KB_LAUNCH_CHECK();
    return 0;
}

extern "C" size_t kb_transient_workspace_bytes(int64_t ndof, int64_t nnz) {
    return carve(nullptr, nullptr, ndof, nnz) + 4 * align256(sizeof(double) * (size_t)(ndof + 1)) +
           align256(sizeof(int32_t) * (size_t)(num_blocks(nnz) + 1)) + align256(2 * sizeof(double) * 4096);
}

extern "C" int kb_transient_run(int64_t ndof, int64_t nnz, const int32_t *d_indptr, const int32_t *d_indices,
                                const double *d_Kff, const double *d_A, int a_symmetric, const double *d_Fm,
                                const int32_t *d_renum, const double *d_applied, double *d_T, double dt, int nsteps,
                                const int32_t *h_snap_steps, int nsnap, double *d_snapshots, double rtol, int maxiter,
                                void *d_workspace, size_t workspace_bytes, kb_solve_info *h_info, void *stream) {
    // d_applied may be NULL when no DOF is prescribed (it is only dereferenced at fixed DOFs)
    if (!d_indptr || !d_indices || !d_Kff || !d_A || !d_Fm || !d_renum || !d_T || !d_workspace || !h_info ||
        ndof < 1 || nsteps < 0 || nsnap < 0 || (nsnap > 0 && (!h_snap_steps || !d_snapshots)))
        return KB_EINVAL;
    if (workspace_bytes < kb_transient_workspace_bytes(ndof, nnz)) return KB_EWORKSPACE;
    cudaStream_t s = (cudaStream_t)stream;
    const int64_t launches0 = g_kb_launches;
    char *base = (char *)d_workspace;
    const size_t kws = carve(nullptr, nullptr, ndof, nnz);
    size_t off = kws;
    const size_t vb = align256(sizeof(double) * (size_t)(ndof + 1));
    double *kt = (double *)(base + off); off += vb;
    double *rhs = (double *)(base + off); off += vb;
    double *y = (double *)(base + off); off += vb;
    off += vb;
    int32_t *rowblk = (int32_t *)(base + off); off += align256(sizeof(int32_t) * (size_t)(num_blocks(nnz) + 1));
    double *partials = (double *)(base + off);
    Work wv;                               // only partials/ticket are used by launch_vec<0> (no dots): give valid pointers
    carve(&wv, base, ndof, nnz);
    const int64_t nblk = num_blocks(nnz);
    k_partition<<<(unsigned)kb_cdiv(nblk + 1, 256), 256, 0, s>>>(ndof, nblk, d_indptr, rowblk);
    KB_LAUNCH_CHECK();
    KB_CUDA(cudaMemsetAsync(y, 0, sizeof(double) * (size_t)ndof, s));
    int snap = 0, rc;
    int64_t total_iters = 0;
    double worst = 0.0;
    int status = 0, failed_step = 0, steps_done = 0;
    // snapshot of the initial state (step index 0)
    while (snap < nsnap && h_snap_steps[snap] == 0) {
        KB_CUDA(cudaMemcpyAsync(d_snapshots + (size_t)snap * ndof, d_T, sizeof(double) * (size_t)ndof,
                                cudaMemcpyDeviceToDevice, s));
        ++snap;
    }
    for (int step = 1; step <= nsteps; ++step) {
        // rhs = Kff T^n + Fm
        if ((rc = launch_spmv_raw<0, NoFinal>(ndof, nnz, d_indptr, d_indices, d_Kff, rowblk, partials, nullptr, d_T, kt,
                                              nullptr, nullptr, NoFinal{}, s))) return rc;
        if ((rc = launch_vec<0>(ndof, BodyRhs{kt, d_Fm, rhs}, NoFin{}, wv, nullptr, s))) return rc;
        // A y = rhs  (A = M + theta dt Kff; y of the previous step is the initial guess)
        kb_solve_info info;
        if (a_symmetric)
            rc = krylov_solve<0>(ndof, nnz, d_indptr, d_indices, d_A, rhs, y, rtol, maxiter, 8, d_workspace, kws, &info, s,
                                 step > 1);
        else
            rc = krylov_solve<1>(ndof, nnz, d_indptr, d_indices, d_A, rhs, y, rtol, maxiter, 8, d_workspace, kws, &info, s,
                                 step > 1);
        if (rc) return rc;
        total_iters += info.iterations;
        if (info.rhs_norm > 0 && info.residual / info.rhs_norm > worst) worst = info.residual / info.rhs_norm;
        if (info.status != 0 || !isfinite(info.residual)) {
            // a failed inner solve stops the time loop: T keeps the last good state (step - 1), later snapshots are
            // not written, and the caller learns which step failed
            status = info.status != 0 ? info.status : KB_EBREAKDOWN;
            failed_step = step;
            if (!isfinite(info.residual)) worst = NAN;
            break;
        }
        steps_done = step;
        // T^{n+1}
        if ((rc = launch_vec<0>(ndof, BodyStep{d_renum, d_applied, y, d_T, dt}, NoFin{}, wv, nullptr, s))) return rc;
        while (snap < nsnap && h_snap_steps[snap] == step) {
            KB_CUDA(cudaMemcpyAsync(d_snapshots + (size_t)snap * ndof, d_T, sizeof(double) * (size_t)ndof,
                                    cudaMemcpyDeviceToDevice, s));
            ++snap;
        }
    }
    KB_CUDA(cudaStreamSynchronize(s));
    memset(h_info, 0, sizeof(*h_info));
    h_info->iterations = (int32_t)(total_iters > INT32_MAX ? INT32_MAX : total_iters);
    h_info->status = status;
    h_info->residual = worst;           // worst relative residual of the per-step solves
    h_info->rhs_norm = 1.0;
    h_info->failed_step = failed_step;
    h_info->steps_done = steps_done;
    h_info->kernel_launches = g_kb_launches - launches0;
    return 0;
}

// ------------------------------------------------------------------------------------------------ building blocks
// The same fused kernels as above, exposed one by one with their scalars passed as DEVICE pointers, for solver loops
// driven from the host side when a collective sits between a dot product and its use (row-block partitioned multi-GPU
// solve, kiltro_b200/distributed.py): local partial dots come back in d_out2 (2 doubles) and are all-reduced by the caller.
namespace {
struct FinStore { double *out; __device__ void operator()(double a, double b, double = 0.0, double = 0.0) const { out[0] = a; out[1] = b; } };
struct BodyDot2 { const double *a, *b; __device__ void operator()(int64_t i, double &d0, double &d1) const { const double bi = b[i]; d0 += a[i] * bi; d1 += bi * bi; } };
struct BodyMulP {           // out = a .* b
    const double *a, *b; double *out;
    __device__ void operator()(int64_t i, double &, double &) const { out[i] = a[i] * b[i]; }
};
struct BodyResidualP {      // r = b - q ; d0 = <r,r> ; d1 = <b,b>
    const double *b, *q; double *r;
    __device__ void operator()(int64_t i, double &d0, double &d1) const {
        const double bi = b[i], ri = bi - q[i];
        r[i] = ri; d0 += ri * ri; d1 += bi * bi;
    }
};

// scratch for the partial sums of the building blocks: [2*(max grid)] doubles + ticket, carved from a caller buffer
struct Scratch { double *partials; unsigned int *ticket; };
Scratch scratch_of(void *d_scratch) {
    Scratch sc;
    sc.ticket = (unsigned int *)d_scratch;
    sc.partials = (double *)((char *)d_scratch + 256);
    return sc;
}
template <int DOTS, class Body, class Final>
int launch_vec_s(int64_t n, Body body, Final fin, const Scratch &sc, cudaStream_t s) {
    if (n <= 0) return 0;
    const int grid = kb_grid(n, 256, 8);
    k_vec<DOTS, Body, Final><<<grid, 256, 0, s>>>(n, body, sc.partials, sc.ticket, nullptr, fin);
    KB_LAUNCH_CHECK();
    return 0;
}
}  // namespace

// bytes of device scratch the building blocks need for a system with nnz non-zeros (zero-initialise it once)
extern "C" size_t kb_blocks_scratch_bytes(int64_t nnz) {
    return 256 + sizeof(double) * (size_t)(4 * num_blocks(nnz) + 4 * 8 * (int64_t)kb_num_sms() + 64);
}

// y = A x over `nrows` rows + local partial dots d_out2 = {<w,y>, <y,y>}
extern "C" int kb_spmv_dots(int64_t nrows, int64_t nnz, const int32_t *d_indptr, const int32_t *d_indices,
                            const int16_t *d_cols16, const uint8_t *d_row_pat, const int16_t *d_pat_tab,
                            const double *d_vals, const int32_t *d_rowblk, const double *d_x,
                            double *d_y, const double *d_w, double *d_out2, void *d_scratch, void *stream) {
    if (!d_indptr || !d_indices || (!d_vals && nnz > 0) || !d_rowblk || !d_x || !d_y || !d_w || !d_out2 || !d_scratch)
        return KB_EINVAL;
    if (nrows <= 0) return 0;
    const Scratch sc = scratch_of(d_scratch);
    return launch_spmv_raw<1, FinStore>(nrows, nnz, d_indptr, d_indices, d_vals, d_rowblk, sc.partials, sc.ticket, d_x,
                                        d_y, d_w, nullptr, FinStore{d_out2}, (cudaStream_t)stream, d_cols16, nullptr,
                                        kbpeer::Wait{nullptr, nullptr, 0ull, nullptr, 0}, d_row_pat, d_pat_tab);
}
// 4-dot form: out4 = {<w,y>, <y,y>, <w2,y>, <w2,w>}  (merged BiCGSTAB update)
struct FinStore4 { double *out; __device__ void operator()(double a, double b, double c, double d) const { out[0] = a; out[1] = b; out[2] = c; out[3] = d; } };
extern "C" int kb_spmv_dots4(int64_t nrows, int64_t nnz, const int32_t *d_indptr, const int32_t *d_indices,
                             const int16_t *d_cols16, const uint8_t *d_row_pat, const int16_t *d_pat_tab,
                             const double *d_vals, const int32_t *d_rowblk, const double *d_x,
                             double *d_y, const double *d_w, const double *d_w2, double *d_out4, void *d_scratch,
                             void *stream) {
    if (!d_indptr || !d_indices || (!d_vals && nnz > 0) || !d_rowblk || !d_x || !d_y || !d_w || !d_w2 || !d_out4 ||
        !d_scratch)
        return KB_EINVAL;
    if (nrows <= 0) return 0;
    const Scratch sc = scratch_of(d_scratch);
    return launch_spmv_raw<2, FinStore4>(nrows, nnz, d_indptr, d_indices, d_vals, d_rowblk, sc.partials, sc.ticket, d_x,
                                         d_y, d_w, nullptr, FinStore4{d_out4}, (cudaStream_t)stream, d_cols16, d_w2,
                                         kbpeer::Wait{nullptr, nullptr, 0ull, nullptr, 0}, d_row_pat, d_pat_tab);
}
// x += alpha p + omega s ; r = s - omega t ; p = r + beta (p - omega v) ; out2[0] = <r,r>   (s is passed in r)
extern "C" int kb_bicgstab_update_p(int64_t n, const double *d_alpha, const double *d_omega, const double *d_beta,
                                    const double *d_t, const double *d_v, double *d_x, double *d_r, double *d_p,
                                    double *d_out2, void *d_scratch, void *stream) {
    if (!d_alpha || !d_omega || !d_beta || !d_t || !d_v || !d_x || !d_r || !d_p || !d_out2 || !d_scratch) return KB_EINVAL;
    const Scratch sc = scratch_of(d_scratch);
    return launch_vec4<1>(n, HBiUpdateP{d_alpha, d_omega, d_beta, d_t, d_v, d_x, d_r, d_p}, FinStore{d_out2}, sc.partials,
                          sc.ticket, nullptr, (cudaStream_t)stream);
}
// ---- scalar-fused forms for host-driven loops: the all-reduced dot products arrive as device pointers and every
// thread derives alpha / omega / beta from them (no tiny scalar kernels between the collectives and the vector work).
namespace {
struct HBiS2 {              // alpha = rho / d[0] ; r -= alpha v ; alpha is stored for the update kernel
    const double *rho, *d, *v; double *r, *alpha_out;
    struct S { double a; }; struct L { double r, v; };
    __device__ S scalars() const {
        const double a = rho[0] / d[0];
        if (blockIdx.x == 0 && threadIdx.x == 0) alpha_out[0] = a;
        return S{a};
    }
    __device__ void load(int64_t i, L &l) const { l.r = r[i]; l.v = __ldcs(v + i); }
    __device__ void apply(int64_t i, const L &l, const S &s, double &, double &) const { r[i] = l.r - s.a * l.v; }
};
struct HBiUpdateP2 {        // omega, rho_new, beta from d4 = {<s,t>,<t,t>,<rhat,t>,<rhat,s>} ; merged x / r / p update
    const double *rho_in, *alpha, *d4, *t, *v; double *x, *r, *p, *rho_out;
    struct S { double a, om, b; }; struct L { double x, p, s, t, v; };
    __device__ S scalars() const {
        const double om = d4[1] > 0.0 ? d4[0] / d4[1] : 0.0;
        const double rho_new = d4[3] - om * d4[2];
        const double a = alpha[0];
        const double b = (rho_new / rho_in[0]) * (a / om);
        if (blockIdx.x == 0 && threadIdx.x == 0) rho_out[0] = rho_new;     // rho_out != rho_in (double-buffered by the caller)
        return S{a, om, b};
    }
    __device__ void load(int64_t i, L &l) const { l.x = x[i]; l.p = p[i]; l.s = r[i]; l.t = __ldcs(t + i); l.v = __ldcs(v + i); }
    __device__ void apply(int64_t i, const L &l, const S &s, double &d0, double &) const {
        x[i] = l.x + (s.a * l.p + s.om * l.s);
        const double ri = l.s - s.om * l.t;
        r[i] = ri;
        p[i] = ri + s.b * (l.p - s.om * l.v);
        d0 += ri * ri;
    }
};
struct HCgUpdate2 {         // alpha = rho / d[0] ; x += alpha p ; r -= alpha q ; d0 = <r,r>
    const double *rho, *d, *p, *q; double *x, *r;
    struct S { double a; }; struct L { double x, p, r, q; };
    __device__ S scalars() const { return S{rho[0] / d[0]}; }
    __device__ void load(int64_t i, L &l) const { l.x = x[i]; l.p = p[i]; l.r = r[i]; l.q = __ldcs(q + i); }
    __device__ void apply(int64_t i, const L &l, const S &s, double &d0, double &) const {
        x[i] = l.x + s.a * l.p;
        const double ri = l.r - s.a * l.q;
        r[i] = ri;
        d0 += ri * ri;
    }
};
struct HCgP2 {              // beta = d[0] / rho_in ; p = r + beta p ; rho_out = d[0]
    const double *rho_in, *d, *r; double *p, *rho_out;
    struct S { double b; }; struct L { double r, p; };
    __device__ S scalars() const {
        const double b = d[0] / rho_in[0];
        if (blockIdx.x == 0 && threadIdx.x == 0) rho_out[0] = d[0];
        return S{b};
    }
    __device__ void load(int64_t i, L &l) const { l.r = r[i]; l.p = p[i]; }
    __device__ void apply(int64_t i, const L &l, const S &s, double &, double &) const { p[i] = l.r + s.b * l.p; }
};
}  // namespace
extern "C" int kb_bicgstab_s2(int64_t n, const double *d_rho, const double *d_dots, const double *d_v, double *d_r,
                              double *d_alpha_out, void *stream) {
    if (!d_rho || !d_dots || !d_v || !d_r || !d_alpha_out) return KB_EINVAL;
    return launch_vec4<0>(n, HBiS2{d_rho, d_dots, d_v, d_r, d_alpha_out}, NoFin{}, nullptr, nullptr, nullptr, (cudaStream_t)stream);
}
extern "C" int kb_bicgstab_update_p2(int64_t n, const double *d_rho_in, double *d_rho_out, const double *d_alpha,
                                     const double *d_dots4, const double *d_t, const double *d_v, double *d_x,
                                     double *d_r, double *d_p, double *d_out2, void *d_scratch, void *stream) {
    if (!d_rho_in || !d_rho_out || d_rho_in == d_rho_out || !d_alpha || !d_dots4 || !d_t || !d_v || !d_x || !d_r || !d_p ||
        !d_out2 || !d_scratch)
        return KB_EINVAL;
    const Scratch sc = scratch_of(d_scratch);
    return launch_vec4<1>(n, HBiUpdateP2{d_rho_in, d_alpha, d_dots4, d_t, d_v, d_x, d_r, d_p, d_rho_out}, FinStore{d_out2},
                          sc.partials, sc.ticket, nullptr, (cudaStream_t)stream);
}
extern "C" int kb_cg_update2(int64_t n, const double *d_rho, const double *d_dots, const double *d_p, const double *d_q,
                             double *d_x, double *d_r, double *d_out2, void *d_scratch, void *stream) {
    if (!d_rho || !d_dots || !d_p || !d_q || !d_x || !d_r || !d_out2 || !d_scratch) return KB_EINVAL;
    const Scratch sc = scratch_of(d_scratch);
    return launch_vec4<1>(n, HCgUpdate2{d_rho, d_dots, d_p, d_q, d_x, d_r}, FinStore{d_out2}, sc.partials, sc.ticket, nullptr,
                          (cudaStream_t)stream);
}
extern "C" int kb_cg_p2(int64_t n, const double *d_rho_in, double *d_rho_out, const double *d_dots, const double *d_r,
                        double *d_p, void *stream) {
    if (!d_rho_in || !d_rho_out || d_rho_in == d_rho_out || !d_dots || !d_r || !d_p) return KB_EINVAL;
    return launch_vec4<0>(n, HCgP2{d_rho_in, d_dots, d_r, d_p, d_rho_out}, NoFin{}, nullptr, nullptr, nullptr, (cudaStream_t)stream);
}
extern "C" int kb_dot2(int64_t n, const double *d_a, const double *d_b, double *d_out2, void *d_scratch, void *stream) {
    if (!d_a || !d_b || !d_out2 || !d_scratch) return KB_EINVAL;
    return launch_vec_s<1>(n, BodyDot2{d_a, d_b}, FinStore{d_out2}, scratch_of(d_scratch), (cudaStream_t)stream);
}
extern "C" int kb_residual(int64_t n, const double *d_b, const double *d_q, double *d_r, double *d_out2, void *d_scratch,
                           void *stream) {
    if (!d_b || !d_q || !d_r || !d_out2 || !d_scratch) return KB_EINVAL;
    return launch_vec_s<1>(n, BodyResidualP{d_b, d_q, d_r}, FinStore{d_out2}, scratch_of(d_scratch), (cudaStream_t)stream);
}
extern "C" int kb_axpy_neg(int64_t n, const double *d_alpha, const double *d_v, double *d_r, void *stream) {
    if (!d_alpha || !d_v || !d_r) return KB_EINVAL;
    return launch_vec4<0>(n, HBiS{d_alpha, d_v, d_r}, NoFin{}, nullptr, nullptr, nullptr, (cudaStream_t)stream);
}
extern "C" int kb_bicgstab_update(int64_t n, const double *d_alpha, const double *d_omega, const double *d_p,
                                  const double *d_t, const double *d_rhat, double *d_x, double *d_r, double *d_out2,
                                  void *d_scratch, void *stream) {
    if (!d_alpha || !d_omega || !d_p || !d_t || !d_rhat || !d_x || !d_r || !d_out2 || !d_scratch) return KB_EINVAL;
    const Scratch sc = scratch_of(d_scratch);
    return launch_vec4<1>(n, HBiUpdate{d_alpha, d_omega, d_p, d_t, d_rhat, d_x, d_r}, FinStore{d_out2}, sc.partials, sc.ticket,
                          nullptr, (cudaStream_t)stream);
}
extern "C" int kb_bicgstab_p(int64_t n, const double *d_beta, const double *d_omega, const double *d_r,
                             const double *d_v, double *d_p, void *stream) {
    if (!d_beta || !d_omega || !d_r || !d_v || !d_p) return KB_EINVAL;
    return launch_vec4<0>(n, HBiP{d_beta, d_omega, d_r, d_v, d_p}, NoFin{}, nullptr, nullptr, nullptr, (cudaStream_t)stream);
}
extern "C" int kb_cg_update(int64_t n, const double *d_alpha, const double *d_p, const double *d_q, double *d_x,
                            double *d_r, double *d_out2, void *d_scratch, void *stream) {
    if (!d_alpha || !d_p || !d_q || !d_x || !d_r || !d_out2 || !d_scratch) return KB_EINVAL;
    const Scratch sc = scratch_of(d_scratch);
    return launch_vec4<1>(n, HCgUpdate{d_alpha, d_p, d_q, d_x, d_r}, FinStore{d_out2}, sc.partials, sc.ticket, nullptr,
                          (cudaStream_t)stream);
}
extern "C" int kb_cg_p(int64_t n, const double *d_beta, const double *d_r, double *d_p, void *stream) {
    if (!d_beta || !d_r || !d_p) return KB_EINVAL;
    return launch_vec4<0>(n, HCgP{d_beta, d_r, d_p}, NoFin{}, nullptr, nullptr, nullptr, (cudaStream_t)stream);
}
extern "C" int kb_vec_mul(int64_t n, const double *d_a, const double *d_b, double *d_out, void *stream) {
    if (!d_a || !d_b || !d_out) return KB_EINVAL;
    return launch_vec_s<0>(n, BodyMulP{d_a, d_b, d_out}, NoFin{}, Scratch{nullptr, nullptr}, (cudaStream_t)stream);
}
// Jacobi scaling pieces: mode 0 -> scale = 1/diag (BiCGSTAB, column scaling), mode 1 -> 1/sqrt(diag) (CG, symmetric)
// 16-bit column offsets (col - row) of a CSR; *d_flag is set to 1 when some offset does not fit (then do not use them).
// With offsets the SpMV computes col = offset + (row index relative to the row view), so callers that pass a row-range
// view must also pass x advanced to the first row of the view.
namespace {
__global__ void __launch_bounds__(256) k_make_cols16(int64_t n, const int32_t *__restrict__ indptr,
                                                     const int32_t *__restrict__ indices, int16_t *__restrict__ out,
                                                     int *__restrict__ flag) {
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    bool wide = false;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride)
        for (int p = indptr[i], pe = indptr[i + 1]; p < pe; ++p) {
            const int d = indices[p] - (int)i;
            if (d < -32768 || d > 32767) wide = true;
            out[p] = (int16_t)d;
        }
    if (wide) *flag = 1;
}
}  // namespace
// Row-pattern dictionary of a CSR pattern (spmv_pat.cuh) for the SpMV building blocks: d_row_pat (nrows rounded up to 16,
// + 16 bytes), d_pat_keys (256 x 8 bytes scratch), d_pat_tab (256 x 16 int16), d_flag[0] = 1 when the matrix cannot be
// expressed (more than 256 patterns, a row longer than 16, an offset beyond 16 bits).
extern "C" int kb_make_row_patterns(int64_t nrows, const int32_t *d_indptr, const int32_t *d_indices, uint8_t *d_row_pat,
                                    void *d_pat_keys, int16_t *d_pat_tab, int32_t *d_flag, void *stream) {
    if (!d_indptr || !d_indices || !d_row_pat || !d_pat_keys || !d_pat_tab || !d_flag || nrows < 0) return KB_EINVAL;
    cudaStream_t s = (cudaStream_t)stream;
    // k_pattern_insert / k_pattern_verify raise slot [1] of the flag block they are given (slot [0] belongs to the 16-bit
    // band test of the solver set-up and is not touched by them): hand them d_flag - 1, so that slot is the caller's word
    int *flag2 = d_flag - 1;
    KB_CUDA(cudaMemsetAsync(d_flag, 0, sizeof(int), s));
    KB_CUDA(cudaMemsetAsync(d_pat_keys, 0, sizeof(unsigned long long) * P_SLOTS, s));
    KB_CUDA(cudaMemsetAsync(d_row_pat, 0, (size_t)((nrows + 15) / 16 * 16 + 16), s));
    if (nrows > 0) {
        const int g = kb_grid(nrows, 256, 8);
        k_pattern_insert<<<g, 256, 0, s>>>(nrows, d_indptr, d_indices, (unsigned long long *)d_pat_keys, d_pat_tab, d_row_pat, flag2);
        KB_LAUNCH_CHECK();
        k_pattern_verify<<<g, 256, 0, s>>>(nrows, d_indptr, d_indices, d_pat_tab, d_row_pat, flag2);
        KB_LAUNCH_CHECK();
    }
    return 0;
}

extern "C" int kb_make_cols16(int64_t nrows, const int32_t *d_indptr, const int32_t *d_indices, int16_t *d_cols16,
                              int32_t *d_flag, void *stream) {
    if (!d_indptr || !d_indices || !d_cols16 || !d_flag || nrows < 0) return KB_EINVAL;
    KB_CUDA(cudaMemsetAsync(d_flag, 0, sizeof(int32_t), (cudaStream_t)stream));
    if (nrows == 0) return 0;
    k_make_cols16<<<kb_grid(nrows, 256, 8), 256, 0, (cudaStream_t)stream>>>(nrows, d_indptr, d_indices, d_cols16, d_flag);
    KB_LAUNCH_CHECK();
    return 0;
}
extern "C" int kb_diag_scale(int64_t nrows, const int32_t *d_indptr, const int32_t *d_indices, const double *d_vals,
                             int mode, double *d_scale, void *stream) {
    if (!d_indptr || !d_indices || !d_vals || !d_scale || nrows < 0 || (mode != 0 && mode != 1)) return KB_EINVAL;
    if (nrows == 0) return 0;
    const int g = kb_grid(nrows, 256, 8);
    if (mode == 0) k_diag_scale<0><<<g, 256, 0, (cudaStream_t)stream>>>(nrows, d_indptr, d_indices, d_vals, d_scale);
    else k_diag_scale<1><<<g, 256, 0, (cudaStream_t)stream>>>(nrows, d_indptr, d_indices, d_vals, d_scale);
    KB_LAUNCH_CHECK();
    return 0;
}
extern "C" int kb_scale_matrix(int64_t nrows, const int32_t *d_indptr, const int32_t *d_indices, const double *d_vals,
                               int mode, const double *d_scale, double *d_out, void *stream) {
    if (!d_indptr || !d_indices || !d_vals || !d_scale || !d_out || nrows < 0 || (mode != 0 && mode != 1)) return KB_EINVAL;
    if (nrows == 0) return 0;
    const int g = kb_grid(nrows, 256, 8);
    if (mode == 0) k_scale_matrix<0><<<g, 256, 0, (cudaStream_t)stream>>>(nrows, d_indptr, d_indices, d_vals, d_scale, d_out);
    else k_scale_matrix<1><<<g, 256, 0, (cudaStream_t)stream>>>(nrows, d_indptr, d_indices, d_vals, d_scale, d_out);
    KB_LAUNCH_CHECK();
    return 0;
}

#include "krylov_peer.cuh"   // K9 peer-memory solvers (same translation unit: they use the launchers above)
#include "krylov_mg.cuh"     // opt-in: BiCGSTAB preconditioned by the multigrid V-cycle of multigrid.cu

extern "C" int kb_profile_spmv_begin(int every) {
    if (!g_prof.have_events) {
        for (int i = 0; i < 2 * SpmvProfile::CAP; ++i) KB_CUDA(cudaEventCreate(&g_prof.ev[i]));
        g_prof.have_events = true;
    }
    g_prof.every = every > 0 ? every : 1;
    g_prof.used = 0; g_prof.samples = 0; g_prof.total_ms = 0.0;
    g_prof.on = true;
    return 0;
}

extern "C" int kb_profile_spmv_end(int64_t *h_samples, double *h_total_ms) {
    g_prof.on = false;
    if (h_samples) *h_samples = g_prof.samples;
    if (h_total_ms) *h_total_ms = g_prof.total_ms;
    return 0;
}

extern "C" int64_t kb_spmv_num_blocks(int64_t nnz) { return num_blocks(nnz); }

extern "C" int kb_spmv_partition(int64_t nrows, int64_t nnz, const int32_t *d_indptr, int32_t *d_rowblk, void *stream) {
    if (!d_indptr || !d_rowblk || nrows < 0 || nnz < 0) return KB_EINVAL;
    const int64_t nblk = num_blocks(nnz);
    k_partition<<<(unsigned)kb_cdiv(nblk + 1, 256), 256, 0, (cudaStream_t)stream>>>(nrows, nblk, d_indptr, d_rowblk);
    KB_LAUNCH_CHECK();
    return 0;
}

extern "C" int kb_spmv(int64_t nrows, int64_t nnz, const int32_t *d_indptr, const int32_t *d_indices,
                       const double *d_vals, const int32_t *d_rowblk, const double *d_x, double *d_y, void *stream) {
    if (!d_indptr || !d_indices || (!d_vals && nnz > 0) || !d_rowblk || !d_x || !d_y || nrows < 0) return KB_EINVAL;
    if (nrows == 0) return 0;
    return launch_spmv_raw<0, NoFinal>(nrows, nnz, d_indptr, d_indices, d_vals, d_rowblk, nullptr, nullptr, d_x, d_y,
                                       nullptr, nullptr, NoFinal{}, (cudaStream_t)stream);
}

/* development switch: 1 forces the one-shot (non-TMA) SpMV kernel */
extern "C" void kb_spmv_force_simple(int on) { g_force_simple_spmv = on; }
/* development switch: 1 keeps 32-bit column indices in the solvers even when the 16-bit offset form is available */
extern "C" void kb_spmv_force_idx32(int on) { g_force_idx32 = on; }
extern "C" void kb_spmv_force_no_patterns(int on) { g_no_patterns = on; }

extern "C" size_t kb_krylov_workspace_bytes(int64_t nrows, int64_t nnz) { return carve(nullptr, nullptr, nrows, nnz); }

extern "C" int kb_cg_solve(int64_t nrows, int64_t nnz, const int32_t *d_indptr, const int32_t *d_indices,
                           const double *d_vals, const double *d_b, double *d_x, double rtol, int maxiter,
                           int check_every, void *d_workspace, size_t workspace_bytes, kb_solve_info *h_info,
                           void *stream) {
    return krylov_solve<0>(nrows, nnz, d_indptr, d_indices, d_vals, d_b, d_x, rtol, maxiter, check_every, d_workspace,
                           workspace_bytes, h_info, (cudaStream_t)stream);
}

extern "C" int kb_bicgstab_solve(int64_t nrows, int64_t nnz, const int32_t *d_indptr, const int32_t *d_indices,
                                 const double *d_vals, const double *d_b, double *d_x, double rtol, int maxiter,
                                 int check_every, void *d_workspace, size_t workspace_bytes, kb_solve_info *h_info,
                                 void *stream) {
    return krylov_solve<1>(nrows, nnz, d_indptr, d_indices, d_vals, d_b, d_x, rtol, maxiter, check_every, d_workspace,
                           workspace_bytes, h_info, (cudaStream_t)stream);
}
