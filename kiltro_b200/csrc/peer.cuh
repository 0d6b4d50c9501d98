// NVLink / NVSwitch peer-memory exchange used by the row-block partitioned solvers (K9): instead of one NCCL collective
// per halo exchange and per dot product, the kernels themselves move the data.
//   * dot products: the last CTA of a kernel writes its rank's partial sums into a slot of EVERY rank's control block
//     (peer stores over NVLink), fences, then raises a per-rank sequence flag; the consumer kernel on each rank waits for
//     all P flags and adds the P partials in rank order (deterministic all-reduce, ~NVLink latency, no extra launch);
//   * halo rows: the vector kernel that produces the SpMV input writes its first / last owned mesh row straight into the
//     neighbours' halo slots of the same vector (the vectors live in IPC-mapped memory) and the last CTA raises the
//     neighbours' halo flags; the SpMV waits for its two flags before it starts.
// Every wait has a clock-based timeout that sets an error word instead of hanging the GPU.
// Ordering: data stores -> __threadfence_system() -> flag store (release); reader: acquire load of the flag, then reads.
// Slots are double-buffered by event parity: a rank can only post event e+2 after it consumed e+1, which every rank
// posts only after consuming e, so a slot is never overwritten before all readers are done with it.
#pragma once
#include "common.cuh"

namespace kbpeer {

constexpr int MAXP = 8;
// control block layout, in 8-byte words
constexpr int W_DOTS = 0;                 // [parity 2][rank MAXP][4]
constexpr int W_DOTFLAG = 64;             // [parity 2][rank MAXP]
constexpr int W_HALOFLAG = 80;            // [vector NVEC][side 2]   side 0: written by the rank below, 1: by the rank above
constexpr int NVEC = 3;                   // exchanged vectors per region: 0 = p, 1 = r (BiCGSTAB's second SpMV input), 2 = T (transient)
constexpr int W_ERROR = 96;
constexpr int CTRL_WORDS = 128;

struct Ctx {
    int P, rank;
    unsigned long long *ctrl[MAXP];       // control block of every rank as mapped into this process (ctrl[rank] = own)
};

struct Wait {                             // flags (in the own control block) an SpMV waits for before it gathers x
    const unsigned long long *f0, *f1;    // f0: raised by the rank below (halo row under the first owned mesh row), f1: above
    unsigned long long val;
    unsigned long long *err;
    int edge;                             // rows per mesh row: only the CTAs whose rows touch the first / last `edge` rows of
                                          // the view read halo entries and have to wait (0 = every CTA waits)
};

__device__ __forceinline__ unsigned long long ld_acquire_sys(const unsigned long long *p) {
    unsigned long long v;
    asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_release_sys(unsigned long long *p, unsigned long long v) {
    asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}

// returns false on timeout (~2 s at 2 GHz)
__device__ __forceinline__ bool spin_until_ge(const unsigned long long *flag, unsigned long long val) {
    if (flag == nullptr) return true;
    const long long t0 = clock64();
    while (ld_acquire_sys(flag) < val) {
        if (clock64() - t0 > 4000000000LL) return false;
        __nanosleep(32);
    }
    return true;
}

// once a wait of this rank has timed out, later waits give up at once (the host returns KB_EPEER at its next look)
__device__ __forceinline__ bool already_failed(const unsigned long long *err) {
    return err != nullptr && *reinterpret_cast<const volatile unsigned long long *>(err) != 0ull;
}

__device__ __forceinline__ void wait_flags(const Wait &w) {
    if (w.f0 == nullptr && w.f1 == nullptr) return;
    if (already_failed(w.err)) return;
    if (!spin_until_ge(w.f0, w.val) || !spin_until_ge(w.f1, w.val)) {
        if (w.err) *w.err = 1ull;
    }
}
// The CTA owning rows [r_first, r_last) of an nrows-row view waits only for the halo rows it will read: rows of the first
// mesh row reference the halo row below (f0), rows of the last mesh row the one above (f1); every other CTA starts at once.
__device__ __forceinline__ void wait_flags_rows(const Wait &w, int r_first, int r_last, int nrows) {
    if (w.f0 == nullptr && w.f1 == nullptr) return;
    if (w.edge <= 0) { wait_flags(w); return; }
    if (r_last <= r_first || already_failed(w.err)) return;
    bool ok = true;
    if (r_first < w.edge) ok = spin_until_ge(w.f0, w.val) && ok;
    if (r_last > nrows - w.edge) ok = spin_until_ge(w.f1, w.val) && ok;
    if (!ok && w.err) *w.err = 1ull;
}

__device__ __forceinline__ unsigned long long ld_relaxed_sys(const unsigned long long *p) {
    unsigned long long v;
    asm volatile("ld.relaxed.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_relaxed_sys(unsigned long long *p, unsigned long long v) {
    asm volatile("st.relaxed.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ void fence_acq_rel_sys() { asm volatile("fence.acq_rel.sys;" ::: "memory"); }

// post this rank's partial sums of event `ev` to every rank (one thread).  ONE fence between the data stores and the flag
// stores (fence + relaxed store = a release): P release stores in a row each wait for the one before to be performed
// across NVLink -- P serialised round trips per reduction, the part of the iteration overhead that grew with P.
__device__ __forceinline__ void post_dots(const Ctx &c, unsigned long long ev, double a0, double a1, double a2, double a3) {
    const int par = (int)(ev & 1ull);
    for (int r = 0; r < c.P; ++r) {
        double *slot = reinterpret_cast<double *>(c.ctrl[r] + W_DOTS) + (par * MAXP + c.rank) * 4;
        slot[0] = a0; slot[1] = a1; slot[2] = a2; slot[3] = a3;
    }
    fence_acq_rel_sys();
    for (int r = 0; r < c.P; ++r) st_relaxed_sys(c.ctrl[r] + W_DOTFLAG + par * MAXP + c.rank, ev);
}

// wait for the partial sums of event `ev` from all ranks and add them in rank order (one thread); false on timeout.
// The P flags are polled together with relaxed loads (independent, in flight at once) and ONE acquire fence follows the
// successful poll, instead of P acquire loads one after the other.
__device__ __forceinline__ bool collect_dots(const Ctx &c, unsigned long long ev, double (&d)[4]) {
    const int par = (int)(ev & 1ull);
    unsigned long long *own = c.ctrl[c.rank];
    bool ok = !already_failed(own + W_ERROR);
    d[0] = d[1] = d[2] = d[3] = 0.0;
    if (ok) {
        const unsigned long long *flags = own + W_DOTFLAG + par * MAXP;
        const long long t0 = clock64();
        for (;;) {
            unsigned long long f[MAXP];
#pragma unroll
            for (int r = 0; r < MAXP; ++r) f[r] = r < c.P ? ld_relaxed_sys(flags + r) : ev;
            bool all = true;
#pragma unroll
            for (int r = 0; r < MAXP; ++r) all = all && (f[r] >= ev);
            if (all) break;
            if (clock64() - t0 > 4000000000LL) { ok = false; break; }
            __nanosleep(32);
        }
        fence_acq_rel_sys();
    }
    for (int r = 0; r < c.P; ++r) {
        const volatile double *slot = reinterpret_cast<const volatile double *>(own + W_DOTS) + (par * MAXP + r) * 4;
        d[0] += slot[0]; d[1] += slot[1]; d[2] += slot[2]; d[3] += slot[3];
    }
    if (!ok) own[W_ERROR] = 2ull;
    return ok;
}

}  // namespace kbpeer
