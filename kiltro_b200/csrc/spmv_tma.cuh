// K5 (production form): persistent, TMA-fed CSR SpMV for sm_100a.
//
// Same row-block decomposition as spmv.cuh (a row block = ~NNZ_TARGET non-zeros, kb_spmv_partition), but the CTAs are
// persistent (2 per SM), each walks a contiguous range of row blocks, and a dedicated producer warp streams the block's
// values / column indices / row pointers into a ring of shared-memory stages with 1-D bulk async copies
// (cp.async.bulk.shared.global, SASS UBLKCP) that complete on mbarriers.  The 8 consumer warps never wait on DRAM: they
// read the staged slice, gather x through L1/L2, overwrite the staged values with the products in place, and one thread
// per row adds its products left to right (fixed order, reproducible).  Stages are handed back through "empty"
// mbarriers.  Why: the one-shot kernel has ~5 dependent DRAM round trips per CTA lifetime and reached 48 % of the HBM
// roofline at 16 M rows (profiles/r01_*); with S stages in flight the copy engine keeps S*35 KB per CTA outstanding.
//
// Alignment contract of bulk copies (16-byte global/shared addresses, 16-byte multiples): slices are widened downwards
// to a multiple of 4 entries and cut at the last multiple of 4; the <= 3 trailing entries are read with ordinary loads.
// vals / indices / indptr must be 16-byte aligned (checked by the launcher, which otherwise uses spmv.cuh).
#pragma once
#include "common.cuh"
#include "spmv.cuh"
#include "peer.cuh"

namespace kbspmv {

constexpr int T_CONSUMERS = 256;                 // 8 consumer warps
constexpr int T_THREADS = T_CONSUMERS + 32;      // + 1 producer warp
constexpr int T_STAGES = 3;
constexpr int T_VCAP = 2560;                     // staged entries per stage (NNZ_TARGET 2304 + longest fast-path row + pad)
constexpr int T_BATCH = 9;                       // gathers in flight per row (a quad-4 row has 9 entries)
constexpr int T_RCAP = 1032;                     // staged row pointers per stage

struct __align__(16) StageMeta { int r0, r1, p0, p1, pa, pcut, ra, rcut, slow, pad0, pad1, pad2; };

// IDX = int: column indices as stored in the CSR; IDX = short: 16-bit offsets (col - row), available when the matrix
// bandwidth fits (structured FE numberings do): 10 instead of 12 bytes per non-zero cross HBM.  Measured at 16 M rows
// inside BiCGSTAB: 0.321 ms (16-bit) vs 0.352 ms (32-bit) per launch; the solvers pick the 16-bit form when every
// |col - row| fits (checked on device while the Jacobi-scaled copy is made).
template <class IDX>
struct __align__(128) StageT {
    double vals[T_VCAP];
    IDX cols[T_VCAP];
    int rptr[T_RCAP];
    StageMeta meta;
};

template <class IDX> constexpr size_t tma_smem_bytes() {
    return sizeof(StageT<IDX>) * T_STAGES + 2 * T_STAGES * sizeof(uint64_t) + 512;
}

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t *bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_%=:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra DONE_%=;\n"
        "bra WAIT_%=;\n"
        "DONE_%=:\n"
        "}\n" ::"r"(smem_u32(bar)), "r"(parity)
        : "memory");
}
__device__ __forceinline__ void bulk_g2s(void *dst, const void *src, uint32_t bytes, uint64_t *bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     smem_u32(dst)),
                 "l"(src), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}
// L2 residency hints: the matrix streams through once (evict-first) so that the x vector, whose rows are re-used by the
// row blocks one mesh row later, stays resident (evict-last) instead of being pushed out by 1.7 GB of values/indices.
__device__ __forceinline__ uint64_t policy_evict_first() {
    uint64_t p;
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(p));
    return p;
}
__device__ __forceinline__ uint64_t policy_evict_last() {
    uint64_t p;
    asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(p));
    return p;
}
__device__ __forceinline__ void bulk_g2s_hint(void *dst, const void *src, uint32_t bytes, uint64_t *bar, uint64_t pol) {
    asm volatile(
        "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;" ::"r"(
            smem_u32(dst)),
        "l"(src), "r"(bytes), "r"(smem_u32(bar)), "l"(pol)
        : "memory");
}
__device__ __forceinline__ double ldg_hint(const double *p, uint64_t pol) {
    double v;
    asm volatile("ld.global.nc.L2::cache_hint.f64 %0, [%1], %2;" : "=d"(v) : "l"(p), "l"(pol));
    return v;
}
__device__ __forceinline__ void consumer_sync() { asm volatile("bar.sync 1, %0;" ::"n"(T_CONSUMERS) : "memory"); }

// sum of one double per consumer thread (fixed order); result valid in thread 0.  red: 8 doubles.
__device__ __forceinline__ double consumer_sum(double v, double *red) {
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    v = warp_sum(v);
    if (lane == 0) red[w] = v;
    consumer_sync();
    double r = 0.0;
    if (w == 0) {
        r = lane < (T_CONSUMERS / 32) ? red[lane] : 0.0;
        r = warp_sum(r);
    }
    consumer_sync();
    return r;
}

template <int DOTS, class Final, class IDX>
__global__ void __launch_bounds__(T_THREADS, 2)
k_spmv_tma(int64_t nblk, const int32_t *__restrict__ indptr, const IDX *__restrict__ indices,
           const double *__restrict__ vals, const int32_t *__restrict__ rowblk, const double *__restrict__ x,
           double *__restrict__ y, const double *__restrict__ w, double *__restrict__ partials,
           unsigned int *__restrict__ ticket, const int *__restrict__ done, Final fin,
           const double *__restrict__ w2 = nullptr, kbpeer::Wait pwait = kbpeer::Wait{nullptr, nullptr, 0ull, nullptr, 0}) {
    // DOTS: 0 none; 1 -> {<w,y>, <y,y>}; 2 -> additionally {<w2,y>, <w2,w>} (the merged BiCGSTAB update, krylov.cu)
    // pwait: halo flags raised by the neighbouring ranks (peer.cuh); the kernel does not touch x before they are up
    if (done != nullptr && *done != 0) return;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    typedef StageT<IDX> Stage;
    constexpr int GRAN = 16 / (int)sizeof(IDX);            // entries per 16 bytes of the index array (bulk-copy granule)
    constexpr bool REL = sizeof(IDX) == 2;                 // 16-bit indices are offsets relative to the row
    Stage *stages = reinterpret_cast<Stage *>(smem_raw);
    uint64_t *full = reinterpret_cast<uint64_t *>(smem_raw + sizeof(Stage) * T_STAGES);
    uint64_t *empty = full + T_STAGES;
    double *red = reinterpret_cast<double *>(empty + T_STAGES);
    int *s_last = reinterpret_cast<int *>(red + 16);

    // contiguous range of row blocks of this CTA
    const int64_t b_begin = nblk * (int64_t)blockIdx.x / gridDim.x;
    const int64_t b_end = nblk * (int64_t)(blockIdx.x + 1) / gridDim.x;
    const int nmine = (int)(b_end - b_begin);

    if (threadIdx.x == 0) {
        for (int s = 0; s < T_STAGES; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], T_CONSUMERS / 32); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        if (pwait.f0 != nullptr || pwait.f1 != nullptr)
            kbpeer::wait_flags_rows(pwait, __ldg(rowblk + b_begin), __ldg(rowblk + b_end), __ldg(rowblk + nblk));
    }
    __syncthreads();

    if (threadIdx.x >= T_CONSUMERS) {
        // ------------------------------------------------------------------ producer warp
        // All 32 lanes fetch the (row, nnz) boundaries of the next 32 row blocks at once (one dependent-load chain per
        // 32 blocks instead of one per block); lane 0 arms the barriers and issues the bulk copies.
        const int lane = threadIdx.x & 31;
        const uint64_t pol_stream = policy_evict_first();
        // Slices are widened UP to the copy granule as well whenever the extra entries still lie inside the arrays (they
        // belong to the next row block): only the very last block of the matrix keeps a tail that consumers must fetch
        // with ordinary loads -- otherwise every block would stall one warp on a dependent global load.
        const int nrows_total = __ldg(rowblk + nblk);
        const int p_lim = __ldg(indptr + nrows_total) & ~(GRAN - 1);
        const int r_lim = (nrows_total + 1) & ~3;
        int r0 = 0, p0 = 0;
        if (nmine > 0) { r0 = __ldg(rowblk + b_begin); p0 = __ldg(indptr + r0); }
        int rn = 0, pn = 0;                                     // boundaries of blocks [base+1 .. base+32]
        for (int base = 0; base < nmine; base += 32) {
            if (base + lane < nmine) { rn = __ldg(rowblk + b_begin + base + lane + 1); pn = __ldg(indptr + rn); }
            const int cntb = (nmine - base) < 32 ? (nmine - base) : 32;
            for (int k = 0; k < cntb; ++k) {
                const int i = base + k;
                const int r1 = __shfl_sync(KB_FULL, rn, k), p1 = __shfl_sync(KB_FULL, pn, k);
                if (lane == 0) {
                    const int s = i % T_STAGES;
                    if (i >= T_STAGES) mbar_wait(&empty[s], ((i / T_STAGES) - 1) & 1);
                    Stage &st = stages[s];
                    StageMeta m;
                    m.r0 = r0; m.r1 = r1; m.p0 = p0; m.p1 = p1;
                    const int p1up = (p1 + GRAN - 1) & ~(GRAN - 1), r1up = (r1 + 1 + 3) & ~3;
                    m.pa = p0 & ~(GRAN - 1); m.pcut = p1up < p_lim ? p1up : p_lim;
                    if (m.pcut < m.pa) m.pcut = m.pa;
                    m.ra = r0 & ~3; m.rcut = r1up < r_lim ? r1up : r_lim;
                    if (m.rcut < m.ra) m.rcut = m.ra;
                    m.slow = (p1up - m.pa) > T_VCAP ? 1 : 0;
                    const bool rows_staged = (r1up - m.ra) <= T_RCAP && !m.slow;
                    if (!rows_staged) m.rcut = m.ra;            // nothing staged: consumers read indptr directly
                    m.pad0 = m.pad1 = m.pad2 = 0;
                    st.meta = m;
                    const uint32_t nv = m.slow ? 0u : (uint32_t)(m.pcut - m.pa);
                    const uint32_t nr = (uint32_t)(m.rcut - m.ra);
                    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // stage was last written by consumers
                    mbar_expect_tx(&full[s], nv * (8u + (uint32_t)sizeof(IDX)) + nr * 4u);
                    if (nv) {
                        bulk_g2s_hint(st.vals, vals + m.pa, nv * 8u, &full[s], pol_stream);
                        bulk_g2s_hint(st.cols, indices + m.pa, nv * (uint32_t)sizeof(IDX), &full[s], pol_stream);
                    }
                    if (nr) bulk_g2s(st.rptr, indptr + m.ra, nr * 4u, &full[s]);
                }
                r0 = r1; p0 = p1;
                __syncwarp();
            }
        }
    } else {
        // ------------------------------------------------------------------ consumers
        double d0 = 0.0, d1 = 0.0, d2 = 0.0, d3 = 0.0;
        const int tid = threadIdx.x;
        const uint64_t pol_keep = policy_evict_last();
        for (int i = 0; i < nmine; ++i) {
            const int s = i % T_STAGES;
            mbar_wait(&full[s], (i / T_STAGES) & 1);
            Stage &st = stages[s];
            const StageMeta m = st.meta;
            if (!m.slow) {
                // One thread per row, straight from the staged slice: lanes of a warp read row-strided shared memory
                // (9 doubles apart for quad-4: conflict-free) and gather x at consecutive addresses, so the products
                // never travel back through shared memory and the L1 sees coalesced gathers.
                const int jcut = m.pcut - m.pa;
                for (int r = m.r0 + tid; r < m.r1; r += T_CONSUMERS) {
                    double wv = 0.0, w2v = 0.0;
                    if (DOTS) wv = __ldcs(w + r);
                    if (DOTS == 2) w2v = __ldg(w2 + r);
                    const int ia = r - m.ra;
                    const int a = ((r < m.rcut) ? st.rptr[ia] : __ldg(indptr + r)) - m.pa;
                    const int b = ((r + 1 < m.rcut) ? st.rptr[ia + 1] : __ldg(indptr + r + 1)) - m.pa;
                    double sum = 0.0;
                    int q = a;
                    const int bs = b < jcut ? b : jcut;             // staged part of the row
                    // all gathers of a batch are issued before the first FMA: one L1/L2 round trip per <= T_BATCH entries
                    for (; q < bs; q += T_BATCH) {
                        int c[T_BATCH];
                        double xv[T_BATCH];
#pragma unroll
                        for (int k = 0; k < T_BATCH; ++k) c[k] = (int)st.cols[(q + k < bs) ? q + k : q] + (REL ? r : 0);
#pragma unroll
                        for (int k = 0; k < T_BATCH; ++k) xv[k] = ldg_hint(x + c[k], pol_keep);
#pragma unroll
                        for (int k = 0; k < T_BATCH; ++k)
                            if (q + k < bs) sum = fma(st.vals[q + k], xv[k], sum);
                    }
                    q = bs > a ? bs : a;
                    for (; q < b; ++q)                                                      // < GRAN tail entries
                        sum = fma(__ldg(vals + m.pa + q), __ldg(x + ((int)__ldg(indices + m.pa + q) + (REL ? r : 0))), sum);
                    y[r] = sum;
                    if (DOTS) { d0 += wv * sum; d1 += sum * sum; }
                    if (DOTS == 2) { d2 += w2v * sum; d3 += w2v * wv; }
                }
            } else {
                // very long row(s): the consumers walk the rows of this block straight from global memory
                for (int r = m.r0; r < m.r1; ++r) {
                    const int a = __ldg(indptr + r), b = __ldg(indptr + r + 1);
                    double sum = 0.0;
                    for (int q = a + tid; q < b; q += T_CONSUMERS) sum += vals[q] * __ldg(x + ((int)indices[q] + (REL ? r : 0)));
                    sum = consumer_sum(sum, red);
                    if (tid == 0) {
                        y[r] = sum;
                        if (DOTS) { d0 += __ldg(w + r) * sum; d1 += sum * sum; }
                        if (DOTS == 2) { d2 += __ldg(w2 + r) * sum; d3 += __ldg(w2 + r) * __ldg(w + r); }
                    }
                }
            }
            __syncwarp();
            if ((tid & 31) == 0) mbar_arrive(&empty[s]);
        }
        if (DOTS) {
            d0 = consumer_sum(d0, red);
            d1 = consumer_sum(d1, red);
            if (DOTS == 2) { d2 = consumer_sum(d2, red); d3 = consumer_sum(d3, red); }
            if (tid == 0) {
                partials[blockIdx.x] = d0;
                partials[gridDim.x + blockIdx.x] = d1;
                if (DOTS == 2) { partials[2 * gridDim.x + blockIdx.x] = d2; partials[3 * gridDim.x + blockIdx.x] = d3; }
                __threadfence();
                const unsigned int t = atomicInc(ticket, gridDim.x - 1);
                *s_last = (t == gridDim.x - 1);
            }
            consumer_sync();
            if (*s_last) {
                __threadfence();
                double a0 = 0.0, a1 = 0.0, a2 = 0.0, a3 = 0.0;
                for (int k = tid; k < (int)gridDim.x; k += T_CONSUMERS) {
                    a0 += __ldcg(partials + k);
                    a1 += __ldcg(partials + gridDim.x + k);
                    if (DOTS == 2) { a2 += __ldcg(partials + 2 * gridDim.x + k); a3 += __ldcg(partials + 3 * gridDim.x + k); }
                }
                a0 = consumer_sum(a0, red);
                a1 = consumer_sum(a1, red);
                if (DOTS == 2) { a2 = consumer_sum(a2, red); a3 = consumer_sum(a3, red); }
                if (tid == 0) fin(a0, a1, a2, a3);
            }
        }
    }
}

static inline bool tma_eligible(const void *indptr, const void *indices, const void *vals) {
    return (((uintptr_t)indptr | (uintptr_t)indices | (uintptr_t)vals) & 15u) == 0;
}

static inline int tma_grid(int64_t nblk) {
    const int64_t cap = 2 * (int64_t)kb_num_sms();
    return (int)(nblk < cap ? (nblk > 0 ? nblk : 1) : cap);
}

}  // namespace kbspmv
