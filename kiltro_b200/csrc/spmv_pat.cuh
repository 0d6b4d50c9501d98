// K5, row-pattern form: the persistent TMA-fed CSR SpMV of spmv_tma.cuh for matrices whose rows repeat a small set of
// column-offset patterns (every structured FE numbering: the 9-point stencil of Mesh.Rectangle has a handful of distinct
// (col - row) tuples, a line mesh three).  The column indices are not streamed at all: HBM delivers 8 bytes per non-zero
// (the value) plus, per row, the row pointer and ONE byte naming the row's pattern; the <= 256 patterns (16 offsets each)
// sit in shared memory.  12 -> 10 (16-bit offsets) -> 8 bytes per non-zero.  Same row-block decomposition, same
// one-row-per-thread consumers, same left-to-right summation and fused dot products as k_spmv_tma: results are
// bit-identical.  The solver setup builds the dictionary on device (k_pattern_insert / k_pattern_verify in krylov.cu)
// and falls back to the 16-bit / 32-bit kernels when a matrix has more than 256 patterns or a row longer than 16.
#pragma once
#include "spmv_tma.cuh"

namespace kbspmv {

constexpr int P_SLOTS = 256;                     // dictionary size (pattern id = one byte)
constexpr int P_MAXLEN = 16;                     // offsets per pattern
constexpr int P_PATCAP = T_RCAP + 24;            // staged pattern bytes per stage (row slice widened to 16-byte granules)
constexpr int P_STAGES = 4;                      // one stage more than spmv_tma.cuh: a stage carries ~20 % fewer bytes, and
                                                 // the copy engine needs the same number of bytes in flight to cover DRAM
                                                 // latency (with 3 stages the kernel ran at 70 % instead of 77 % of DRAM peak)
typedef int16_t pat_off_t;                       // offsets are stored in 16 bits (the builder checks that they fit)
constexpr int P_CONSUMERS = 256;                 // 8 consumer warps = the rows of one ~2304-nnz row block (one row per thread).
                                                 // More warps would only idle unless the row blocks grew with them, and 4
                                                 // stages of larger blocks no longer fit two CTAs per SM.
constexpr int P_THREADS = P_CONSUMERS + 32;      // + 1 producer warp

static_assert(P_CONSUMERS == T_CONSUMERS, "consumer_sync / consumer_sum of spmv_tma.cuh are reused");

struct __align__(128) StageP {
    double vals[T_VCAP];
    int rptr[T_RCAP];
    uint8_t pat[P_PATCAP];
    StageMeta meta;                              // pad0 / pad1: first / one-past-last row of the staged pattern bytes
};

constexpr size_t pat_smem_bytes() {
    return sizeof(StageP) * P_STAGES + 2 * P_STAGES * sizeof(uint64_t) + 512 + sizeof(pat_off_t) * P_SLOTS * P_MAXLEN;
}

template <int DOTS, class Final>
__global__ void __launch_bounds__(P_THREADS, 2)
k_spmv_pat(int64_t nblk, const int32_t *__restrict__ indptr, const uint8_t *__restrict__ row_pat,
           const pat_off_t *__restrict__ pat_tab, const double *__restrict__ vals, const int32_t *__restrict__ rowblk,
           const double *__restrict__ x, double *__restrict__ y, const double *__restrict__ w,
           double *__restrict__ partials, unsigned int *__restrict__ ticket, const int *__restrict__ done, Final fin,
           const double *__restrict__ w2 = nullptr, kbpeer::Wait pwait = kbpeer::Wait{nullptr, nullptr, 0ull, nullptr, 0}) {
    // pwait: halo flags raised by the neighbouring ranks (peer.cuh); the kernel does not touch x before they are up
    extern __shared__ __align__(128) unsigned char smem_raw[];
    constexpr int GRAN = 2;                                // values per 16 bytes (bulk-copy granule)
    StageP *stages = reinterpret_cast<StageP *>(smem_raw);
    uint64_t *full = reinterpret_cast<uint64_t *>(smem_raw + sizeof(StageP) * P_STAGES);
    uint64_t *empty = full + P_STAGES;
    double *red = reinterpret_cast<double *>(empty + P_STAGES);
    int *s_last = reinterpret_cast<int *>(red + 16);
    pat_off_t *s_tab = reinterpret_cast<pat_off_t *>(smem_raw + sizeof(StageP) * P_STAGES + 2 * P_STAGES * sizeof(uint64_t) + 512);

    const int64_t b_begin = nblk * (int64_t)blockIdx.x / gridDim.x;
    const int64_t b_end = nblk * (int64_t)(blockIdx.x + 1) / gridDim.x;
    const int nmine = (int)(b_end - b_begin);

    // set-up that does not depend on the previous kernel of the stream (the dictionary is written once per solver
    // set-up): with a programmatic dependent launch it overlaps that kernel's tail
    for (int i = threadIdx.x; i < P_SLOTS * P_MAXLEN; i += P_THREADS) s_tab[i] = __ldg(pat_tab + i);
    if (threadIdx.x == 0) {
        for (int s = 0; s < P_STAGES; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], P_CONSUMERS / 32); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    kb_pdl_wait();
    if (done != nullptr && *done != 0) return;             // uniform over the grid
    if (threadIdx.x == 0 && (pwait.f0 != nullptr || pwait.f1 != nullptr))
        kbpeer::wait_flags_rows(pwait, __ldg(rowblk + b_begin), __ldg(rowblk + b_end), __ldg(rowblk + nblk));
    __syncthreads();

    if (threadIdx.x >= P_CONSUMERS) {
        // ------------------------------------------------------------------ producer warp (see spmv_tma.cuh)
        const int lane = threadIdx.x & 31;
        const uint64_t pol_stream = policy_evict_first();
        const int nrows_total = __ldg(rowblk + nblk);
        const int p_lim = __ldg(indptr + nrows_total) & ~(GRAN - 1);
        const int r_lim = (nrows_total + 1) & ~3;
        const int pat_lim = (nrows_total + 15) & ~15;           // row_pat is allocated (and zero-filled) up to here
        int r0 = 0, p0 = 0;
        if (nmine > 0) { r0 = __ldg(rowblk + b_begin); p0 = __ldg(indptr + r0); }
        int rn = 0, pn = 0;
        for (int base = 0; base < nmine; base += 32) {
            if (base + lane < nmine) { rn = __ldg(rowblk + b_begin + base + lane + 1); pn = __ldg(indptr + rn); }
            const int cntb = (nmine - base) < 32 ? (nmine - base) : 32;
            for (int k = 0; k < cntb; ++k) {
                const int i = base + k;
                const int r1 = __shfl_sync(KB_FULL, rn, k), p1 = __shfl_sync(KB_FULL, pn, k);
                if (lane == 0) {
                    const int s = i % P_STAGES;
                    if (i >= P_STAGES) mbar_wait(&empty[s], ((i / P_STAGES) - 1) & 1);
                    StageP &st = stages[s];
                    StageMeta m;
                    m.r0 = r0; m.r1 = r1; m.p0 = p0; m.p1 = p1;
                    const int p1up = (p1 + GRAN - 1) & ~(GRAN - 1), r1up = (r1 + 1 + 3) & ~3;
                    m.pa = p0 & ~(GRAN - 1); m.pcut = p1up < p_lim ? p1up : p_lim;
                    if (m.pcut < m.pa) m.pcut = m.pa;
                    m.ra = r0 & ~3; m.rcut = r1up < r_lim ? r1up : r_lim;
                    if (m.rcut < m.ra) m.rcut = m.ra;
                    m.slow = (p1up - m.pa) > T_VCAP ? 1 : 0;      // cannot happen with rows <= 16 entries; kept as a guard
                    m.pad0 = r0 & ~15;
                    const int rp1 = (r1 + 15) & ~15;
                    m.pad1 = rp1 < pat_lim ? rp1 : pat_lim;
                    if (m.pad1 < m.pad0) m.pad1 = m.pad0;
                    const bool rows_staged = (r1up - m.ra) <= T_RCAP && (m.pad1 - m.pad0) <= P_PATCAP && !m.slow;
                    if (!rows_staged) { m.rcut = m.ra; m.pad1 = m.pad0; }   // consumers read indptr / row_pat directly
                    m.pad2 = 0;
                    st.meta = m;
                    const uint32_t nv = m.slow ? 0u : (uint32_t)(m.pcut - m.pa);
                    const uint32_t nr = (uint32_t)(m.rcut - m.ra), np = (uint32_t)(m.pad1 - m.pad0);
                    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                    mbar_expect_tx(&full[s], nv * 8u + nr * 4u + np);
                    if (nv) bulk_g2s_hint(st.vals, vals + m.pa, nv * 8u, &full[s], pol_stream);
                    if (nr) bulk_g2s(st.rptr, indptr + m.ra, nr * 4u, &full[s]);
                    if (np) bulk_g2s(st.pat, row_pat + m.pad0, np, &full[s]);
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
            const int s = i % P_STAGES;
            mbar_wait(&full[s], (i / P_STAGES) & 1);
            StageP &st = stages[s];
            const StageMeta m = st.meta;
            const int jcut = m.slow ? 0 : m.pcut - m.pa;
            for (int r = m.r0 + tid; r < m.r1; r += P_CONSUMERS) {
                double wv = 0.0, w2v = 0.0;
                if (DOTS) wv = __ldcs(w + r);
                if (DOTS == 2) w2v = __ldg(w2 + r);
                const int ia = r - m.ra;
                const int a = ((r < m.rcut) ? st.rptr[ia] : __ldg(indptr + r)) - m.pa;
                const int b = ((r + 1 < m.rcut) ? st.rptr[ia + 1] : __ldg(indptr + r + 1)) - m.pa;
                const int pid = (r < m.pad1) ? (int)st.pat[r - m.pad0] : (int)__ldg(row_pat + r);
                const pat_off_t *po = s_tab + pid * P_MAXLEN;
                double sum = 0.0;
                const int bs = b < jcut ? b : jcut;             // staged part of the row
                {
                    // first T_BATCH (= 9) entries: the offsets arrive with two vector loads (a pattern is 32 aligned
                    // bytes) instead of nine scalar ones -- with nine, ptxas interleaves them with the first gathers and
                    // a shared scoreboard makes the remaining gathers wait for a DRAM round trip -- then all gathers are
                    // issued before the first FMA
                    const uint4 o0 = *reinterpret_cast<const uint4 *>(po);
                    const uint32_t o1 = *reinterpret_cast<const uint32_t *>(po + 8);
                    const uint32_t ow[5] = {o0.x, o0.y, o0.z, o0.w, o1};
                    const int nb = bs - a;                       // staged entries of this row (may be <= 0)
                    int c[T_BATCH];
                    double xv[T_BATCH];
#pragma unroll
                    for (int k = 0; k < T_BATCH; ++k) {
                        const int off = (k & 1) ? ((int)ow[k >> 1] >> 16) : (int)(short)(ow[k >> 1] & 0xffffu);
                        c[k] = r + ((k < nb) ? off : 0);
                    }
#pragma unroll
                    for (int k = 0; k < T_BATCH; ++k) xv[k] = ldg_hint(x + c[k], pol_keep);
#pragma unroll
                    for (int k = 0; k < T_BATCH; ++k)
                        if (k < nb) sum = fma(st.vals[a + k], xv[k], sum);
                }
                for (int q = a + T_BATCH; q < bs; ++q)           // rows longer than 9 entries (rare)
                    sum = fma(st.vals[q], __ldg(x + (r + (int)po[q - a])), sum);
                for (int q = (bs > a ? bs : a); q < b; ++q)      // unstaged tail (last block of the matrix only)
                    sum = fma(__ldg(vals + m.pa + q), __ldg(x + (r + (int)po[q - a])), sum);
                y[r] = sum;
                if (DOTS) { d0 += wv * sum; d1 += sum * sum; }
                if (DOTS == 2) { d2 += w2v * sum; d3 += w2v * wv; }
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
                for (int k = tid; k < (int)gridDim.x; k += P_CONSUMERS) {
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

}  // namespace kbspmv
