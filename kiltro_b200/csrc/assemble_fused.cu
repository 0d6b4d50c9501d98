// K1+K3 fused: numeric assembly with element matrices computed on chip and never written to HBM.
//
// Replaces the element loop of FiniteElements/Assembly.py:46-72 (GetElementalMatrices -> scatter-add into CSR data ->
// Flux) for the whole mesh in one launch.  Persistent CTAs (2 per SM) walk the tiles of the plan (tileplan.cuh); per tile:
//   phase A  one thread per incident element: the nodal coordinates / velocities were gathered into shared memory by
//            cp.async (LDGSTS) while the PREVIOUS tile was being processed; the thread evaluates the fp64 2x2-Gauss
//            element matrix in registers (element.cuh) and parks it in shared memory component-major; it then starts
//            the gathers of its element of the NEXT tile into the same staging slots;
//   phase B  one thread per owned row: walks the row's (element, local node, slot positions) records in ascending
//            element order -- the order of the reference's sequential `+=` (Assembly.py:46,60-61) -- and adds the
//            element row into the row's CSR slots held in shared memory (positions precomputed by the symbolic phase;
//            no atomics, no searching, deterministic);
//   phase C  the tile's CSR values leave through coalesced streaming stores (a tile's rows are runs of consecutive
//            rows, so its entries are long contiguous pieces of K.data).
// (A thread-per-entry phase B without the shared read-modify-write was tried: 2.9 ms instead of 1.6 ms at 16 M DOF --
// more instructions in an issue-bound kernel.)
// Row metadata (row ids, pair ranges, first records) of the next tile is fetched one tile ahead too, so no phase waits
// on a dependent global load.  The mass matrix (transient) is a second pass of phase B over the same buffers.
//
// HBM traffic per launch: the mesh (coordinates / velocity through the tile-local connectivity copy, ~1.14x because of
// the tile rim), the plan (4 B per (row, element) pair + 12 B per row) and one write of K.data; `indices` is not read.
#include "element.cuh"
#include "tileplan.cuh"
#include "cpasync.cuh"

using namespace kbel;
using namespace kbtile;
using namespace kbasync;

ElemOpts kb_make_opts(const kb_material *m, int mode, int petrov, int want_mass);   // element.cu

namespace {

template <int NDIM> struct ElemIO;
template <> struct ElemIO<2> {
    static constexpr int NPE = 4;
    static constexpr int STAGE_DOUBLES = 16;          // per element: 4 nodes x (x,y) + 4 nodes x (ux,uy)
    typedef int4 Conn;
    static __device__ __forceinline__ Conn load_conn(const int32_t *tile_conn, int e) {
        return __ldg(reinterpret_cast<const int4 *>(tile_conn) + e);
    }
    // staging layout: chunk-major double2 [8][EMAX] -> consecutive threads touch consecutive 16-byte words
    template <bool HAS_U>
    static __device__ __forceinline__ void gather_async(double *stage, int tid, const double *points,
                                                        const double *velocity, const Conn &c) {
        const int n[4] = {c.x, c.y, c.z, c.w};
        double2 *s2 = reinterpret_cast<double2 *>(stage);
#pragma unroll
        for (int a = 0; a < 4; ++a) {
            cp_async16(s2 + a * EMAX + tid, reinterpret_cast<const double2 *>(points) + n[a]);
            if (HAS_U) cp_async16(s2 + (4 + a) * EMAX + tid, reinterpret_cast<const double2 *>(velocity) + n[a]);
        }
    }
    template <bool HAS_U, bool MASS, bool GENERAL>
    static __device__ __forceinline__ bool eval(const double *stage, int tid, const ElemOpts &o, double (&ke)[16],
                                                double (&fe)[4], double (&me)[16]) {
        const double2 *s2 = reinterpret_cast<const double2 *>(stage);
        double X[4][2], U[4][2];
#pragma unroll
        for (int a = 0; a < 4; ++a) {
            const double2 p = s2[a * EMAX + tid];
            X[a][0] = p.x; X[a][1] = p.y;
            if (HAS_U) { const double2 u = s2[(4 + a) * EMAX + tid]; U[a][0] = u.x; U[a][1] = u.y; }
            else { U[a][0] = 0.0; U[a][1] = 0.0; }
        }
        if (GENERAL) { quad4_general(X, U, o, ke, fe, me); return true; }
        return quad4_fast(X, U, o, ke, fe, me);
    }
};
template <> struct ElemIO<1> {
    static constexpr int NPE = 2;
    static constexpr int STAGE_DOUBLES = 4;
    typedef int2 Conn;
    static __device__ __forceinline__ Conn load_conn(const int32_t *tile_conn, int e) {
        return __ldg(reinterpret_cast<const int2 *>(tile_conn) + e);
    }
    template <bool HAS_U>
    static __device__ __forceinline__ void gather_async(double *stage, int tid, const double *points,
                                                        const double *velocity, const Conn &c) {
        cp_async8(stage + 0 * EMAX + tid, points + c.x);
        cp_async8(stage + 1 * EMAX + tid, points + c.y);
        if (HAS_U) { cp_async8(stage + 2 * EMAX + tid, velocity + c.x); cp_async8(stage + 3 * EMAX + tid, velocity + c.y); }
    }
    template <bool HAS_U, bool MASS, bool GENERAL>
    static __device__ __forceinline__ bool eval(const double *stage, int tid, const ElemOpts &o, double (&ke)[4],
                                                double (&fe)[2], double (&me)[4]) {
        double X[2] = {stage[0 * EMAX + tid], stage[1 * EMAX + tid]};
        double U[2] = {0.0, 0.0};
        if (HAS_U) { U[0] = stage[2 * EMAX + tid]; U[1] = stage[3 * EMAX + tid]; }
        line2(X, U, o, ke, fe, me);
        return true;
    }
};

__device__ __forceinline__ int lower_bound_g(const int32_t *a, int n, int v) {
    int lo = 0, hi = n;
    while (lo < hi) {
        const int mid = (lo + hi) >> 1;
        if (__ldg(a + mid) < v) lo = mid + 1; else hi = mid;
    }
    return lo;
}

// shared-memory carve-up (bytes) -- also used by the launcher
template <int NDIM> struct Smem {
    static constexpr int NPE = ElemIO<NDIM>::NPE, CAP = NPE * NPE;
    static constexpr size_t oKe = 0;
    static constexpr size_t oFe = oKe + 8 * (size_t)EMAX * CAP;
    static constexpr size_t oStage = oFe + 8 * (size_t)EMAX * NPE;
    static constexpr size_t oV = oStage + 8 * (size_t)EMAX * ElemIO<NDIM>::STAGE_DOUBLES;
    static constexpr size_t oDelta = oV + 8 * (size_t)VMAX;
    static constexpr size_t oSegOff = oDelta + 4 * (size_t)RMAX;
    static constexpr size_t oScan = oSegOff + 4 * (size_t)(RMAX + 4);
    static constexpr size_t bytes = oScan + 4 * 64;
};

// GENERAL = false: every tile, with the sum-factorised quad-4 routine; an element it declines (non-convex / inverted /
// extreme scale) flags its tile in bad_tile and raises *nbad.  GENERAL = true: the fix-up launch that follows -- only
// the flagged tiles, general element routine; exits at once when *nbad == 0 (the normal case).
template <int NDIM, bool HAS_U, bool MASS, bool GENERAL>
__global__ void __launch_bounds__(THREADS, 2)
k_assemble_fused(const double *__restrict__ points, const double *__restrict__ velocity, ElemOpts o,
                 const int32_t *__restrict__ indptr, const int32_t *__restrict__ indices,
                 const int32_t *__restrict__ tile_rows, const int32_t *__restrict__ tile_row_ptr,
                 const int32_t *__restrict__ tile_conn, const int32_t *__restrict__ tile_elem_ptr,
                 const uint32_t *__restrict__ pair_rec, const int32_t *__restrict__ pair_ptr, double *__restrict__ K,
                 double *__restrict__ M, double *__restrict__ Flux, int ntiles, uint8_t *bad_tile, int *nbad) {
    constexpr int NPE = ElemIO<NDIM>::NPE, CAP = NPE * NPE;
    typedef Smem<NDIM> L;
    typedef typename ElemIO<NDIM>::Conn Conn;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    double *sKe = reinterpret_cast<double *>(smem_raw + L::oKe);        // [CAP][EMAX] component-major
    double *sFe = reinterpret_cast<double *>(smem_raw + L::oFe);        // [NPE][EMAX]
    double *sStage = reinterpret_cast<double *>(smem_raw + L::oStage);  // cp.async landing zone of the gathers
    double *sV = reinterpret_cast<double *>(smem_raw + L::oV);          // CSR values of the tile's rows
    int *sDelta = reinterpret_cast<int *>(smem_raw + L::oDelta);        // per run: global position - local position
    int *sSegOff = reinterpret_cast<int *>(smem_raw + L::oSegOff);      // local offsets of the runs of consecutive rows
    int *sScan = reinterpret_cast<int *>(smem_raw + L::oScan);

    const int tid = threadIdx.x;
    const int G = gridDim.x;
    if (GENERAL && *reinterpret_cast<volatile int *>(nbad) == 0) return;
    auto next_tile = [&](int t) {
        t += G;
        if (GENERAL) while (t < ntiles && !bad_tile[t]) t += G;
        return t;
    };
    int t = next_tile((int)blockIdx.x - G);
    if (t >= ntiles) return;

    // ---- prologue: everything the first tile needs
    int rb = __ldg(tile_row_ptr + t), nr = __ldg(tile_row_ptr + t + 1) - rb;
    int eb = __ldg(tile_elem_ptr + t), ne = __ldg(tile_elem_ptr + t + 1) - eb;
    if (tid < ne) {
        const Conn c = ElemIO<NDIM>::load_conn(tile_conn, eb + tid);
        ElemIO<NDIM>::template gather_async<HAS_U>(sStage, tid, points, velocity, c);
    }
    cp_async_commit();
    int row = 0, rowp = -2, jp0 = 0, jp1 = 0;
    if (tid < nr) {
        row = __ldg(tile_rows + rb + tid);
        if (tid > 0) rowp = __ldg(tile_rows + rb + tid - 1);
        jp0 = __ldg(pair_ptr + rb + tid);
        jp1 = __ldg(pair_ptr + rb + tid + 1);
    }

    for (; t < ntiles;) {
        const int tn = next_tile(t);
        // extents of the next tile (consumed after phase A)
        int rb_n = 0, nr_n = 0, eb_n = 0, ne_n = 0;
        if (tn < ntiles) {
            rb_n = __ldg(tile_row_ptr + tn); nr_n = __ldg(tile_row_ptr + tn + 1) - rb_n;
            eb_n = __ldg(tile_elem_ptr + tn); ne_n = __ldg(tile_elem_ptr + tn + 1) - eb_n;
        }
        // second-level row metadata of THIS tile: addresses were fetched a tile ago, results are used after phase A
        int p0 = 0, p1 = 0;
        uint4 rec4 = make_uint4(0u, 0u, 0u, 0u);
        if (tid < nr) {
            p0 = __ldg(indptr + row);
            p1 = __ldg(indptr + row + 1);
            const int np = jp1 - jp0;
            if (np > 0) rec4.x = __ldg(pair_rec + jp0);
            if (np > 1) rec4.y = __ldg(pair_rec + jp0 + 1);
            if (np > 2) rec4.z = __ldg(pair_rec + jp0 + 2);
            if (np > 3) rec4.w = __ldg(pair_rec + jp0 + 3);
        }

        // ---- phase A: element matrices -> shared memory, component-major (conflict-free stores and reads)
        cp_async_wait_all();                              // this thread's gathers (issued one tile ago) have landed
        double me[CAP];
        if (tid < ne) {
            double ke[CAP], fe[NPE];
            if (!ElemIO<NDIM>::template eval<HAS_U, MASS, GENERAL>(sStage, tid, o, ke, fe, me)) {
                bad_tile[t] = 1;                          // redone by the fix-up launch
                *nbad = 1;
            }
#pragma unroll
            for (int i = 0; i < CAP; ++i) sKe[i * EMAX + tid] = ke[i];
#pragma unroll
            for (int i = 0; i < NPE; ++i) sFe[i * EMAX + tid] = fe[i];
        }
        // the staging slots of this thread are free again: start the gathers of its element of the next tile
        if (tn < ntiles && tid < ne_n) {
            const Conn c = ElemIO<NDIM>::load_conn(tile_conn, eb_n + tid);
            ElemIO<NDIM>::template gather_async<HAS_U>(sStage, tid, points, velocity, c);
        }
        cp_async_commit();
        // first-level row metadata of the next tile
        int row_n = 0, rowp_n = -2, jp0_n = 0, jp1_n = 0;
        if (tn < ntiles && tid < nr_n) {
            row_n = __ldg(tile_rows + rb_n + tid);
            if (tid > 0) rowp_n = __ldg(tile_rows + rb_n + tid - 1);
            jp0_n = __ldg(pair_ptr + rb_n + tid);
            jp1_n = __ldg(pair_ptr + rb_n + tid + 1);
        }

        // ---- row bookkeeping: one block scan gives the local offset of every row's entries (low 16 bits) and the index
        // of the run of consecutive rows it belongs to (high bits); a run's entries are contiguous in K.data.
        const int len = tid < nr ? p1 - p0 : 0;
        const int segstart = (tid < nr && row != rowp + 1) ? 1 : 0;
        int packed_total;
        const int ex = block_scan_excl(len | (segstart << 16), sScan, &packed_total);
        const int off = ex & 0xffff, total = packed_total & 0xffff, nseg = packed_total >> 16;
        if (segstart) { sSegOff[ex >> 16] = off; sDelta[ex >> 16] = p0 - off; }
        if (tid == 0) sSegOff[nseg] = total;
        for (int i = tid; i < total; i += THREADS) sV[i] = 0.0;
        __syncthreads();

#pragma unroll 1
        for (int pass = 0; pass < (MASS ? 2 : 1); ++pass) {
            // ---- phase B: row-owner accumulation in ascending element order
            if (tid < nr) {
                double f = 0.0;
                for (int j = jp0; j < jp1; ++j) {
                    const int k4 = j - jp0;
                    const uint32_t rec = k4 == 0 ? rec4.x : k4 == 1 ? rec4.y : k4 == 2 ? rec4.z : k4 == 3 ? rec4.w
                                                                                               : __ldg(pair_rec + j);
                    const int slot = (int)((rec >> 18) & 0x1fffu), a = (int)((rec >> 16) & 3u);
                    const double *src = &sKe[(a * NPE) * EMAX + slot];
                    if (!(rec & 0x80000000u)) {
#pragma unroll
                        for (int b = 0; b < NPE; ++b) sV[off + (int)((rec >> (4 * b)) & 15u)] += src[b * EMAX];
                    } else {                              // row longer than 16 entries: locate the columns by search
#pragma unroll
                        for (int b = 0; b < NPE; ++b)
                            sV[off + lower_bound_g(indices + p0, len, __ldg(tile_conn + (int64_t)(eb + slot) * NPE + b))] +=
                                src[b * EMAX];
                    }
                    if (pass == 0) f += sFe[a * EMAX + slot];         // source vector                Assembly.py:69-72
                }
                if (pass == 0) Flux[row] = f;
            }
            __syncthreads();
            // ---- phase C: coalesced write-out, one warp per run of consecutive rows
            double *out = pass == 0 ? K : M;
            for (int sg = tid >> 5; sg < nseg; sg += THREADS / 32) {
                const int a0 = sSegOff[sg], b0 = sSegOff[sg + 1], dl = sDelta[sg];
                for (int i = a0 + (tid & 31); i < b0; i += 32) __stcs(out + i + dl, sV[i]);
            }
            if (MASS && pass == 0) {
                __syncthreads();
                if (tid < ne) {
#pragma unroll
                    for (int i = 0; i < CAP; ++i) sKe[i * EMAX + tid] = me[i];
                }
                for (int i = tid; i < total; i += THREADS) sV[i] = 0.0;
                __syncthreads();
            }
        }
        __syncthreads();                                  // the next tile overwrites the shared buffers
        t = tn;
        rb = rb_n; nr = nr_n; eb = eb_n; ne = ne_n;
        row = row_n; rowp = rowp_n; jp0 = jp0_n; jp1 = jp1_n;
    }
}

template <int NDIM, bool HAS_U, bool MASS>
int launch_fused(int64_t ntiles, cudaStream_t s, const double *points, const double *velocity, const ElemOpts &o,
                 const int32_t *indptr, const int32_t *indices, const int32_t *tile_rows, const int32_t *tile_row_ptr,
                 const int32_t *tile_conn, const int32_t *tile_elem_ptr, const uint32_t *pair_rec,
                 const int32_t *pair_ptr, double *K, double *M, double *Flux, uint8_t *scratch) {
    static KbOncePerDevice configured;
    constexpr size_t bytes = Smem<NDIM>::bytes;
    if (configured.need()) {
        KB_CUDA(cudaFuncSetAttribute(k_assemble_fused<NDIM, HAS_U, MASS, false>,
                                     cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes));
        KB_CUDA(cudaFuncSetAttribute(k_assemble_fused<NDIM, HAS_U, MASS, true>,
                                     cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes));
    }
    const int64_t cap = (int64_t)kb_num_sms() * 2;           // persistent: 2 CTAs per SM
    const int grid = (int)(ntiles < cap ? ntiles : cap);
    const size_t flag_bytes = ((size_t)ntiles + 15) / 16 * 16;
    int *nbad = reinterpret_cast<int *>(scratch + flag_bytes);
    if (NDIM == 2) KB_CUDA(cudaMemsetAsync(scratch, 0, flag_bytes + 16, s));
    k_assemble_fused<NDIM, HAS_U, MASS, false><<<grid, THREADS, bytes, s>>>(
        points, velocity, o, indptr, indices, tile_rows, tile_row_ptr, tile_conn, tile_elem_ptr, pair_rec, pair_ptr, K, M,
        Flux, (int)ntiles, scratch, nbad);
    KB_LAUNCH_CHECK();
    if (NDIM == 2) {                                          // line elements never decline
        k_assemble_fused<NDIM, HAS_U, MASS, true><<<grid, THREADS, bytes, s>>>(
            points, velocity, o, indptr, indices, tile_rows, tile_row_ptr, tile_conn, tile_elem_ptr, pair_rec, pair_ptr,
            K, M, Flux, (int)ntiles, scratch, nbad);
        KB_LAUNCH_CHECK();
    }
    return 0;
}

}  // namespace

extern "C" int kb_assemble_fused(int ndim, const double *d_points, const int32_t *d_elements, int64_t nelem,
                                 int64_t nnode, const double *d_velocity, const kb_material *material, int mode,
                                 int petrov, const int32_t *d_indptr, const int32_t *d_indices, int64_t ntiles,
                                 const int32_t *d_tile_rows, const int32_t *d_tile_row_ptr,
                                 const int32_t *d_tile_conn, const int32_t *d_tile_elem_ptr,
                                 const uint32_t *d_pair_rec, const int32_t *d_pair_ptr, double *d_K, double *d_M,
                                 double *d_Flux, void *d_scratch, void *stream) {
    if (!d_points || !d_elements || !material || !d_indptr || !d_indices || !d_tile_rows || !d_tile_row_ptr ||
        !d_tile_conn || !d_tile_elem_ptr || !d_pair_rec || !d_pair_ptr || !d_K || !d_Flux || !d_scratch)
        return KB_EINVAL;
    if ((ndim != 1 && ndim != 2) || ntiles < 0 || nelem < 0 || nnode < 0) return KB_EINVAL;
    if (mode != KB_MODE_REFERENCE && mode != KB_MODE_CONSISTENT) return KB_EINVAL;
    if (ntiles == 0) return 0;
    cudaStream_t s = (cudaStream_t)stream;
    const ElemOpts o = kb_make_opts(material, mode, petrov, d_M != nullptr);
#define KB_FUSED(ND, HU, MS)                                                                                      \
    return launch_fused<ND, HU, MS>(ntiles, s, d_points, d_velocity, o, d_indptr, d_indices, d_tile_rows,            \
                                    d_tile_row_ptr, d_tile_conn, d_tile_elem_ptr, d_pair_rec, d_pair_ptr, d_K, d_M,    \
                                    d_Flux, (uint8_t *)d_scratch)
    const bool hu = d_velocity != nullptr, ms = d_M != nullptr;
    if (ndim == 2) {
        if (hu && ms) KB_FUSED(2, true, true); else if (hu) KB_FUSED(2, true, false);
        else if (ms) KB_FUSED(2, false, true); else KB_FUSED(2, false, false);
    } else {
        if (hu && ms) KB_FUSED(1, true, true); else if (hu) KB_FUSED(1, true, false);
        else if (ms) KB_FUSED(1, false, true); else KB_FUSED(1, false, false);
    }
#undef KB_FUSED
}
