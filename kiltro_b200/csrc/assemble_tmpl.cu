// K1+K3 fused, template form (quad-4): numeric assembly for meshes whose tiles repeat (tileplan.cuh, tiletmpl.cu).
//
// Same contract as k_assemble_fused (assemble_fused.cu; replaces the element loop of FiniteElements/Assembly.py:46-72),
// same tiles, same summation order -- but the per-(row, element) records, row ids and row pointers of a tile are not
// read from HBM: the CTA keeps the tile's TEMPLATE in shared memory (reloaded only when the class of the next tile
// differs) and fetches, per tile, the tile-local connectivity and one (first row, indptr) pair per run of rows.
//   phase A   one thread per incident element: nodal data were gathered into shared memory by cp.async one tile ahead;
//             fp64 element matrix in registers (element.cuh, sum-factorised) -> shared memory, component-major with an
//             odd stride;
//   phase BC  one thread per CSR entry of the tile: adds the entry's <= 4 contributions in ascending element order
//             (byte offsets from the template; unused ones point at a zero word, so there is no branch) and streams the
//             value straight to K.data -- consecutive threads write consecutive addresses inside a run of rows.
// No atomics, deterministic; K equals the generic kernel's K bit for bit.
#include "element.cuh"
#include "tileplan.cuh"
#include "cpasync.cuh"

using namespace kbel;
using namespace kbtile;
using namespace kbasync;

ElemOpts kb_make_opts(const kb_material *m, int mode, int petrov, int want_mass);   // element.cu

namespace {

__device__ __forceinline__ double lds_f64(uint32_t addr) {
    double v;
    asm volatile("ld.shared.f64 %0, [%1];" : "=d"(v) : "r"(addr));
    return v;
}

// split barrier (mbarrier, one arrival per warp): "every warp has finished reading the shared buffers of the previous
// tile".  Arrive right after phase BC, wait only just before the element matrices of the next tile are stored -- the
// whole fp64 element computation of a warp overlaps the phase BC of the slower warps.
__device__ __forceinline__ void mbar_init(uint64_t *b, int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_addr(b)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t *b) {
    asm volatile("{\n\t.reg .b64 st;\n\tmbarrier.arrive.shared::cta.b64 st, [%0];\n\t}" ::"r"(smem_addr(b)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *b, uint32_t parity) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "WAIT_%=:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@!p bra WAIT_%=;\n\t}" ::"r"(smem_addr(b)), "r"(parity) : "memory");
}

struct SmemT {
    double ke[16 * KSTRIDE + 1];               // [component][slot], + the zero word at KZERO
    double fe[4 * FSTRIDE + 1];                // + zero word at FZERO
    double2 stage[8 * EMAX];                   // cp.async landing zone: 4 x coordinates, 4 x velocities per element
    TileTmpl tmpl;
    int delta[RUNMAX + 1];                     // per run: global entry position - local entry position
    int rdelta[RUNMAX + 1];                    // per run: global row - local row
    uint64_t bar_free;                         // split barrier: shared buffers of the previous tile are free
};

template <bool HAS_U>
__device__ __forceinline__ void gather_async(double2 *stage, int tid, const double *points, const double *velocity,
                                             const int4 &c) {
    const int n[4] = {c.x, c.y, c.z, c.w};
#pragma unroll
    for (int a = 0; a < 4; ++a) {
        cp_async16(stage + a * EMAX + tid, reinterpret_cast<const double2 *>(points) + n[a]);
        if (HAS_U) cp_async16(stage + (4 + a) * EMAX + tid, reinterpret_cast<const double2 *>(velocity) + n[a]);
    }
}

// loads whose issue point matters (metadata fetched one or two tiles ahead): volatile, so the compiler keeps them where
// they are written instead of sinking them to their first use
__device__ __forceinline__ int ldg_pin(const int32_t *p) {
    int v;
    asm volatile("ld.global.nc.s32 %0, [%1];" : "=r"(v) : "l"(p));
    return v;
}
__device__ __forceinline__ int ldg_pin(const uint8_t *p) {
    int v;
    asm volatile("ld.global.nc.u8 %0, [%1];" : "=r"(v) : "l"(p));
    return v;
}
__device__ __forceinline__ int4 ldg_pin4(const int32_t *p) {
    int4 v;
    asm volatile("ld.global.nc.v4.s32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p));
    return v;
}

// GENERAL = false: every tile, sum-factorised element routine; an element it declines (non-convex / inverted / extreme
//                  scale: quad4_fast returns false) flags its tile in bad_tile and raises *nbad;
// GENERAL = true:  the fix-up launch that follows -- only the flagged tiles, general element routine for all their
//                  elements, same summation order; exits at once when *nbad == 0 (the normal case).
template <bool HAS_U, bool MASS, bool GENERAL>
__global__ void __launch_bounds__(THREADS, 2)
k_assemble_tmpl(const double *__restrict__ points, const double *__restrict__ velocity, ElemOpts o,
                const int32_t *__restrict__ indptr, const int32_t *__restrict__ tile_conn,
                const int32_t *__restrict__ tile_elem_ptr, const uint8_t *__restrict__ tile_class,
                const int32_t *__restrict__ trun_ptr, const int32_t *__restrict__ trun_row,
                const TileTmpl *__restrict__ templates, double *__restrict__ K, double *__restrict__ M,
                double *__restrict__ Flux, int ntiles, uint8_t *bad_tile, int *nbad) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    SmemT &sm = *reinterpret_cast<SmemT *>(smem_raw);
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
    if (tid == 0) { sm.ke[KZERO] = 0.0; sm.fe[FZERO] = 0.0; mbar_init(&sm.bar_free, THREADS / 32); }
    bool pending = false;                                   // an arrival round of bar_free not yet waited for
    uint32_t parity = 0;

    // ---- prologue: tile t completely, the extents of the tile after it
    int eb = __ldg(tile_elem_ptr + t), ne = __ldg(tile_elem_ptr + t + 1) - eb;
    if (tid < ne) gather_async<HAS_U>(sm.stage, tid, points, velocity, __ldg(reinterpret_cast<const int4 *>(tile_conn) + eb + tid));
    cp_async_commit();
    int cls = __ldg(tile_class + t), cur = -1;
    int frow = 0;
    {
        const int rp = __ldg(trun_ptr + t), nrun = __ldg(trun_ptr + t + 1) - rp;
        frow = __ldg(trun_row + rp + min(tid, max(nrun - 1, 0)));
    }
    int tn = next_tile(t);
    int eb_n = 0, ne_n = 0, cls_n = 0, rp_n = 0, nrun_n = 0;
    if (tn < ntiles) {
        eb_n = __ldg(tile_elem_ptr + tn); ne_n = __ldg(tile_elem_ptr + tn + 1) - eb_n;
        cls_n = __ldg(tile_class + tn);
        rp_n = __ldg(trun_ptr + tn); nrun_n = __ldg(trun_ptr + tn + 1) - rp_n;
    }
    const uint32_t ke_base = smem_addr(sm.ke), fe_base = smem_addr(sm.fe);
    if (fe_base + (uint32_t)sizeof(sm.fe) > 0xffffu) __trap();            // source addresses are kept in 16 bits
    const bool has_q = (o.area * o.q) != 0.0;

    while (t < ntiles) {
        // ---- issue now what is consumed after phase A: the next tile's connectivity and run rows (their addresses were
        // fetched a tile ago), this tile's row pointers, and the extents of the tile after the next
        const int tnn = tn < ntiles ? next_tile(tn) : tn;
        // (unconditional, index clamped: a predicated load plus a zero in the other branch makes the zeroing wait for
        // the load -- same destination registers)
        const int4 conn_n = ldg_pin4(tile_conn + 4 * (int64_t)(eb_n + min(tid, max(ne_n - 1, 0))));
        const int frow_n = ldg_pin(trun_row + rp_n + min(tid, max(nrun_n - 1, 0)));
        int eb_nn = 0, ne_nn = 0, cls_nn = 0, rp_nn = 0, nrun_nn = 0;
        if (tnn < ntiles) {
            eb_nn = ldg_pin(tile_elem_ptr + tnn); ne_nn = ldg_pin(tile_elem_ptr + tnn + 1);
            cls_nn = ldg_pin(tile_class + tnn);
            rp_nn = ldg_pin(trun_ptr + tnn); nrun_nn = ldg_pin(trun_ptr + tnn + 1);
        }
        if (cls != cur) {                                   // uniform: (re)load the template of this tile's class
            if (pending) { mbar_wait(&sm.bar_free, parity); parity ^= 1u; pending = false; }
            __syncthreads();                                // (also publishes the barrier init on the first tile)
            const uint4 *src = reinterpret_cast<const uint4 *>(templates + cls);
            uint4 *dst = reinterpret_cast<uint4 *>(&sm.tmpl);
            for (int i = tid; i < (int)(sizeof(TileTmpl) / 16); i += THREADS) dst[i] = __ldg(src + i);
            cur = cls;
            __syncthreads();
            // turn the byte offsets of the shared copy into shared-window addresses (they still fit 16 bits)
            {
                uint32_t *e = reinterpret_cast<uint32_t *>(&sm.tmpl.ent_src[0][0]);
                for (int i = tid; i < VMAX * 2; i += THREADS) e[i] += ke_base * 0x10001u;
                uint32_t *r = reinterpret_cast<uint32_t *>(&sm.tmpl.row_src[0][0]);
                for (int i = tid; i < RMAX * 2; i += THREADS) r[i] += fe_base * 0x10001u;
            }
            __syncthreads();
        }
        const int nrun = sm.tmpl.nrun, nent = sm.tmpl.nent, nrow = sm.tmpl.nrow;
        const int ip = ldg_pin(indptr + frow);              // frow is a valid row (or 0) in every thread

        // ---- phase A: element matrices -> shared memory
        cp_async_wait_all();                                // this thread's gathers (issued one tile ago) have landed
        double me[16];
        if (tid < ne) {
            double X[4][2], U[4][2], ke[16], fe[4];
#pragma unroll
            for (int a = 0; a < 4; ++a) {
                const double2 p = sm.stage[a * EMAX + tid];
                X[a][0] = p.x; X[a][1] = p.y;
                if (HAS_U) { const double2 u = sm.stage[(4 + a) * EMAX + tid]; U[a][0] = u.x; U[a][1] = u.y; }
                else { U[a][0] = 0.0; U[a][1] = 0.0; }
            }
            if (GENERAL) {
                quad4_general(X, U, o, ke, fe, me);
            } else if (!quad4_fast(X, U, o, ke, fe, me)) {
                bad_tile[t] = 1;                            // redone by the fix-up launch
                *nbad = 1;
            }
            if (pending) mbar_wait(&sm.bar_free, parity);   // the previous tile's phase BC is over in every warp
#pragma unroll
            for (int i = 0; i < 16; ++i) sm.ke[i * KSTRIDE + tid] = ke[i];
            if (has_q) {
#pragma unroll
                for (int i = 0; i < 4; ++i) sm.fe[i * FSTRIDE + tid] = fe[i];
            }
        } else if (pending) {
            mbar_wait(&sm.bar_free, parity);
        }
        if (pending) { parity ^= 1u; pending = false; }
        // the staging slots of this thread are free again: start the gathers of its element of the next tile
        if (tid < ne_n) gather_async<HAS_U>(sm.stage, tid, points, velocity, conn_n);
        cp_async_commit();
        if (tid < nrun) { sm.delta[tid] = ip - (int)sm.tmpl.run_ent[tid]; sm.rdelta[tid] = frow - (int)sm.tmpl.run_row[tid]; }
        __syncthreads();

#pragma unroll 1
        for (int pass = 0; pass < (MASS ? 2 : 1); ++pass) {
            // ---- phase BC: one thread per CSR entry
            double *out = pass == 0 ? K : M;
#pragma unroll 4
            for (int i = tid; i < nent; i += THREADS) {
                const uint2 sr = *reinterpret_cast<const uint2 *>(&sm.tmpl.ent_src[i][0]);
                const int dl = sm.delta[sm.tmpl.ent_run[i]];
                double v = lds_f64(sr.x & 0xffffu);
                v += lds_f64(sr.x >> 16);
                v += lds_f64(sr.y & 0xffffu);
                v += lds_f64(sr.y >> 16);
                __stcs(out + (int64_t)(i + dl), v);
            }
            if (pass == 0 && tid < nrow) {                  // source vector                         Assembly.py:69-72
                double f = 0.0;
                if (has_q) {
                    const uint2 sr = *reinterpret_cast<const uint2 *>(&sm.tmpl.row_src[tid][0]);
                    f = lds_f64(sr.x & 0xffffu);
                    f += lds_f64(sr.x >> 16);
                    f += lds_f64(sr.y & 0xffffu);
                    f += lds_f64(sr.y >> 16);
                }
                Flux[tid + sm.rdelta[sm.tmpl.row_run[tid]]] = f;
            }
            if (MASS && pass == 0) {
                __syncthreads();
                if (tid < ne) {
#pragma unroll
                    for (int i = 0; i < 16; ++i) sm.ke[i * KSTRIDE + tid] = me[i];
                }
                __syncthreads();
            }
        }
        __syncwarp();
        if ((tid & 31) == 0) mbar_arrive(&sm.bar_free);     // this warp no longer reads the shared buffers of tile t
        pending = true;
        t = tn; tn = tnn;
        eb = eb_n; ne = ne_n; cls = cls_n; frow = frow_n;
        eb_n = eb_nn; ne_n = ne_nn - eb_nn; cls_n = cls_nn; rp_n = rp_nn; nrun_n = nrun_nn - rp_nn;
    }
}

template <bool HAS_U, bool MASS>
int launch_tmpl(int64_t ntiles, cudaStream_t s, const double *points, const double *velocity, const ElemOpts &o,
                const int32_t *indptr, const int32_t *tile_conn, const int32_t *tile_elem_ptr,
                const uint8_t *tile_class, const int32_t *trun_ptr, const int32_t *trun_row, const void *templates,
                double *K, double *M, double *Flux, uint8_t *scratch) {
    static KbOncePerDevice configured;
    constexpr size_t bytes = sizeof(SmemT);
    if (configured.need()) {
        KB_CUDA(cudaFuncSetAttribute(k_assemble_tmpl<HAS_U, MASS, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes));
        KB_CUDA(cudaFuncSetAttribute(k_assemble_tmpl<HAS_U, MASS, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes));
    }
    const int64_t cap = (int64_t)kb_num_sms() * 2;           // persistent: 2 CTAs per SM
    const int grid = (int)(ntiles < cap ? ntiles : cap);
    const size_t flag_bytes = ((size_t)ntiles + 15) / 16 * 16;
    int *nbad = reinterpret_cast<int *>(scratch + flag_bytes);
    KB_CUDA(cudaMemsetAsync(scratch, 0, flag_bytes + 16, s));
    k_assemble_tmpl<HAS_U, MASS, false><<<grid, THREADS, bytes, s>>>(points, velocity, o, indptr, tile_conn, tile_elem_ptr,
                                                                     tile_class, trun_ptr, trun_row,
                                                                     (const TileTmpl *)templates, K, M, Flux,
                                                                     (int)ntiles, scratch, nbad);
    KB_LAUNCH_CHECK();
    k_assemble_tmpl<HAS_U, MASS, true><<<grid, THREADS, bytes, s>>>(points, velocity, o, indptr, tile_conn, tile_elem_ptr,
                                                                    tile_class, trun_ptr, trun_row,
                                                                    (const TileTmpl *)templates, K, M, Flux,
                                                                    (int)ntiles, scratch, nbad);
    KB_LAUNCH_CHECK();
    return 0;
}

}  // namespace

extern "C" int kb_assemble_fused_tmpl(const double *d_points, int64_t nnode, const double *d_velocity,
                                      const kb_material *material, int mode, int petrov, const int32_t *d_indptr,
                                      int64_t ntiles, const int32_t *d_tile_conn, const int32_t *d_tile_elem_ptr,
                                      const uint8_t *d_tile_class, const int32_t *d_trun_ptr,
                                      const int32_t *d_trun_row, const void *d_templates, double *d_K, double *d_M,
                                      double *d_Flux, void *d_scratch, void *stream) {
    if (!d_points || !material || !d_indptr || !d_tile_conn || !d_tile_elem_ptr || !d_tile_class || !d_trun_ptr ||
        !d_trun_row || !d_templates || !d_K || !d_Flux || !d_scratch)
        return KB_EINVAL;
    if (ntiles < 0 || nnode < 0) return KB_EINVAL;
    if (mode != KB_MODE_REFERENCE && mode != KB_MODE_CONSISTENT) return KB_EINVAL;
    if (ntiles == 0) return 0;
    cudaStream_t s = (cudaStream_t)stream;
    const ElemOpts o = kb_make_opts(material, mode, petrov, d_M != nullptr);
#define KB_TMPL(HU, MS)                                                                                               \
    return launch_tmpl<HU, MS>(ntiles, s, d_points, d_velocity, o, d_indptr, d_tile_conn, d_tile_elem_ptr, d_tile_class, \
                               d_trun_ptr, d_trun_row, d_templates, d_K, d_M, d_Flux, (uint8_t *)d_scratch)
    const bool hu = d_velocity != nullptr, ms = d_M != nullptr;
    if (hu && ms) KB_TMPL(true, true); else if (hu) KB_TMPL(true, false);
    else if (ms) KB_TMPL(false, true); else KB_TMPL(false, false);
#undef KB_TMPL
}
