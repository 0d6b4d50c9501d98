// K1+K3 fused: numeric assembly with element matrices computed on chip and never written to HBM.
//
// Replaces the element loop of FiniteElements/Assembly.py:46-72 (GetElementalMatrices -> scatter-add into CSR data ->
// Flux) for the whole mesh in one launch.  One CTA per tile of the plan (tileplan.cuh):
//   phase A  one thread per incident element: gather connectivity / coordinates / velocity (vector loads, L1-shared
//            between neighbours), evaluate the fp64 2x2-Gauss element matrix in registers (element.cuh), park it in
//            shared memory;
//   phase B  one thread per owned row: walk the row's (element, local node) pairs in ascending element order -- the
//            order of the reference's sequential `+=` -- and add the element row into the row's CSR slots held in
//            shared memory (positions precomputed by the symbolic phase; no atomics, no searching, deterministic);
//   phase C  the tile's CSR values leave through coalesced streaming stores (a tile's rows are runs of consecutive
//            rows, so its entries are long contiguous pieces of K.data).
// The mass matrix (transient) reuses the same shared buffers in a second pass over phases B and C.
//
// HBM traffic per launch is the mesh (connectivity, coordinates, velocity: read ~1.14x because of the tile rim), the
// plan (4 B per (row, element) pair + 12 B per row) and one write of K.data; column indices are not read at all.
// The kernel is co-limited by the FP64 pipe (~400 DFMA per element) and HBM; see DESIGN.md for the roofline.
#include "element.cuh"
#include "tileplan.cuh"

using namespace kbel;
using namespace kbtile;

ElemOpts kb_make_opts(const kb_material *m, int mode, int petrov, int want_mass);   // element.cu

namespace {

__device__ __forceinline__ void prefetch_l1(const void *p) { asm volatile("prefetch.global.L1 [%0];" ::"l"(p)); }

template <int NDIM> struct ElemIO;
template <> struct ElemIO<2> {
    static constexpr int NPE = 4;
    typedef int4 Conn;
    static __device__ __forceinline__ Conn load_conn(const int32_t *tile_conn, int e) {
        return __ldg(reinterpret_cast<const int4 *>(tile_conn) + e);
    }
    template <bool HAS_U>
    static __device__ __forceinline__ void prefetch(const double *points, const double *velocity, const Conn &c) {
        const int n[4] = {c.x, c.y, c.z, c.w};
#pragma unroll
        for (int a = 0; a < 4; ++a) {
            prefetch_l1(reinterpret_cast<const double2 *>(points) + n[a]);
            if (HAS_U) prefetch_l1(reinterpret_cast<const double2 *>(velocity) + n[a]);
        }
    }
    template <bool HAS_U, bool MASS>
    static __device__ __forceinline__ void eval(const double *points, const double *velocity, const Conn &c,
                                                const ElemOpts &o, double (&ke)[16], double (&fe)[4],
                                                double (&me)[16]) {
        const int n[4] = {c.x, c.y, c.z, c.w};
        double X[4][2], U[4][2];
#pragma unroll
        for (int a = 0; a < 4; ++a) {
            const double2 p = __ldg(reinterpret_cast<const double2 *>(points) + n[a]);
            X[a][0] = p.x; X[a][1] = p.y;
            if (HAS_U) {
                const double2 u = __ldg(reinterpret_cast<const double2 *>(velocity) + n[a]);
                U[a][0] = u.x; U[a][1] = u.y;
            } else { U[a][0] = 0.0; U[a][1] = 0.0; }
        }
        quad4(X, U, o, ke, fe, me);
    }
};
template <> struct ElemIO<1> {
    static constexpr int NPE = 2;
    typedef int2 Conn;
    static __device__ __forceinline__ Conn load_conn(const int32_t *tile_conn, int e) {
        return __ldg(reinterpret_cast<const int2 *>(tile_conn) + e);
    }
    template <bool HAS_U>
    static __device__ __forceinline__ void prefetch(const double *points, const double *velocity, const Conn &c) {
        prefetch_l1(points + c.x); prefetch_l1(points + c.y);
        if (HAS_U) { prefetch_l1(velocity + c.x); prefetch_l1(velocity + c.y); }
    }
    template <bool HAS_U, bool MASS>
    static __device__ __forceinline__ void eval(const double *points, const double *velocity, const Conn &c,
                                                const ElemOpts &o, double (&ke)[4], double (&fe)[2], double (&me)[4]) {
        double X[2] = {__ldg(points + c.x), __ldg(points + c.y)};
        double U[2] = {0.0, 0.0};
        if (HAS_U) { U[0] = __ldg(velocity + c.x); U[1] = __ldg(velocity + c.y); }
        line2(X, U, o, ke, fe, me);
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

template <int NDIM, bool HAS_U, bool MASS>
__global__ void __launch_bounds__(THREADS, 2)
k_assemble_fused(const double *__restrict__ points, const double *__restrict__ velocity, ElemOpts o,
                 const int32_t *__restrict__ indptr, const int32_t *__restrict__ indices,
                 const int32_t *__restrict__ tile_rows, const int32_t *__restrict__ tile_row_ptr,
                 const int32_t *__restrict__ tile_conn,
                 const int32_t *__restrict__ tile_elem_ptr, const uint32_t *__restrict__ pair_rec,
                 const int32_t *__restrict__ pair_ptr, double *__restrict__ K, double *__restrict__ M,
                 double *__restrict__ Flux, int ntiles) {
    constexpr int NPE = ElemIO<NDIM>::NPE, CAP = NPE * NPE;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    double *sKe = reinterpret_cast<double *>(smem_raw);                 // [EMAX * CAP]
    double *sFe = sKe + EMAX * CAP;                                     // [EMAX * NPE]
    double *sV = sFe + EMAX * NPE;                                      // [VMAX]
    unsigned short *sE2R = reinterpret_cast<unsigned short *>(sV + VMAX);   // [VMAX]
    int *sDelta = reinterpret_cast<int *>(sE2R + VMAX);                 // [RMAX]
    int *sScan = sDelta + RMAX;                                         // [64]

    typedef typename ElemIO<NDIM>::Conn Conn;
    const int tid = threadIdx.x;
    int t = blockIdx.x;
    if (t >= ntiles) return;
    // Persistent CTA: tiles t, t+grid, t+2*grid, ...  The next tile's extents and this thread's next element
    // connectivity are fetched one tile ahead and the node data they point to are pulled into L1 (prefetch.global.L1),
    // so the element phase of the next tile starts on L1 hits instead of a three-deep chain of DRAM round trips.
    int rb = __ldg(tile_row_ptr + t), nr = __ldg(tile_row_ptr + t + 1) - rb;
    int eb = __ldg(tile_elem_ptr + t), ne = __ldg(tile_elem_ptr + t + 1) - eb;
    Conn conn = ElemIO<NDIM>::load_conn(tile_conn, eb + (tid < ne ? tid : 0));

    for (; t < ntiles; t += gridDim.x) {
        const int tn = t + gridDim.x;
        int rb_n = 0, nr_n = 0, eb_n = 0, ne_n = 0;
        if (tn < ntiles) {
            rb_n = __ldg(tile_row_ptr + tn); nr_n = __ldg(tile_row_ptr + tn + 1) - rb_n;
            eb_n = __ldg(tile_elem_ptr + tn); ne_n = __ldg(tile_elem_ptr + tn + 1) - eb_n;
        }
        // ---- row metadata first: these loads overlap the element phase
        int row = -1, p0 = 0, len = 0, jp0 = 0, jp1 = 0;
        if (tid < nr) {
            row = __ldg(tile_rows + rb + tid);
            jp0 = __ldg(pair_ptr + rb + tid);
            jp1 = __ldg(pair_ptr + rb + tid + 1);
            p0 = __ldg(indptr + row);
            len = __ldg(indptr + row + 1) - p0;
        }
        uint32_t rec4[4] = {0u, 0u, 0u, 0u};             // the first four (row, element) records (a quad-4 node has <= 4)
#pragma unroll
        for (int k = 0; k < 4; ++k)
            if (jp0 + k < jp1) rec4[k] = __ldg(pair_rec + jp0 + k);

        // ---- phase A: element matrices -> shared memory, component-major (sKe[i][slot]: conflict-free stores and reads)
        double me[CAP];
        if (tid < ne) {
            double ke[CAP], fe[NPE];
            ElemIO<NDIM>::template eval<HAS_U, MASS>(points, velocity, conn, o, ke, fe, me);
#pragma unroll
            for (int i = 0; i < CAP; ++i) sKe[i * EMAX + tid] = ke[i];
#pragma unroll
            for (int i = 0; i < NPE; ++i) sFe[i * EMAX + tid] = fe[i];
        }
        // connectivity of this thread's element in the NEXT tile (consumed after phase B)
        Conn conn_n = conn;
        if (tn < ntiles) conn_n = ElemIO<NDIM>::load_conn(tile_conn, eb_n + (tid < ne_n ? tid : 0));

        int total;
        const int off = block_scan_excl(len, sScan, &total);
        if (tid < nr) {
            sDelta[tid] = p0 - off;
            for (int k = 0; k < len; ++k) sE2R[off + k] = (unsigned short)tid;
        }
        for (int i = tid; i < total; i += THREADS) sV[i] = 0.0;
        __syncthreads();

#pragma unroll 1
        for (int pass = 0; pass < (MASS ? 2 : 1); ++pass) {
            // ---- phase B: row-owner accumulation in ascending element order
            if (tid < nr) {
                double f = 0.0;
                for (int j = jp0; j < jp1; ++j) {
                    const int k4 = j - jp0;
                    uint32_t rec;
                    if (k4 < 4) rec = k4 == 0 ? rec4[0] : k4 == 1 ? rec4[1] : k4 == 2 ? rec4[2] : rec4[3];
                    else rec = __ldg(pair_rec + j);
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
                    if (pass == 0) f += sFe[a * EMAX + slot];
                }
                if (pass == 0) Flux[row] = f;
            }
            __syncthreads();
            if (pass == 0 && tn < ntiles && tid < ne_n) ElemIO<NDIM>::template prefetch<HAS_U>(points, velocity, conn_n);
            // ---- phase C: coalesced write-out
            double *out = pass == 0 ? K : M;
            for (int i = tid; i < total; i += THREADS) __stcs(out + i + sDelta[sE2R[i]], sV[i]);
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
        rb = rb_n; nr = nr_n; eb = eb_n; ne = ne_n; conn = conn_n;
    }
}

template <int NDIM> constexpr size_t fused_smem_bytes() {
    return 8 * (size_t)(EMAX * ElemIO<NDIM>::NPE * ElemIO<NDIM>::NPE + EMAX * ElemIO<NDIM>::NPE + VMAX) + 2 * (size_t)VMAX +
           4 * (size_t)RMAX + 4 * 64;
}

template <int NDIM, bool HAS_U, bool MASS>
int launch_fused(int64_t ntiles, cudaStream_t s, const double *points, const double *velocity,
                 const ElemOpts &o, const int32_t *indptr, const int32_t *indices, const int32_t *tile_rows,
                 const int32_t *tile_row_ptr, const int32_t *tile_conn, const int32_t *tile_elem_ptr,
                 const uint32_t *pair_rec, const int32_t *pair_ptr, double *K, double *M, double *Flux) {
    static bool configured = false;
    constexpr size_t bytes = fused_smem_bytes<NDIM>();
    if (!configured) {
        KB_CUDA(cudaFuncSetAttribute(k_assemble_fused<NDIM, HAS_U, MASS>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                     (int)bytes));
        configured = true;
    }
    const int64_t cap = (int64_t)kb_num_sms() * 2;           // persistent: 2 CTAs per SM
    const int grid = (int)(ntiles < cap ? ntiles : cap);
    k_assemble_fused<NDIM, HAS_U, MASS><<<grid, THREADS, bytes, s>>>(
        points, velocity, o, indptr, indices, tile_rows, tile_row_ptr, tile_conn, tile_elem_ptr, pair_rec, pair_ptr, K, M,
        Flux, (int)ntiles);
    KB_LAUNCH_CHECK();
    return 0;
}

}  // namespace

extern "C" int kb_assemble_fused(int ndim, const double *d_points, const int32_t *d_elements, int64_t nelem,
                                 int64_t nnode, const double *d_velocity, const kb_material *material, int mode,
                                 int petrov, const int32_t *d_indptr, const int32_t *d_indices, int64_t ntiles,
                                 const int32_t *d_tile_rows, const int32_t *d_tile_row_ptr,
                                 const int32_t *d_tile_conn, const int32_t *d_tile_elem_ptr,
                                 const uint32_t *d_pair_rec, const int32_t *d_pair_ptr, double *d_K, double *d_M,
                                 double *d_Flux, void *stream) {
    if (!d_points || !d_elements || !material || !d_indptr || !d_indices || !d_tile_rows || !d_tile_row_ptr ||
        !d_tile_conn || !d_tile_elem_ptr || !d_pair_rec || !d_pair_ptr || !d_K || !d_Flux)
        return KB_EINVAL;
    if ((ndim != 1 && ndim != 2) || ntiles < 0 || nelem < 0 || nnode < 0) return KB_EINVAL;
    if (mode != KB_MODE_REFERENCE && mode != KB_MODE_CONSISTENT) return KB_EINVAL;
    if (ntiles == 0) return 0;
    cudaStream_t s = (cudaStream_t)stream;
    const ElemOpts o = kb_make_opts(material, mode, petrov, d_M != nullptr);
#define KB_FUSED(ND, HU, MS)                                                                                     \
    return launch_fused<ND, HU, MS>(ntiles, s, d_points, d_velocity, o, d_indptr, d_indices, d_tile_rows,            \
                                    d_tile_row_ptr, d_tile_conn, d_tile_elem_ptr, d_pair_rec, d_pair_ptr, d_K, d_M,    \
                                    d_Flux)
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
