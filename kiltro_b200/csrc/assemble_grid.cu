// K1+K3 fused, grid form (quad-4): numeric assembly for meshes with the numbering of Mesh.Rectangle
// (MeshGeneration/Mesh.py:33-58: node (i, j) = j * nxn + i, element (i, j) = [c, c+1, c+nxn+1, c+nxn], c = j * nxn + i),
// checked once on device by kb_grid_verify.  Same contract and the same summation order as k_assemble_tmpl /
// k_assemble_fused (replaces the element loop of FiniteElements/Assembly.py:46-72; results equal theirs bit for bit),
// but nothing goes through shared memory except the CSR values on their way out, and there is no block-level barrier:
//
//   * a WARP sweeps a strip of 31 node columns upwards.  Lane l computes the element matrices of element column
//     c = n0 - 1 + l, one element per step, in registers (element.cuh, sum-factorised) and keeps the two upper rows of
//     the matrix of the step before;
//   * the CSR row of node (c + 1, j) gets its <= 4 contributions per entry from the lane's own two element matrices
//     (elements (c, j-1), (c, j)) and from the two of lane l + 1 (elements (c+1, j-1), (c+1, j)), which arrive by
//     warp shuffles -- added in ascending element order, the reference's sequential `+=` (Assembly.py:46,60-61);
//   * the 31 rows of a step are contiguous in K.data: the lanes park their <= 9 values in a per-warp shared-memory
//     buffer at (indptr[row] - indptr[first row]) and the warp streams the block out with fully coalesced stores;
//   * per step a lane loads the two upper nodes of its element (coordinates and velocity, 16 bytes each, consecutive
//     lanes -> consecutive addresses); the connectivity and the column indices are not read at all.
//
// HBM traffic: 32 B per node (+ 1/31 + 1/G_STRIP overlap), 4 B per row of indptr, 8 B per non-zero out.
// No atomics, deterministic.  Lane 31 only feeds lane 30 (3 % redundant element work), the first step of a strip
// recomputes the element row below it (1 / G_STRIP).
#include <cstdlib>
#include "element.cuh"

using namespace kbel;

ElemOpts kb_make_opts(const kb_material *m, int mode, int petrov, int want_mass);   // element.cu

namespace {

constexpr int G_THREADS = 256;                 // 8 warps per CTA, 2 CTAs per SM
constexpr int G_WARPS = G_THREADS / 32;
constexpr int G_COLS = 31;                     // node columns owned by a warp
constexpr int G_STRIP = 32;                    // node rows per task (default; 64: 0.477 ms, 32: 0.451 ms, 128: 0.508 ms @S16)
constexpr int G_BUF = 9 * 32;                  // doubles parked per warp and step (31 rows x <= 9 entries)
constexpr int G_PFD = 2;                       // L2 prefetch distance in steps (node rows; default; 8 is slower than 4, 2 faster)

// a load whose issue point matters (one step ahead of its use): volatile, so the compiler does not sink it to the use
__device__ __forceinline__ double2 ldg_pin(const double2 *p) {
    double2 v;
    asm volatile("ld.global.nc.v2.f64 {%0, %1}, [%2];" : "=d"(v.x), "=d"(v.y) : "l"(p));
    return v;
}
__device__ __forceinline__ int ldg_pin(const int32_t *p) {
    int v;
    asm volatile("ld.global.nc.s32 %0, [%1];" : "=r"(v) : "l"(p));
    return v;
}
__device__ __forceinline__ void prefetch_l2(const void *p) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p)); }

__device__ __forceinline__ double shfl_next(double v) { return __shfl_down_sync(KB_FULL, v, 1); }
template <int W>
__device__ __forceinline__ double (&pick(double (&a)[16], double (&b)[16]))[16] {
    if constexpr (W == 0) return a; else return b;
}

__global__ void __launch_bounds__(256) k_grid_verify(const int4 *__restrict__ elements, int64_t nelem, int nxn,
                                                     int32_t *__restrict__ flag) {
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    const int nex = nxn - 1;
    bool bad = false;
    for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < nelem; e += stride) {
        const int4 c = __ldg(elements + e);
        const int64_t j = e / nex, i = e - j * nex;
        const int64_t n = j * nxn + i;
        bad |= (c.x != n) || (c.y != n + 1) || (c.z != n + nxn + 1) || (c.w != n + nxn);
    }
    if (bad) *flag = 1;
}

// WHICH = 0: K (+ Flux); WHICH = 1: M.  GENERAL: see assemble_tmpl.cu (second launch, general element routine for every
// element, runs only when the first launch met an element the sum-factorised routine declines).
// PF (development switch, KB_GRID_VARIANT): 0 = loads where they are used; 1 = row pointers loaded (one coalesced load per
// step) before the element is computed + L2 prefetch of the node rows G_PFD steps ahead; 2 = 1 + the upper nodes of the next
// element are loaded into registers one step ahead.
template <bool HAS_U, int WHICH, bool GENERAL, int PF>
__global__ void __launch_bounds__(G_THREADS, 2)
k_assemble_grid(const double2 *__restrict__ points, const double2 *__restrict__ velocity, ElemOpts o, int nxn, int nyn,
                const int32_t *__restrict__ indptr, double *__restrict__ out, double *__restrict__ Flux,
                int ntask_x, int ntasks, unsigned int *__restrict__ counter, int *__restrict__ nbad, int strip, int pfd) {
    __shared__ double s_buf[G_WARPS][G_BUF];
    if (GENERAL && *reinterpret_cast<volatile int *>(nbad) == 0) return;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    double *buf = s_buf[warp];
    const int nex = nxn - 1, ney = nyn - 1;
    const bool has_q = (o.area * o.q) != 0.0;
    bool bad = false;

    for (;;) {
        unsigned int task = 0;
        if (lane == 0) task = atomicAdd(counter, 1u);
        task = __shfl_sync(KB_FULL, task, 0);
        if (task >= (unsigned int)ntasks) break;
        const int ty = (int)(task / (unsigned int)ntask_x), tx = (int)(task - (unsigned int)ty * ntask_x);
        const int n0 = tx * G_COLS;                          // first owned node column
        const int j0 = ty * strip, j1 = min(j0 + strip, nyn);
        const int c = n0 - 1 + lane;                         // this lane's element column
        const int n = c + 1;                                 // ... and node column
        const bool col_ok = c >= 0 && c < nex;               // the element column exists
        const bool own = lane < G_COLS && n < nxn;           // the lane owns a node column
        const int cl = min(max(c, 0), nex - 1);              // clamped: loads stay in range, results are discarded

        // nodes of the row below the first element row handled (lower nodes of the first element)
        double X[4][2], U[4][2];
        const int jstart = max(j0 - 1, 0);
        {
            const int64_t b = (int64_t)jstart * nxn + cl;
            const double2 p0 = __ldg(points + b), p1 = __ldg(points + b + 1);
            X[3][0] = p0.x; X[3][1] = p0.y; X[2][0] = p1.x; X[2][1] = p1.y;
            if (HAS_U) {
                const double2 u0 = __ldg(velocity + b), u1 = __ldg(velocity + b + 1);
                U[3][0] = u0.x; U[3][1] = u0.y; U[2][0] = u1.x; U[2][1] = u1.y;
            } else {
                U[3][0] = U[3][1] = U[2][0] = U[2][1] = 0.0;
            }
        }
        // rows 2 (upper right node) and 3 (upper left node) of the element matrix of the step before, and its fe
        double p2[4] = {0.0, 0.0, 0.0, 0.0}, p3[4] = {0.0, 0.0, 0.0, 0.0}, pf2 = 0.0, pf3 = 0.0;
        const int nown = min(G_COLS, nxn - n0);                          // owned node columns of this strip
        double2 nq0, nq1, nu0, nu1;                                      // PF == 2: upper nodes of the next element
        if (PF == 2) {
            const int64_t b = (int64_t)min(jstart + 1, nyn - 1) * nxn + cl;
            nq0 = __ldg(points + b); nq1 = __ldg(points + b + 1);
            if (HAS_U) { nu0 = __ldg(velocity + b); nu1 = __ldg(velocity + b + 1); }
        }

        // element rows je = jstart .. j1 - 1 (je = j0 - 1 only primes p2 / p3); node row je is emitted after element je
        for (int je = jstart; je < j1; ++je) {
            // ---- element (c, je): its lower nodes are the upper nodes of the step before
            X[0][0] = X[3][0]; X[0][1] = X[3][1]; X[1][0] = X[2][0]; X[1][1] = X[2][1];
            U[0][0] = U[3][0]; U[0][1] = U[3][1]; U[1][0] = U[2][0]; U[1][1] = U[2][1];
            double ke[16], fe[4], me[16];
            const bool el_ok = col_ok && je < ney;
            int ip = 0;                                                  // PF >= 1: indptr[row0 + lane], loaded early
            if (PF >= 1) {
                ip = ldg_pin(indptr + (max(je, j0) * nxn + n0 + min(lane, nown)));
                const int bp = min(je + 1 + pfd, nyn - 1) * nxn + cl;      // (node indices fit 32 bits)
                prefetch_l2(points + bp);
                if (HAS_U) prefetch_l2(velocity + bp);
                if (lane < 16) prefetch_l2(indptr + (min(je + pfd, nyn - 1) * nxn + n0 + 2 * lane));
            }
            if (PF == 2) {
                X[3][0] = nq0.x; X[3][1] = nq0.y; X[2][0] = nq1.x; X[2][1] = nq1.y;
                if (HAS_U) { U[3][0] = nu0.x; U[3][1] = nu0.y; U[2][0] = nu1.x; U[2][1] = nu1.y; }
                const int b = min(je + 2, nyn - 1) * nxn + cl;
                nq0 = ldg_pin(points + b); nq1 = ldg_pin(points + b + 1);
                if (HAS_U) { nu0 = ldg_pin(velocity + b); nu1 = ldg_pin(velocity + b + 1); }
            } else {
                const int b = min(je + 1, nyn - 1) * nxn + cl;
                const double2 q0 = __ldg(points + b), q1 = __ldg(points + b + 1);
                X[3][0] = q0.x; X[3][1] = q0.y; X[2][0] = q1.x; X[2][1] = q1.y;
                if (HAS_U) {
                    const double2 u0 = __ldg(velocity + b), u1 = __ldg(velocity + b + 1);
                    U[3][0] = u0.x; U[3][1] = u0.y; U[2][0] = u1.x; U[2][1] = u1.y;
                }
            }
            if (GENERAL) {
                quad4_general(X, U, o, ke, fe, me);
            } else if (!quad4_fast(X, U, o, ke, fe, me)) {
                bad |= el_ok;
            }
            double (&mat)[16] = pick<WHICH>(ke, me);
            if (!el_ok) {
#pragma unroll
                for (int i = 0; i < 16; ++i) mat[i] = 0.0;
#pragma unroll
                for (int i = 0; i < 4; ++i) fe[i] = 0.0;
            }
            if (je >= j0) {
                // ---- CSR row of node (n, je): A = element (c, je-1) [rows kept in p2], B = (c+1, je-1) [lane+1's p3],
                //      C = (c, je) [row 1 of mat], D = (c+1, je) [lane+1's row 0]
                double b3[4], d0[4];
#pragma unroll
                for (int k = 0; k < 4; ++k) { b3[k] = shfl_next(p3[k]); d0[k] = shfl_next(mat[k]); }
                const double e0 = p2[0];
                const double e1 = p2[1] + b3[0];
                const double e2 = b3[1];
                const double e3 = p2[3] + mat[4];
                double e4 = p2[2] + b3[3]; e4 += mat[5]; e4 += d0[0];
                const double e5 = b3[2] + d0[1];
                const double e6 = mat[7];
                const double e7 = mat[6] + d0[3];
                const double e8 = d0[2];
                const int row = je * nxn + n;
                const int row0 = je * nxn + n0;
                int base, total, pos;
                if (PF >= 1) {
                    base = __shfl_sync(KB_FULL, ip, 0);
                    total = __shfl_sync(KB_FULL, ip, nown) - base;
                    pos = ip - base;
                } else {
                    base = __ldg(indptr + row0);
                    total = __ldg(indptr + row0 + nown) - base;
                    pos = own ? __ldg(indptr + row) - base : 0;
                }
                if (own) {
                    const bool xl = n > 0, xr = n < nxn - 1, yd = je > 0, yu = je < nyn - 1;
                    double *bq = buf + pos;
                    if (xl && xr && yd && yu) {                          // interior row: nine entries, fixed offsets
                        bq[0] = e0; bq[1] = e1; bq[2] = e2; bq[3] = e3; bq[4] = e4;
                        bq[5] = e5; bq[6] = e6; bq[7] = e7; bq[8] = e8;
                    } else {
                        if (yd) {
                            if (xl) *bq++ = e0;
                            *bq++ = e1;
                            if (xr) *bq++ = e2;
                        }
                        if (xl) *bq++ = e3;
                        *bq++ = e4;
                        if (xr) *bq++ = e5;
                        if (yu) {
                            if (xl) *bq++ = e6;
                            *bq++ = e7;
                            if (xr) *bq++ = e8;
                        }
                    }
                }
                if (WHICH == 0) {                                        // source vector          Assembly.py:69-72
                    double f = 0.0;
                    if (has_q) {                                         // (uniform over the grid)
                        const double fb = shfl_next(pf3), fd = shfl_next(fe[0]);
                        f = pf2 + fb; f += fe[1]; f += fd;
                    }
                    if (own) Flux[row] = f;
                }
                __syncwarp();
                {
                    double *op = out + base + lane;
                    const double *bp = buf + lane;
#pragma unroll
                    for (int k = 0; k < 9; ++k)
                        if (lane + 32 * k < total) __stcs(op + 32 * k, bp[32 * k]);
                }
                __syncwarp();
            }
#pragma unroll
            for (int k = 0; k < 4; ++k) { p2[k] = mat[8 + k]; p3[k] = mat[12 + k]; }
            pf2 = fe[2]; pf3 = fe[3];
        }
    }
    if (!GENERAL && bad) *nbad = 1;
}

int g_grid_variant = -1;      // development switch (environment KB_GRID_VARIANT): see PF above

template <bool HAS_U, int WHICH, int PF>
int launch_grid_pf(cudaStream_t s, const double *points, const double *velocity, const ElemOpts &o, int nxn, int nyn,
                   const int32_t *indptr, double *out, double *Flux, unsigned int *counters, int *nbad) {
    static int strip = 0, pfd = 0;                           // development switches KB_GRID_STRIP / KB_GRID_PFD
    if (strip == 0) {
        const char *e = getenv("KB_GRID_STRIP"), *f = getenv("KB_GRID_PFD");
        strip = e && atoi(e) > 0 ? atoi(e) : G_STRIP;
        pfd = f && atoi(f) > 0 ? atoi(f) : G_PFD;
    }
    const int ntx = (nxn + G_COLS - 1) / G_COLS, nty = (nyn + strip - 1) / strip;
    const int ntasks = ntx * nty;
    const int64_t cap = (int64_t)kb_num_sms() * 2;
    const int grid = (int)((ntasks + G_WARPS - 1) / G_WARPS < cap ? (ntasks + G_WARPS - 1) / G_WARPS : cap);
    k_assemble_grid<HAS_U, WHICH, false, PF><<<grid, G_THREADS, 0, s>>>((const double2 *)points, (const double2 *)velocity, o,
                                                                        nxn, nyn, indptr, out, Flux, ntx, ntasks, counters, nbad,
                                                                        strip, pfd);
    KB_LAUNCH_CHECK();
    k_assemble_grid<HAS_U, WHICH, true, 0><<<grid, G_THREADS, 0, s>>>((const double2 *)points, (const double2 *)velocity, o, nxn,
                                                                      nyn, indptr, out, Flux, ntx, ntasks, counters + 1, nbad,
                                                                      strip, pfd);
    KB_LAUNCH_CHECK();
    return 0;
}

template <bool HAS_U, int WHICH>
int launch_grid(cudaStream_t s, const double *points, const double *velocity, const ElemOpts &o, int nxn, int nyn,
                const int32_t *indptr, double *out, double *Flux, unsigned int *counters, int *nbad) {
    if (g_grid_variant < 0) {
        const char *e = getenv("KB_GRID_VARIANT");
        g_grid_variant = e ? atoi(e) : 1;
    }
    if (g_grid_variant == 0) return launch_grid_pf<HAS_U, WHICH, 0>(s, points, velocity, o, nxn, nyn, indptr, out, Flux, counters, nbad);
    if (g_grid_variant == 2) return launch_grid_pf<HAS_U, WHICH, 2>(s, points, velocity, o, nxn, nyn, indptr, out, Flux, counters, nbad);
    return launch_grid_pf<HAS_U, WHICH, 1>(s, points, velocity, o, nxn, nyn, indptr, out, Flux, counters, nbad);
}

}  // namespace

extern "C" int kb_grid_verify(const int32_t *d_elements, int64_t nelem, int64_t nxn, int64_t nyn, int32_t *d_flag,
                              void *stream) {
    if (!d_elements || !d_flag || nxn < 2 || nyn < 2) return KB_EINVAL;
    if (nxn * nyn > (int64_t)INT32_MAX) return KB_ETOOBIG;
    cudaStream_t s = (cudaStream_t)stream;
    KB_CUDA(cudaMemsetAsync(d_flag, 0, sizeof(int32_t), s));
    if (nelem != (nxn - 1) * (nyn - 1)) {
        const int32_t one = 1;
        KB_CUDA(cudaMemcpyAsync(d_flag, &one, sizeof(one), cudaMemcpyHostToDevice, s));
        KB_CUDA(cudaStreamSynchronize(s));
        return 0;
    }
    k_grid_verify<<<kb_grid(nelem, 256, 8), 256, 0, s>>>((const int4 *)d_elements, nelem, (int)nxn, d_flag);
    KB_LAUNCH_CHECK();
    return 0;
}

extern "C" int kb_assemble_grid(const double *d_points, const double *d_velocity, const kb_material *material, int mode,
                                int petrov, int64_t nxn, int64_t nyn, const int32_t *d_indptr, double *d_K, double *d_M,
                                double *d_Flux, void *d_scratch, void *stream) {
    if (!d_points || !material || !d_indptr || !d_K || !d_Flux || !d_scratch) return KB_EINVAL;
    if (nxn < 2 || nyn < 2) return KB_EINVAL;
    if (nxn * nyn > (int64_t)INT32_MAX - 64) return KB_ETOOBIG;         // node indices (+ the prefetch overshoot) fit 32 bits
    if (mode != KB_MODE_REFERENCE && mode != KB_MODE_CONSISTENT) return KB_EINVAL;
    cudaStream_t s = (cudaStream_t)stream;
    const ElemOpts o = kb_make_opts(material, mode, petrov, d_M != nullptr);
    unsigned int *counters = (unsigned int *)d_scratch;         // [0..3] task counters, [4] bad-element flag
    int *nbad = (int *)d_scratch + 4;
    KB_CUDA(cudaMemsetAsync(d_scratch, 0, 32, s));
    int rc;
    if (d_velocity) rc = launch_grid<true, 0>(s, d_points, d_velocity, o, (int)nxn, (int)nyn, d_indptr, d_K, d_Flux, counters, nbad);
    else rc = launch_grid<false, 0>(s, d_points, d_velocity, o, (int)nxn, (int)nyn, d_indptr, d_K, d_Flux, counters, nbad);
    if (rc) return rc;
    if (d_M) {
        if (d_velocity) rc = launch_grid<true, 1>(s, d_points, d_velocity, o, (int)nxn, (int)nyn, d_indptr, d_M, d_Flux, counters + 2, nbad);
        else rc = launch_grid<false, 1>(s, d_points, d_velocity, o, (int)nxn, (int)nyn, d_indptr, d_M, d_Flux, counters + 2, nbad);
    }
    return rc;
}
