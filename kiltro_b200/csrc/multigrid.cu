// Opt-in geometric-multigrid preconditioner for quad-4 meshes numbered like Mesh.Rectangle (kiltro/MeshGeneration/Mesh.py:33-58).
//
// The reference solves K_b x = F_b with a sparse direct method (kiltro/FEMSolver.py:350-364); the prescribed replacement is
// Jacobi-preconditioned BiCGSTAB (krylov.cu), whose iteration count grows like 1/h.  This file provides ONE V-cycle as a
// fixed linear operator z = M^-1 v for the right-preconditioned BiCGSTAB of krylov_mg.cuh (kb_mg_bicgstab_solve):
//
//   * hierarchy: level l+1 has about half the elements of level l per direction (nx/2+1 nodes; non-nested when the element
//     count is odd), down to <= `coarsest` nodes; coordinates, velocities and Dirichlet flags of a coarse node are those of
//     the nearest finer node;
//   * operators are held as NINE-POINT STENCILS in fp32: per node the 8 off-diagonal coefficients c_k = a_k / a_diag of the
//     row scaled to a unit diagonal, interleaved (32 B per node, two 16-byte loads), + the diagonal d and 1/d (d = 0 marks
//     a node that is no unknown: Dirichlet node or padding).  Rows are padded to a multiple of 32 floats and every array
//     has MG_GUARD zero rows below and above, so the kernels need neither index arithmetic per neighbour nor boundary
//     branches.  Level 0 is converted from the CSR matrix the solver was handed (kb_apply_dirichlet_reduce's K_b); the
//     coarser levels are REDISCRETISED on the coarser grid by the path's own element routine (element.cuh) with the
//     Petrov-Galerkin weights of TemperaturePG (kiltro/VariationalPrinciple.py:216-261,311-342) -- Galerkin products
//     R A P of the central operator diverge once the coarse cell Peclet number passes 1;
//   * smoother: nu damped-Jacobi sweeps x += omega (bhat - Ahat x) before and after; 40 B per node cross HBM per sweep
//     instead of the 101 B/row of the fp64 CSR SpMV.  fp32 is enough for a preconditioner: the outer recurrence, its
//     residual and the verdict stay fp64 (right preconditioning: x and r remain consistent whatever M is);
//   * transfer: bilinear interpolation between the two rectangles from 1-D tables (kb_mg_transfer_table), restriction =
//     its transpose, both as gather kernels (deterministic);
//   * coarsest level: dense inverse (Gauss-Jordan on device, one CTA) applied as a mat-vec.
//
// Everything is stream-ordered; the V-cycle performs no host synchronisation and no allocation (caller-owned workspace).
#include <math.h>
#include <string.h>
#include <new>
#include <vector>
#include "multigrid.cuh"
#include "element.cuh"

using namespace kbel;
ElemOpts kb_make_opts(const kb_material *m, int mode, int petrov, int want_mass);   // element.cu

namespace {

constexpr int MG_MAXLEV = 24;
constexpr int MG_DENSE_MAX = 1023;        // rows of the coarsest dense system (shared-memory row / factor buffers)
constexpr int MG_DENSE_SMEM = 100;        // up to here the Gauss-Jordan elimination runs in shared memory
constexpr int MG_GUARD = 8;               // zero rows below / above every padded array (>= NS + unroll overshoot of k_mg_leg)

// stencil slots k = 0..7: (dy,dx) = (-1,-1) (-1,0) (-1,1) (0,-1) (0,1) (1,-1) (1,0) (1,1); node pp's coefficients: cf[8 pp + k]
struct MgLevel {
    int nx = 0, ny = 0, pitch = 0;
    int64_t np = 0;                       // ny * pitch (padded nodes, without the guard rows)
    float *cf = nullptr;                  // 8 coefficients per padded node
    float *d = nullptr, *dinv = nullptr;  // diagonal (0: not an unknown) and its reciprocal (0 there too)
    float *cbx = nullptr, *cby = nullptr; // levels >= 1: Petrov-Galerkin weight coefficients of the node (upwind-biased restriction)
    float *bh = nullptr, *xa = nullptr, *xb = nullptr, *r = nullptr;    // fp32 vectors
    double *pts = nullptr, *vel = nullptr;                              // levels >= 1
    uint8_t *fixed = nullptr;             // compact (nx * ny)
    // transfer to level l+1 (this level is the fine one)
    int *jx = nullptr, *jy = nullptr, *sx = nullptr, *sy = nullptr;
    float *wx = nullptr, *wy = nullptr;
    std::vector<int> h_jx, h_jy, h_sx, h_sy;
    std::vector<float> h_wx, h_wy;
};

}  // namespace

struct kb_mg {
    int nlev = 0;
    MgLevel L[MG_MAXLEV];
    int nu = 2;
    float omega[4] = {0.67f, 0.67f, 0.67f, 0.67f};   // damping of sweep 1..nu (pre- and post-smoothing alike)
    int64_t nfree = 0;
    int reps[MG_MAXLEV];            // coarse-grid corrections per visit of level l (1 everywhere: V-cycle); kb_mg_set_cycle
    int *free2pp = nullptr;
    int *n2f = nullptr;             // level 0, padded like the vectors: unknown of a node, -1 elsewhere
    int tail_first = -1;            // first level of the shared-memory tail (k_mg_tail), -1: none
    int ncoarse = 0;
    double *dense = nullptr;
    float *ainv = nullptr;
    int *flag = nullptr;
    bool ready = false;
    void *ws_base = nullptr;        // workspace the device pointers above were carved from
    KbMgSolveCache solve_cache = {nullptr, nullptr, nullptr, 0, 0, 0};
};

namespace {

size_t a256(size_t x) { return (x + 255) & ~(size_t)255; }
// elements of a padded array with `w` values per node: guard rows on both sides + 32 values of slack at either end
size_t padded_elems(const MgLevel &l, int w) { return (size_t)(l.ny + 2 * MG_GUARD) * l.pitch * w + 64; }
size_t padded_origin(const MgLevel &l, int w) { return (size_t)MG_GUARD * l.pitch * w + 32; }

// 1-D linear interpolation between nf and nc uniformly spaced nodes on the same interval: fine node i sits at
// s = i (nc-1)/(nf-1) in coarse units -> j = min(floor(s), nc-2), w = s - j  (x_f[i] = (1-w) x_c[j] + w x_c[j+1]);
// start[c] = first fine i with j(i) >= c, c = 0..nc (the transpose walks [start[c-1], start[c+1]) ).
void transfer_table(int nf, int nc, int *j, float *w, int *start) {
    for (int i = 0; i < nf; ++i) {
        const long long num = (long long)i * (nc - 1);
        long long jj = nf > 1 ? num / (nf - 1) : 0;
        if (jj > nc - 2) jj = nc - 2;
        if (jj < 0) jj = 0;
        j[i] = (int)jj;
        w[i] = nf > 1 ? (float)((double)(num - jj * (nf - 1)) / (double)(nf - 1)) : 0.f;
    }
    int i = 0;
    for (int c = 0; c <= nc; ++c) {
        while (i < nf && j[i] < c) ++i;
        start[c] = i;
    }
}

size_t carve(kb_mg *mg, char *base, bool assign) {
    size_t off = 0;
    auto take = [&](size_t bytes) { char *p = base + off; off += a256(bytes); return p; };
    for (int l = 0; l < mg->nlev; ++l) {
        MgLevel &L = mg->L[l];
        const size_t nn = (size_t)L.nx * L.ny;
        { float *p = (float *)take(sizeof(float) * padded_elems(L, 8)) + padded_origin(L, 8); if (assign) L.cf = p; }
        float *v[8];
        for (int k = 0; k < 8; ++k) v[k] = (float *)take(sizeof(float) * padded_elems(L, 1)) + padded_origin(L, 1);
        if (assign) { L.d = v[0]; L.dinv = v[1]; L.bh = v[2]; L.xa = v[3]; L.xb = v[4]; L.r = v[5]; L.cbx = v[6]; L.cby = v[7]; }
        { uint8_t *p = (uint8_t *)take(nn); if (assign) L.fixed = p; }
        if (l > 0) {
            double *p = (double *)take(sizeof(double) * 2 * nn), *q = (double *)take(sizeof(double) * 2 * nn);
            if (assign) { L.pts = p; L.vel = q; }
        }
        if (l + 1 < mg->nlev) {
            const MgLevel &Cn = mg->L[l + 1];
            int *a = (int *)take(sizeof(int) * L.nx), *b = (int *)take(sizeof(int) * L.ny);
            int *c = (int *)take(sizeof(int) * (Cn.nx + 1)), *d = (int *)take(sizeof(int) * (Cn.ny + 1));
            float *e = (float *)take(sizeof(float) * L.nx), *f = (float *)take(sizeof(float) * L.ny);
            if (assign) { L.jx = a; L.jy = b; L.sx = c; L.sy = d; L.wx = e; L.wy = f; }
        }
    }
    const MgLevel &L0 = mg->L[0];
    const size_t nn0 = (size_t)L0.nx * L0.ny;
    {
        int *p = (int *)take(sizeof(int) * nn0);
        int *r = (int *)take(sizeof(int) * padded_elems(L0, 1)) + padded_origin(L0, 1);
        if (assign) { mg->free2pp = p; mg->n2f = r; }
    }
    const size_t nc = (size_t)mg->ncoarse;
    { double *p = (double *)take(sizeof(double) * nc * 2 * nc); if (assign) mg->dense = p; }
    { float *p = (float *)take(sizeof(float) * nc * nc); if (assign) mg->ainv = p; }
    { int *p = (int *)take(sizeof(int) * 8); if (assign) mg->flag = p; }
    return off;
}

__device__ __forceinline__ float4 ld4(const float *p) { return *reinterpret_cast<const float4 *>(p); }
__device__ __forceinline__ float4 ld4s(const float *p) { return __ldcs(reinterpret_cast<const float4 *>(p)); }

// ------------------------------------------------------------------------------------------------ set-up kernels
// reduced DOF -> padded index (level 0), the padded node -> unknown map and the fixed-node mask, from the
// renumbering of kb_dirichlet_split (n2f is pre-filled with -1)
__global__ void __launch_bounds__(256) k_mg_free_map(int64_t nn, int nx, int pitch, const int32_t *__restrict__ renum,
                                                     int *__restrict__ free2pp, int *__restrict__ n2f,
                                                     uint8_t *__restrict__ fixed) {
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; p < nn; p += stride) {
        const int f = renum[p];
        const int pp = (int)(p / nx) * pitch + (int)(p % nx);
        fixed[p] = f < 0;
        n2f[pp] = f >= 0 ? f : -1;
        if (f >= 0) free2pp[f] = pp;
    }
}

// level 0: CSR rows of the reduced system -> unit-diagonal stencil coefficients.  The stencil slot of an entry is found by
// comparing its column with the unknowns of the row's eight grid neighbours (read from the padded node -> unknown map:
// consecutive rows read consecutive addresses) instead of gathering the node of every column.  An entry that belongs to
// none of them -- outside the 9-point neighbourhood -- raises flag[0].
__global__ void __launch_bounds__(256) k_mg_dia_from_csr(int64_t nfree, const int32_t *__restrict__ indptr,
                                                         const int32_t *__restrict__ indices, const double *__restrict__ vals,
                                                         const int *__restrict__ free2pp, const int *__restrict__ n2f, int pitch,
                                                         float *__restrict__ cf, float *__restrict__ d,
                                                         float *__restrict__ dinv, int *__restrict__ flag) {
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    bool bad = false;
    for (int64_t f = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; f < nfree; f += stride) {
        const int pp = free2pp[f];
        const int *c = n2f + pp;
        const int nb[8] = {c[-pitch - 1], c[-pitch], c[-pitch + 1], c[-1], c[1], c[pitch - 1], c[pitch], c[pitch + 1]};
        const int e0 = indptr[f], e1 = indptr[f + 1];
        double dg = 0.0;
        double off[8] = {0.0, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0};
        for (int e = e0; e < e1; ++e) {
            const int col = indices[e];
            const double v = vals[e];
            if (col == (int)f) { dg += v; continue; }
            bool hit = false;
#pragma unroll
            for (int k = 0; k < 8; ++k)
                if (col == nb[k]) { off[k] += v; hit = true; }
            bad |= !hit;
        }
        if (!(dg != 0.0) || !isfinite(dg)) dg = 1.0;
        const double inv = 1.0 / dg;
        d[pp] = (float)dg;
        dinv[pp] = (float)inv;
        float4 *dst = reinterpret_cast<float4 *>(cf + 8 * (int64_t)pp);
        dst[0] = make_float4((float)(off[0] * inv), (float)(off[1] * inv), (float)(off[2] * inv), (float)(off[3] * inv));
        dst[1] = make_float4((float)(off[4] * inv), (float)(off[5] * inv), (float)(off[6] * inv), (float)(off[7] * inv));
    }
    if (bad) flag[0] = 1;
}

// geometry of a coarse level: every coarse node takes the nearest node of the finer level
__global__ void __launch_bounds__(256) k_mg_sample(int nxc, int nyc, int nxf, int nyf, const double *__restrict__ pf,
                                                   const double *__restrict__ vf, const uint8_t *__restrict__ fixf,
                                                   double *__restrict__ pc, double *__restrict__ vc,
                                                   uint8_t *__restrict__ fixc) {
    const int64_t nn = (int64_t)nxc * nyc, stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; p < nn; p += stride) {
        const int ic = (int)(p % nxc), jc = (int)(p / nxc);
        const int fi = nxc > 1 ? (int)((2ll * ic * (nxf - 1) + (nxc - 1)) / (2ll * (nxc - 1))) : 0;
        const int fj = nyc > 1 ? (int)((2ll * jc * (nyf - 1) + (nyc - 1)) / (2ll * (nyc - 1))) : 0;
        const int64_t src = (int64_t)fj * nxf + fi;
        pc[2 * p] = pf[2 * src]; pc[2 * p + 1] = pf[2 * src + 1];
        if (vf != nullptr) { vc[2 * p] = vf[2 * src]; vc[2 * p + 1] = vf[2 * src + 1]; }
        fixc[p] = fixf[src];
    }
}

// coarse-level operator, rediscretised: one thread per node adds row `a` of the element matrices of its <= 4 incident
// elements (node order of Mesh.Rectangle: [c, c+1, c+nxn+1, c+nxn]) into the node's 9-point stencil
template <bool VEL>
__global__ void __launch_bounds__(128) k_mg_assemble_dia(int nx, int ny, int pitch, const double *__restrict__ pts,
                                                         const double *__restrict__ vel, const uint8_t *__restrict__ fixed,
                                                         ElemOpts o, float *__restrict__ cf, float *__restrict__ d,
                                                         float *__restrict__ dinv) {
    const int64_t nn = (int64_t)nx * ny, stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; p < nn; p += stride) {
        const int i = (int)(p % nx), j = (int)(p / nx);
        const int64_t pp = (int64_t)j * pitch + i;
        if (fixed[p]) continue;                               // d and c stay 0 (memset by the caller)
        double st[9];
#pragma unroll
        for (int k = 0; k < 9; ++k) st[k] = 0.0;
#pragma unroll 1
        for (int e = 0; e < 4; ++e) {
            const int ex = i - 1 + (e & 1), ey = j - 1 + (e >> 1);
            if (ex < 0 || ey < 0 || ex >= nx - 1 || ey >= ny - 1) continue;
            const int64_t n0 = (int64_t)ey * nx + ex;
            const int64_t nd[4] = {n0, n0 + 1, n0 + nx + 1, n0 + nx};
            double X[4][2], U[4][2];
#pragma unroll
            for (int a = 0; a < 4; ++a) {
                X[a][0] = pts[2 * nd[a]]; X[a][1] = pts[2 * nd[a] + 1];
                U[a][0] = VEL ? vel[2 * nd[a]] : 0.0; U[a][1] = VEL ? vel[2 * nd[a] + 1] : 0.0;
            }
            double Ke[16], Fe[4], Me[16];
            quad4(X, U, o, Ke, Fe, Me);
            // local index of (i,j): (0,0) -> 0, (1,0) -> 1, (1,1) -> 2, (0,1) -> 3 ; stencil slot of local node b
            const int lx = i - ex, ly = j - ey, a = ly == 0 ? lx : 3 - lx;
            const int bx[4] = {ex, ex + 1, ex + 1, ex}, by[4] = {ey, ey, ey + 1, ey + 1};
#pragma unroll
            for (int b = 0; b < 4; ++b) {
                double v = 0.0;
#pragma unroll
                for (int aa = 0; aa < 4; ++aa) v = (aa == a) ? Ke[4 * aa + b] : v;
                const int slot = (by[b] - j + 1) * 3 + (bx[b] - i + 1);
#pragma unroll
                for (int k = 0; k < 9; ++k) st[k] += (k == slot) ? v : 0.0;
            }
        }
        double dg = st[4];
        if (!(dg != 0.0) || !isfinite(dg)) dg = 1.0;
        const double inv = 1.0 / dg;
        d[pp] = (float)dg;
        dinv[pp] = (float)inv;
        float out[8];
#pragma unroll
        for (int k = 0; k < 9; ++k) {
            if (k == 4) continue;
            const int qi = i + (k % 3) - 1, qj = j + (k / 3) - 1;
            float cv = 0.f;
            if (qi >= 0 && qj >= 0 && qi < nx && qj < ny && !fixed[(int64_t)qj * nx + qi]) cv = (float)(st[k] * inv);
            out[k > 4 ? k - 1 : k] = cv;
        }
        float4 *dst = reinterpret_cast<float4 *>(cf + 8 * pp);
        dst[0] = make_float4(out[0], out[1], out[2], out[3]);
        dst[1] = make_float4(out[4], out[5], out[6], out[7]);
    }
}

// Petrov-Galerkin weight coefficients of a coarse node, cb = (alpha / |u|) u with the reference's alpha(Pe)
// (kiltro/VariationalPrinciple.py:311-330; Pe from the node's velocity and the x-spacing of the level, like the element
// routine's h = |X1 - X0|): the coarse operators are assembled with the test functions W = N + cb . grad_parent N, and the
// restriction tests the fine residual with the same W (k_mg_restrict).
__global__ void __launch_bounds__(256) k_mg_pg_weights(int nx, int ny, int pitch, const double *__restrict__ pts,
                                                       const double *__restrict__ vel, double kcond,
                                                       float *__restrict__ cbx, float *__restrict__ cby) {
    const int64_t nn = (int64_t)nx * ny, stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; p < nn; p += stride) {
        const int i = (int)(p % nx), j = (int)(p / nx);
        const int64_t q = i + 1 < nx ? p + 1 : p - 1;
        const double dx = pts[2 * q] - pts[2 * p], dy = pts[2 * q + 1] - pts[2 * p + 1];
        const double ux = vel[2 * p], uy = vel[2 * p + 1];
        const double coef = pg_coefficient(sqrt(ux * ux + uy * uy), sqrt(dx * dx + dy * dy), kcond);
        const int64_t pp = (int64_t)j * pitch + i;
        cbx[pp] = (float)(coef * ux);
        cby[pp] = (float)(coef * uy);
    }
}

// The same operator, tiled: a CTA owns AT_X x AT_Y nodes, evaluates the (AT_X + 1) x (AT_Y + 1) elements around them ONCE
// each into shared memory (k_mg_assemble_dia evaluates every element four times, once per node: the fp64 element routine
// is what that kernel costs), then every node gathers row `a` of its <= 4 elements in the same order -- bit-identical
// results (test_coarse_assembly_tiled_bitwise).
constexpr int AT_X = 32, AT_Y = 8, AT_EX = AT_X + 1, AT_EY = AT_Y + 1;
template <bool VEL>
__global__ void __launch_bounds__(AT_X * AT_Y) k_mg_assemble_dia_tile(int nx, int ny, int pitch, const double *__restrict__ pts,
                                                                     const double *__restrict__ vel,
                                                                     const uint8_t *__restrict__ fixed, ElemOpts o,
                                                                     float *__restrict__ cf, float *__restrict__ d,
                                                                     float *__restrict__ dinv) {
    extern __shared__ double s_ke[];                         // [AT_EX * AT_EY][16]
    const int i0 = blockIdx.x * AT_X, j0 = blockIdx.y * AT_Y;
    for (int e = threadIdx.x; e < AT_EX * AT_EY; e += AT_X * AT_Y) {
        const int ex = i0 - 1 + e % AT_EX, ey = j0 - 1 + e / AT_EX;
        if (ex < 0 || ey < 0 || ex >= nx - 1 || ey >= ny - 1) continue;
        const int64_t n0 = (int64_t)ey * nx + ex;
        const int64_t nd[4] = {n0, n0 + 1, n0 + nx + 1, n0 + nx};
        double X[4][2], U[4][2];
#pragma unroll
        for (int a = 0; a < 4; ++a) {
            X[a][0] = pts[2 * nd[a]]; X[a][1] = pts[2 * nd[a] + 1];
            U[a][0] = VEL ? vel[2 * nd[a]] : 0.0; U[a][1] = VEL ? vel[2 * nd[a] + 1] : 0.0;
        }
        double Ke[16], Fe[4], Me[16];
        quad4(X, U, o, Ke, Fe, Me);
#pragma unroll
        for (int k = 0; k < 16; ++k) s_ke[e * 16 + k] = Ke[k];
    }
    __syncthreads();
    const int i = i0 + threadIdx.x % AT_X, j = j0 + threadIdx.x / AT_X;
    if (i >= nx || j >= ny) return;
    const int64_t p = (int64_t)j * nx + i, pp = (int64_t)j * pitch + i;
    if (fixed[p]) return;                                    // d and c stay 0 (memset by the caller)
    double st[9];
#pragma unroll
    for (int k = 0; k < 9; ++k) st[k] = 0.0;
#pragma unroll
    for (int e = 0; e < 4; ++e) {
        const int ex = i - 1 + (e & 1), ey = j - 1 + (e >> 1);
        if (ex < 0 || ey < 0 || ex >= nx - 1 || ey >= ny - 1) continue;
        const double *Ke = s_ke + ((ey - (j0 - 1)) * AT_EX + (ex - (i0 - 1))) * 16;
        const int lx = i - ex, ly = j - ey, a = ly == 0 ? lx : 3 - lx;
        const int bx[4] = {ex, ex + 1, ex + 1, ex}, by[4] = {ey, ey, ey + 1, ey + 1};
#pragma unroll
        for (int b = 0; b < 4; ++b) {
            const double v = Ke[4 * a + b];
            const int slot = (by[b] - j + 1) * 3 + (bx[b] - i + 1);
#pragma unroll
            for (int k = 0; k < 9; ++k) st[k] += (k == slot) ? v : 0.0;
        }
    }
    double dg = st[4];
    if (!(dg != 0.0) || !isfinite(dg)) dg = 1.0;
    const double inv = 1.0 / dg;
    d[pp] = (float)dg;
    dinv[pp] = (float)inv;
    float out[8];
#pragma unroll
    for (int k = 0; k < 9; ++k) {
        if (k == 4) continue;
        const int qi = i + (k % 3) - 1, qj = j + (k / 3) - 1;
        float cv = 0.f;
        if (qi >= 0 && qj >= 0 && qi < nx && qj < ny && !fixed[(int64_t)qj * nx + qi]) cv = (float)(st[k] * inv);
        out[k > 4 ? k - 1 : k] = cv;
    }
    float4 *dst = reinterpret_cast<float4 *>(cf + 8 * pp);
    dst[0] = make_float4(out[0], out[1], out[2], out[3]);
    dst[1] = make_float4(out[4], out[5], out[6], out[7]);
}

// coarsest level: dense [Ahat | I] (fp64, n x 2n) from the stencil; rows of non-unknowns are identity rows
__global__ void __launch_bounds__(256) k_mg_dense_build(int nx, int ny, int pitch, const float *__restrict__ cf,
                                                        const float *__restrict__ d, double *__restrict__ W) {
    const int n = nx * ny, m = 2 * n, q = blockIdx.x;
    if (q >= n) return;
    double *row = W + (size_t)q * m;
    for (int c = threadIdx.x; c < m; c += blockDim.x) row[c] = (c == q || c == n + q) ? 1.0 : 0.0;
    __syncthreads();
    const int i = q % nx, j = q / nx;
    const int64_t pp = (int64_t)j * pitch + i;
    if (threadIdx.x < 8 && d[pp] != 0.f) {
        const int k8 = threadIdx.x, k = k8 >= 4 ? k8 + 1 : k8;
        const int qi = i + (k % 3) - 1, qj = j + (k / 3) - 1;
        if (qi >= 0 && qj >= 0 && qi < nx && qj < ny) row[qj * nx + qi] = (double)cf[8 * pp + k8];
    }
}

// Gauss-Jordan without pivoting (rows are scaled to a unit diagonal and the rediscretised operators are diagonally
// dominant); a vanishing / non-finite pivot raises flag[1].  One CTA.  The matrix is banded (half-width `band` = nodes per
// grid row + 1): below the pivot, column k is zero outside rows (k, k+band] throughout the elimination (the rows above
// fill in: the inverse is dense), so step k updates rows [0, k+band] over the window [k, n+k] of columns that can be
// non-zero.  Thread = one column of the window x every ngrp-th row.  SMEM: the whole [A | I] lives in shared memory
// (n <= MG_DENSE_SMEM; the default coarsest level) instead of taking an L2 round trip per update.
template <bool SMEM>
__global__ void __launch_bounds__(1024) k_mg_dense_invert(int n, int band, double *__restrict__ Wg, float *__restrict__ ainv,
                                                          int *__restrict__ flag) {
    extern __shared__ double s_dyn[];
    __shared__ double s_row[MG_DENSE_MAX + 1];
    __shared__ double s_fac[MG_DENSE_MAX];
    const int m = 2 * n, tid = threadIdx.x, T = blockDim.x;
    double *W = Wg;
    if (SMEM) {
        W = s_dyn;
        for (int e = tid; e < n * m; e += T) W[e] = Wg[e];
        __syncthreads();
    }
    const int width = n + 1, ngrp = T / width > 0 ? T / width : 1;
    const int mycol = tid % width, mygrp = tid / width;
    for (int k = 0; k < n; ++k) {
        const double piv = W[(size_t)k * m + k];
        if (!(fabs(piv) > 1.0e-12) || !isfinite(piv)) { if (tid == 0) flag[1] = 1; return; }   // uniform over the CTA
        const double inv = 1.0 / piv;
        const int r1 = k + band < n - 1 ? k + band : n - 1;
        for (int c = tid; c <= n; c += T) s_row[c] = W[(size_t)k * m + k + c] * inv;
        for (int r = tid; r <= r1; r += T) s_fac[r] = (r == k) ? 0.0 : W[(size_t)r * m + k];
        __syncthreads();
        if (mygrp < ngrp) {                                  // width <= blockDim.x (n <= MG_DENSE_MAX = 1023)
            const int c = mycol;
            const double rc = s_row[c];
            if (mygrp == 0) W[(size_t)k * m + k + c] = rc;
            // eight independent read-modify-writes in flight (one dependent round trip per row otherwise)
            for (int rb = mygrp; rb <= r1; rb += 8 * ngrp) {
                double wv[8];
#pragma unroll
                for (int u = 0; u < 8; ++u) {
                    const int r = rb + u * ngrp;
                    wv[u] = r <= r1 ? W[(size_t)r * m + k + c] : 0.0;
                }
#pragma unroll
                for (int u = 0; u < 8; ++u) {
                    const int r = rb + u * ngrp;
                    if (r <= r1) {
                        const double f = s_fac[r];
                        if (f != 0.0) W[(size_t)r * m + k + c] = wv[u] - f * rc;
                    }
                }
            }
        }
        __syncthreads();
    }
    for (int e = tid; e < n * n; e += T) ainv[e] = (float)W[(size_t)(e / n) * m + n + (e % n)];
}

// ------------------------------------------------------------------------------------------------ V-cycle kernels
// stage the fp64 reduced vector on the level-0 grid: bh = v / d (0 at non-unknowns), first sweep from zero x = omega bh.
// With q != NULL the staged vector is s = v - alpha[0] q, also written to s_out (BiCGSTAB's s = r - alpha v).
__global__ void __launch_bounds__(256) k_mg_entry(int ny, int pitch, const int *__restrict__ n2f,
                                                  const float *__restrict__ dinv, const double *__restrict__ v,
                                                  double *__restrict__ s_out, const double *__restrict__ alpha,
                                                  const double *__restrict__ q, float *__restrict__ bh,
                                                  float *__restrict__ xa, float omega, const int *__restrict__ done) {
    kb_pdl_wait();
    if (done != nullptr && *done != 0) return;
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= pitch) return;
    const double a = (q != nullptr) ? alpha[0] : 0.0;
    for (int j = blockIdx.y; j < ny; j += gridDim.y) {
        const int64_t pp = (int64_t)j * pitch + i;
        float b = 0.f;
        const int f = n2f[pp];
        if (f >= 0) {
            double vv = v[f];
            if (q != nullptr) { vv -= a * __ldcs(q + f); s_out[f] = vv; }
            b = (float)vv * dinv[pp];
        }
        bh[pp] = b;
        xa[pp] = omega * b;
    }
}

__global__ void __launch_bounds__(256) k_mg_exit(int64_t nfree, const int *__restrict__ free2pp, const float *__restrict__ x,
                                                 double *__restrict__ z, const int *__restrict__ done) {
    kb_pdl_wait();
    if (done != nullptr && *done != 0) return;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t f = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; f < nfree; f += stride) z[f] = (double)x[free2pp[f]];
}

// the arithmetic of one node, shared by the one-kernel-per-sweep form and the fused legs (same order => same bits):
// res = bh - x - sum_k c_k x_k ;  sweep: x + omega res ;  residual: d res
__device__ __forceinline__ float stencil_res(const float (&c)[8], float dl, float dc, float dr, float ml, float mc, float mr,
                                             float ul, float uc, float ur, float bh) {
    // two independent chains of four (a single chain of eight dependent FMAs showed up as fixed-latency stalls)
    float a0 = c[0] * dl, a1 = c[4] * mr;
    a0 = fmaf(c[1], dc, a0); a1 = fmaf(c[5], ul, a1);
    a0 = fmaf(c[2], dr, a0); a1 = fmaf(c[6], uc, a1);
    a0 = fmaf(c[3], ml, a0); a1 = fmaf(c[7], ur, a1);
    return (bh - mc) - (a0 + a1);
}

// MODE 0: damped-Jacobi sweep  out = x + omega (bh - x - sum_k c_k x_k)     MODE 1: residual  out = d (bh - x - sum_k c_k x_k)
// Flat over the padded grid, one node per thread (the guard rows and the zero coefficients of padding / boundary nodes
// make index arithmetic and branches unnecessary).  The one-kernel-per-sweep form: fallback for nu > 3 and the
// cross-check of the fused legs.
template <int MODE>
__global__ void __launch_bounds__(256) k_mg_sweep(int64_t np, int pitch, const float *__restrict__ cf,
                                                  const float *__restrict__ d, const float *__restrict__ bh,
                                                  const float *__restrict__ xi, float *__restrict__ out, float omega,
                                                  const int *__restrict__ done) {
    kb_pdl_wait();
    if (done != nullptr && *done != 0) return;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; p < np; p += stride) {
        const float *xc = xi + p, *xd = xc - pitch, *xu = xc + pitch;
        const float4 lo = ld4s(cf + 8 * p), hi = ld4s(cf + 8 * p + 4);
        const float cc[8] = {lo.x, lo.y, lo.z, lo.w, hi.x, hi.y, hi.z, hi.w};
        const float mc = xc[0];
        const float res = stencil_res(cc, xd[-1], xd[0], xd[1], xc[-1], mc, xc[1], xu[-1], xu[0], xu[1], bh[p]);
        out[p] = MODE == 0 ? fmaf(omega, res, mc) : d[p] * res;
    }
}

// coarse rhs = P^T (fine residual), scaled to the unit-diagonal system, + the first sweep from zero.  A coarse node
// gathers the fine nodes of [start[c-1], start[c+1]) in both directions (<= RMAX per direction with the x weights in
// registers; wider supports -- only when a direction is not coarsened -- fall back to the table walk).
constexpr int RMAX = 6;
// UPWIND: the fine residual is tested with the coarse level's Petrov-Galerkin functions W_c = N_c + cb . grad_parent N_c
// instead of the hat functions N_c alone (cbx / cby: k_mg_pg_weights; grad_parent N_c = +-1/2 x the hat function of the other
// direction inside the two elements left / right of the node, 0 on the node's own grid line): the coarse operators are
// assembled with these test functions, and a restriction that ignores them leaves the coarse problems with the wrong
// right-hand side for exactly the smooth streamline modes the coarse levels are there for (Richardson factor per cycle on
// the bench problem: 0.45 without).
__global__ void __launch_bounds__(128) k_mg_restrict(int nxc, int nyc, int pitch_c, int pitch_f, const float *__restrict__ dinvc,
                                                     const float *__restrict__ cbx, const float *__restrict__ cby,
                                                     const float *__restrict__ rf, const int *__restrict__ jx,
                                                     const float *__restrict__ wx, const int *__restrict__ sx,
                                                     const int *__restrict__ jy, const float *__restrict__ wy,
                                                     const int *__restrict__ sy, float *__restrict__ bhc,
                                                     float *__restrict__ xac, float omega, float overweight,
                                                     const int *__restrict__ done) {
    kb_pdl_wait();
    if (done != nullptr && *done != 0) return;
    const int ic = blockIdx.x * blockDim.x + threadIdx.x;
    if (ic >= pitch_c) return;
    int i0 = 0, cnt = 0;
    float wxr[RMAX], sxr[RMAX];
#pragma unroll
    for (int u = 0; u < RMAX; ++u) { wxr[u] = 0.f; sxr[u] = 0.f; }
    if (ic < nxc) {
        i0 = sx[ic > 0 ? ic - 1 : 0];
        cnt = sx[ic + 1 < nxc ? ic + 1 : nxc] - i0;
        if (cnt <= RMAX) {
#pragma unroll
            for (int u = 0; u < RMAX; ++u)
                if (u < cnt) {
                    const bool right = jx[i0 + u] == ic;
                    const float w = wx[i0 + u];
                    wxr[u] = right ? 1.f - w : w;
                    sxr[u] = right ? (w > 0.f ? -1.f : 0.f) : (w < 1.f ? 1.f : 0.f);
                }
        }
    }
    for (int jc = blockIdx.y; jc < nyc; jc += gridDim.y) {
        const int64_t ppc = (int64_t)jc * pitch_c + ic;
        float b = 0.f;
        const float dv = ic < nxc ? dinvc[ppc] : 0.f;
        if (dv != 0.f) {
            const float hx = cbx != nullptr ? 0.5f * cbx[ppc] : 0.f, hy = cby != nullptr ? 0.5f * cby[ppc] : 0.f;
            const int j0 = sy[jc > 0 ? jc - 1 : 0], j1 = sy[jc + 1 < nyc ? jc + 1 : nyc];
            float acc = 0.f;
            for (int jf = j0; jf < j1; ++jf) {
                const bool up = jy[jf] == jc;
                const float wyf = wy[jf];
                const float wyv = up ? 1.f - wyf : wyf;
                const float syv = up ? (wyf > 0.f ? -1.f : 0.f) : (wyf < 1.f ? 1.f : 0.f);
                const float *row = rf + (int64_t)jf * pitch_f + i0;
                float racc = 0.f, rdx = 0.f;
                if (cnt <= RMAX) {
#pragma unroll
                    for (int u = 0; u < RMAX; ++u)
                        if (u < cnt) { const float rv = row[u]; racc = fmaf(wxr[u], rv, racc); rdx = fmaf(sxr[u], rv, rdx); }
                } else {
                    for (int u = 0; u < cnt; ++u) {
                        const bool right = jx[i0 + u] == ic;
                        const float w = wx[i0 + u], rv = row[u];
                        racc = fmaf(right ? 1.f - w : w, rv, racc);
                        rdx = fmaf(right ? (w > 0.f ? -1.f : 0.f) : (w < 1.f ? 1.f : 0.f), rv, rdx);
                    }
                }
                acc = fmaf(wyv, fmaf(hx, rdx, racc), acc);
                acc = fmaf(hy * syv, racc, acc);
            }
            b = (overweight * acc) * dv;
        }
        bhc[ppc] = b;
        xac[ppc] = omega * b;
    }
}

// x_f += P x_c (also at nodes that are no unknowns: nothing reads them -- their columns are zero in every stencil, the
// residual there is multiplied by d = 0, and the level-0 output is gathered at the unknowns only)
__device__ __forceinline__ float prolong_value(float c00, float c01, float c10, float c11, float wa, float wb) {
    const float lo = fmaf(wa, c01 - c00, c00), hi = fmaf(wa, c11 - c10, c10);
    return fmaf(wb, hi - lo, lo);
}
__global__ void __launch_bounds__(128) k_mg_prolong(int nxf, int nyf, int pitch_f, int pitch_c, const float *__restrict__ xc,
                                                    const int *__restrict__ jx, const float *__restrict__ wx,
                                                    const int *__restrict__ jy, const float *__restrict__ wy,
                                                    float *__restrict__ xf, const int *__restrict__ done) {
    kb_pdl_wait();
    if (done != nullptr && *done != 0) return;
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nxf) return;
    const int a = jx[i];
    const float wa = wx[i];
    for (int j = blockIdx.y; j < nyf; j += gridDim.y) {
        const int64_t pp = (int64_t)j * pitch_f + i;
        const float *r0 = xc + (int64_t)jy[j] * pitch_c + a, *r1 = r0 + pitch_c;
        xf[pp] += prolong_value(r0[0], r0[1], r1[0], r1[1], wa, wy[j]);
    }
}

// coarsest level: x = Ainv bh (one CTA, one warp per row)
__global__ void __launch_bounds__(1024) k_mg_dense_solve(int n, int nx, int pitch, const float *__restrict__ ainv,
                                                         const float *__restrict__ bh, float *__restrict__ x,
                                                         const int *__restrict__ done) {
    kb_pdl_wait();
    if (done != nullptr && *done != 0) return;
    __shared__ float s_b[MG_DENSE_MAX + 1];
    for (int q = threadIdx.x; q < n; q += blockDim.x) s_b[q] = bh[(int64_t)(q / nx) * pitch + q % nx];
    __syncthreads();
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = blockDim.x >> 5;
    for (int r = w; r < n; r += nw) {
        const float *row = ainv + (size_t)r * n;
        float acc = 0.f;
        for (int c = lane; c < n; c += 32) acc = fmaf(row[c], s_b[c], acc);
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(KB_FULL, acc, o);
        if (lane == 0) x[(int64_t)(r / nx) * pitch + r % nx] = acc;
    }
}

// ------------------------------------------------------------------------------------------------ fused legs
// One kernel per leg of the cycle on a level instead of one per sweep (temporal blocking): a WARP sweeps a strip of 32
// node columns upwards, one node row per step, and runs every stage of the leg on it in a software pipeline --
//   down-leg:  stage 0  x = omega bh (level 0: bh staged from the fp64 reduced vector, s = r - alpha v on the way)
//              stages 1..NS-1  damped-Jacobi sweeps, stage NS  residual r = d (bh - Ahat x)           -> x, r (+ bh) out
//   up-leg:    stage 0  x += P x_coarse, stages 1..NS  sweeps                        -> x out (level 0: fp64 reduced z)
// Stage s lags stage s-1 by one row; a lane keeps the last three rows (left, centre, right neighbour values, exchanged
// by warp shuffles) of every stage and the coefficients of the last NS rows in registers, so every coefficient, rhs and
// x value crosses HBM ONCE per leg instead of once per sweep.  No shared memory, no barrier, no branch in the step: the
// guard rows and the clamped column index make every load unconditional.  NS halo columns per side and NS halo rows per
// task are recomputed by the neighbouring strips (valid lanes [NS, 32-NS)); the arithmetic of a node is stencil_res /
// prolong_value exactly as in the one-kernel-per-sweep form, so both forms agree bit for bit (test_fused_legs_bitwise).
struct LegArgs {
    int nx, ny, pitch, rows_per_task, nstrips, ntasks;
    const float *cf, *d, *dinv;
    float omega[4];           // damping of sweep 1, 2, ... of the leg
    const float *bh_in;       // rhs (down-leg of levels >= 1, up-leg)
    float *bh_out;            // level-0 down-leg: rhs staged here
    const double *v;          // level-0 down-leg: fp64 reduced vector in ...
    double *s_out;            // ... and where v - alpha q goes when q != NULL
    const double *alpha, *q;
    const int *n2f;           // level 0
    float *x_out, *r_out;
    const float *x_in;        // up-leg: x after the down-leg
    const float *xc;          // up-leg: coarse correction
    int pitch_c;
    const int *jx, *jy;
    const float *wx, *wy;
    double *z;                // level-0 up-leg: fp64 reduced output
    int pfd;                  // L2 prefetch distance in rows (0 = off)
    int pf1;                  // L1 prefetch distance in rows (0 = off)
    int64_t nfree;
    const int *done;
};

// (two CTAs of 8 warps per SM: nu = 3 needs ~120 registers; capping it at 85 for a third CTA spills and ran 16 % slower,
// 10-warp CTAs at a cap of 96 registers 5 % slower)
template <int NS, bool DOWN, bool L0, int MINB>
__global__ void __launch_bounds__(256, MINB) k_mg_leg(const LegArgs a) {
    kb_pdl_wait();
    if (a.done != nullptr && *a.done != 0) return;
    static_assert(NS >= 1 && NS + 4 <= MG_GUARD, "guard rows cover the pipeline depth + the unroll overshoot + one row of prefetch");
    constexpr int U = NS + 1 > 3 ? NS + 1 : 3;        // ring depth of every per-row history (and unroll factor)
    constexpr int VW = 32 - 2 * NS;                   // columns a warp owns
    const int lane = threadIdx.x & 31;
    const int task = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (task >= a.ntasks) return;
    const int strip = task % a.nstrips, ty = task / a.nstrips;
    const int c = strip * VW - NS + lane;
    const int pitch = a.pitch;
    const int cl = c < 0 ? 0 : (c >= pitch ? pitch - 1 : c);          // column the lane loads (results of lanes outside
                                                                      // the grid are never stored and only ever meet zero
                                                                      // coefficients in their neighbours' stencils)
    const int j0 = ty * a.rows_per_task, j1 = j0 + a.rows_per_task < a.ny ? j0 + a.rows_per_task : a.ny;
    const bool own_col = c >= 0 && c < a.nx && lane >= NS && lane < 32 - NS;
    float cf[U][8], bhh[U], dhh[U], win[NS][U][3];
    int fhh[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
        bhh[u] = 0.f; dhh[u] = 0.f; fhh[u] = -1;
#pragma unroll
        for (int k = 0; k < 8; ++k) cf[u][k] = 0.f;
#pragma unroll
        for (int s = 0; s < NS; ++s) { win[s][u][0] = 0.f; win[s][u][1] = 0.f; win[s][u][2] = 0.f; }
    }
    int pjx = 0; float pwx = 0.f;                      // per-lane constants of the prolongation
    if (!DOWN) { const int ci = cl < a.nx ? cl : a.nx - 1; pjx = a.jx[ci]; pwx = a.wx[ci]; }
    const double alpha = (DOWN && L0 && a.q != nullptr) ? a.alpha[0] : 0.0;

    // L2 prefetch, PFD rows ahead (the loads of a step are consumed one step later: that covers an L2 hit, not a DRAM
    // round trip).  Lanes 0..15 cover the eight 128-byte lines of the warp's coefficient row segment (+ the misaligned
    // ninth), lane pairs 16.. the two lines of the other streams.
    const char *pf_ptr = nullptr;
    int pf_stride = 0;
    if (a.pfd > 0) {
        const int cbase = strip * VW - NS > 0 ? strip * VW - NS : 0;
        if (lane < 9) { pf_ptr = reinterpret_cast<const char *>(a.cf + 8 * (int64_t)cbase) + 128 * lane; pf_stride = pitch * 32; }
        else {
            const int sel = (lane - 9) >> 1, half = (lane - 9) & 1;
            const void *base = nullptr;
            if (sel == 0) base = DOWN ? (const void *)a.d : (const void *)a.bh_in;
            else if (sel == 1) base = DOWN ? (L0 ? (const void *)a.dinv : (const void *)a.bh_in) : (const void *)a.x_in;
            else if (sel == 2 && L0) base = a.n2f;
            if (base != nullptr) { pf_ptr = reinterpret_cast<const char *>(base) + 4 * (int64_t)cbase + 128 * half; pf_stride = pitch * 4; }
        }
    }

    // stage-0 inputs of the next row, loaded one step ahead (the node -> unknown map of the level-0 down-leg two steps
    // ahead: the fp64 vector is addressed through it)
    float p_bh = 0.f, p_d = 0.f, p_di = 0.f, p_x = 0.f, p_c00 = 0.f, p_c01 = 0.f, p_c10 = 0.f, p_c11 = 0.f, p_wy = 0.f;
    double p_v = 0.0;
    int p_f = -1, n_f = -1;
    auto load_stage0 = [&](int row, int off) {                    // off = row * pitch + cl
        if (DOWN) {
            p_d = a.d[off];
            if (L0) {
                const int f_prev = p_f;
                p_f = n_f;
                n_f = a.n2f[off + pitch];
                p_di = a.dinv[off];
                p_v = 0.0;
                if (p_f >= 0) {
                    p_v = a.v[p_f];
                    if (a.q != nullptr) p_v -= alpha * __ldcs(a.q + p_f);
                }
                if (a.pfd > 2 && (lane & 15) == 0 && p_f >= 0 && f_prev >= 0) {   // predicted address of the vector rows ahead
                    int64_t fp = (int64_t)p_f + (int64_t)(a.pfd - 1) * (p_f - f_prev);
                    fp = fp < 0 ? 0 : (fp >= a.nfree ? a.nfree - 1 : fp);
                    asm volatile("prefetch.global.L2 [%0];" ::"l"(a.v + fp));
                    if (a.q != nullptr) asm volatile("prefetch.global.L2 [%0];" ::"l"(a.q + fp));
                }
            } else {
                p_bh = a.bh_in[off];
            }
        } else {
            p_bh = a.bh_in[off];
            p_x = a.x_in[off];
            const int rr = row < 0 ? 0 : (row >= a.ny ? a.ny - 1 : row);
            p_wy = a.wy[rr];
            const float *r0 = a.xc + a.jy[rr] * a.pitch_c + pjx;
            p_c00 = r0[0]; p_c01 = r0[1]; p_c10 = r0[a.pitch_c]; p_c11 = r0[a.pitch_c + 1];
            if (L0) p_f = a.n2f[off];
        }
    };

    const int T0 = j0 - NS, T1 = j1 + NS;             // stage 0 walks rows [T0, T1)
    // running offsets instead of row * pitch + column per access (the index arithmetic was two thirds of the instructions
    // of a step): off = (row t, column cl); a lane that stores has c == cl, and stage s stores at off - s * pitch
    int off = T0 * pitch + cl;
    int tj = T0 - j0;                                  // t - j0: stage s stores when 0 <= tj - s < nrows
    const unsigned nrows = (unsigned)(j1 - j0);
    const float *pcf = a.cf + 8 * (int64_t)off;
    const char *pf_run = pf_ptr != nullptr ? pf_ptr + (int64_t)(T0 + a.pfd) * pf_stride : nullptr;
    const int row_end = a.ny + MG_GUARD;               // rows < row_end exist in every padded array
    if (DOWN && L0) n_f = a.n2f[off];
    load_stage0(T0, off);
    for (int tb = T0; tb < T1; tb += U) {
#pragma unroll
        for (int ph = 0; ph < U; ++ph) {
            const int t = tb + ph;
            // ---- stage 0 on row t
            float b0, x0;
            if (DOWN) {
                if (L0) b0 = (p_f >= 0) ? (float)p_v * p_di : 0.f;
                else b0 = p_bh;
                x0 = a.omega[0] * b0;
                const bool st0 = own_col && (unsigned)tj < nrows;
                if (L0 && st0) {
                    a.bh_out[off] = b0;
                    if (a.q != nullptr && p_f >= 0) a.s_out[p_f] = p_v;
                }
                if (NS == 1 && st0) a.x_out[off] = x0;
            } else {
                b0 = p_bh;
                x0 = p_x + prolong_value(p_c00, p_c01, p_c10, p_c11, pwx, p_wy);
            }
            bhh[ph] = b0;
            dhh[ph] = p_d;
            fhh[ph] = p_f;
            win[0][ph][1] = x0;
            win[0][ph][0] = __shfl_up_sync(KB_FULL, x0, 1);
            win[0][ph][2] = __shfl_down_sync(KB_FULL, x0, 1);
            // ---- loads: the coefficients of row t (first used by stage 1 at the next step), stage-0 inputs of row t+1
            {
                const float4 lo = ld4s(pcf), hi = ld4s(pcf + 4);
                cf[ph][0] = lo.x; cf[ph][1] = lo.y; cf[ph][2] = lo.z; cf[ph][3] = lo.w;
                cf[ph][4] = hi.x; cf[ph][5] = hi.y; cf[ph][6] = hi.z; cf[ph][7] = hi.w;
            }
            load_stage0(t + 1, off + pitch);
            if (pf_run != nullptr && t + a.pfd < row_end) asm volatile("prefetch.global.L2 [%0];" ::"l"(pf_run));
            if (a.pf1 > 0 && t + a.pf1 < row_end) {   // ... and into L1 a couple of rows ahead: the next steps' loads hit L1
                const int offp = off + a.pf1 * pitch;
                asm volatile("prefetch.global.L1 [%0];" ::"l"(pcf + 8 * (int64_t)(a.pf1 * pitch)));
                if (DOWN) {
                    asm volatile("prefetch.global.L1 [%0];" ::"l"(a.d + offp));
                    if (L0) asm volatile("prefetch.global.L1 [%0];" ::"l"(a.dinv + offp));
                    else asm volatile("prefetch.global.L1 [%0];" ::"l"(a.bh_in + offp));
                } else {
                    asm volatile("prefetch.global.L1 [%0];" ::"l"(a.bh_in + offp));
                    asm volatile("prefetch.global.L1 [%0];" ::"l"(a.x_in + offp));
                }
                if (L0) asm volatile("prefetch.global.L1 [%0];" ::"l"(a.n2f + offp));
            }
            // ---- stages 1..NS on rows t-1 .. t-NS
#pragma unroll
            for (int s = 1; s <= NS; ++s) {
                const int sl = (ph - s + 2 * U) % U, sd = (sl + U - 1) % U, su = (sl + 1) % U;
                const float (&w)[U][3] = win[s - 1];
                const float xc0 = w[sl][1];
                const float res = stencil_res(cf[sl], w[sd][0], w[sd][1], w[sd][2], w[sl][0], xc0, w[sl][2],
                                              w[su][0], w[su][1], w[su][2], bhh[sl]);
                const bool store = own_col && (unsigned)(tj - s) < nrows;
                const int offs = off - s * pitch;
                if (DOWN && s == NS) {
                    if (store) a.r_out[offs] = dhh[sl] * res;
                } else {
                    const float val = fmaf(a.omega[DOWN ? s : s - 1], res, xc0);   // down: sweep s+1 of nu, up: sweep s
                    if (s < NS) {
                        win[s][sl][1] = val;
                        win[s][sl][0] = __shfl_up_sync(KB_FULL, val, 1);
                        win[s][sl][2] = __shfl_down_sync(KB_FULL, val, 1);
                    }
                    if (DOWN && s == NS - 1 && store) a.x_out[offs] = val;
                    if (!DOWN && s == NS && store) {
                        if (L0) { if (fhh[sl] >= 0) a.z[fhh[sl]] = (double)val; }
                        else a.x_out[offs] = val;
                    }
                }
            }
            off += pitch; ++tj; pcf += 8 * pitch;
            if (pf_run != nullptr) pf_run += pf_stride;
        }
    }
}

// ------------------------------------------------------------------------------------------------ coarse tail
// The coarsest levels (<= TAIL_NODES nodes each, e.g. 33^2, 17^2, 9^2) as ONE kernel of one CTA with everything in shared
// memory: as separate kernels each of them is three launches of pure latency per visit (~28 us), and the cycle schedule
// visits them several times per cycle.  Same operations in the same order as the per-level kernels -- first sweep from
// zero, nu-1 sweeps, residual, restriction ... dense solve ... prolongation, nu sweeps -- through the same device functions
// (stencil_res, prolong_value), so the result is bit-identical (test_fused_legs_bitwise runs with and without it).
// Arrays live in shared memory with a one-node zero border (pitch nx+2), coefficients as eight planes.
constexpr int TAIL_MAXLEV = 4;
constexpr int TAIL_NODES = 1200;            // a level joins the tail when nx * ny <= TAIL_NODES
struct TailLevelArgs {
    int nx, ny, pitch;                      // global layout
    const float *cf, *d, *dinv;             // global
    const int *jx, *jy, *sx, *sy;           // transfer tables towards the next (coarser) tail level
    const float *wx, *wy;
    int off;                                // first float of the level's block in shared memory
};
struct TailArgs {
    int nlev, nu, ncoarse;
    TailLevelArgs L[TAIL_MAXLEV];
    const float *bh_in;                     // global rhs of the first tail level (padded layout)
    float *x_out;                           // global x of the first tail level after the sub-cycle (padded layout)
    const float *ainv;
    float omega[4];
    const int *done;
};
__host__ __device__ inline int tail_ps(int nx) { return nx + 2; }
__host__ __device__ inline int tail_block_floats(int nx, int ny) { return 14 * (nx + 2) * (ny + 2); }   // 8 c + d + dinv + bh + xa + xb + r

__global__ void __launch_bounds__(1024) k_mg_tail(const TailArgs a) {
    kb_pdl_wait();
    if (a.done != nullptr && *a.done != 0) return;
    extern __shared__ float s_tail[];
    __shared__ float s_b[MG_DENSE_MAX + 1];
    const int tid = threadIdx.x, T = blockDim.x;
    // ---- load: coefficients, d, 1/d of every tail level; vectors zeroed (borders stay zero)
    for (int l = 0; l < a.nlev; ++l) {
        const TailLevelArgs &L = a.L[l];
        const int ps = tail_ps(L.nx), plane = ps * (L.ny + 2);
        float *base = s_tail + L.off;
        for (int e = tid; e < 14 * plane; e += T) base[e] = 0.f;
    }
    __syncthreads();
    for (int l = 0; l < a.nlev; ++l) {
        const TailLevelArgs &L = a.L[l];
        const int ps = tail_ps(L.nx), plane = ps * (L.ny + 2);
        float *base = s_tail + L.off;
        for (int q = tid; q < L.nx * L.ny; q += T) {
            const int i = q % L.nx, j = q / L.nx;
            const int64_t pp = (int64_t)j * L.pitch + i;
            const int sp = (j + 1) * ps + i + 1;
            const float4 lo = ld4(L.cf + 8 * pp), hi = ld4(L.cf + 8 * pp + 4);
            base[0 * plane + sp] = lo.x; base[1 * plane + sp] = lo.y; base[2 * plane + sp] = lo.z; base[3 * plane + sp] = lo.w;
            base[4 * plane + sp] = hi.x; base[5 * plane + sp] = hi.y; base[6 * plane + sp] = hi.z; base[7 * plane + sp] = hi.w;
            base[8 * plane + sp] = L.d[pp];
            base[9 * plane + sp] = L.dinv[pp];
            if (l == 0) base[10 * plane + sp] = a.bh_in[pp];
        }
    }
    __syncthreads();
    auto sweep = [&](int l, int src, int dst, float om, bool residual) {          // planes: 10 bh, 11 xa, 12 xb, 13 r
        const TailLevelArgs &L = a.L[l];
        const int ps = tail_ps(L.nx), plane = ps * (L.ny + 2);
        const float *base = s_tail + L.off;
        const float *x = base + src * plane;
        float *out = s_tail + L.off + dst * plane;
        for (int q = tid; q < L.nx * L.ny; q += T) {
            const int i = q % L.nx, j = q / L.nx;
            const int sp = (j + 1) * ps + i + 1;
            float cc[8];
#pragma unroll
            for (int k = 0; k < 8; ++k) cc[k] = base[k * plane + sp];
            const float mc = x[sp];
            const float res = stencil_res(cc, x[sp - ps - 1], x[sp - ps], x[sp - ps + 1], x[sp - 1], mc, x[sp + 1],
                                          x[sp + ps - 1], x[sp + ps], x[sp + ps + 1], base[10 * plane + sp]);
            out[sp] = residual ? base[8 * plane + sp] * res : fmaf(om, res, mc);
        }
        __syncthreads();
    };
    int cur[TAIL_MAXLEV];
    // ---- down
    for (int l = 0; l + 1 < a.nlev; ++l) {
        const TailLevelArgs &L = a.L[l], &C = a.L[l + 1];
        const int ps = tail_ps(L.nx), plane = ps * (L.ny + 2);
        float *base = s_tail + L.off;
        for (int q = tid; q < L.nx * L.ny; q += T) {                                  // first sweep from zero
            const int sp = (q / L.nx + 1) * ps + q % L.nx + 1;
            base[11 * plane + sp] = a.omega[0] * base[10 * plane + sp];
        }
        __syncthreads();
        int c = 11, o = 12;
        for (int k = 1; k < a.nu; ++k) { sweep(l, c, o, a.omega[k], false); const int t = c; c = o; o = t; }
        sweep(l, c, 13, 0.f, true);
        cur[l] = c;
        // restriction (the arithmetic of k_mg_restrict with the Petrov-Galerkin / over-weighting switches off)
        const int psc = tail_ps(C.nx), planec = psc * (C.ny + 2);
        float *cb = s_tail + C.off;
        const float *r = base + 13 * plane;
        for (int q = tid; q < C.nx * C.ny; q += T) {
            const int ic = q % C.nx, jc = q / C.nx;
            const int spc = (jc + 1) * psc + ic + 1;
            const float dv = cb[9 * planec + spc];
            float b = 0.f;
            if (dv != 0.f) {
                const int i0 = L.sx[ic > 0 ? ic - 1 : 0], cnt = L.sx[ic + 1 < C.nx ? ic + 1 : C.nx] - i0;
                const int j0 = L.sy[jc > 0 ? jc - 1 : 0], j1 = L.sy[jc + 1 < C.ny ? jc + 1 : C.ny];
                float acc = 0.f;
                for (int jf = j0; jf < j1; ++jf) {
                    const float wyf = L.wy[jf];
                    const float wyv = (L.jy[jf] == jc) ? 1.f - wyf : wyf;
                    const float *row = r + (jf + 1) * ps + i0 + 1;
                    float racc = 0.f;
                    for (int u = 0; u < cnt; ++u) {
                        const float w = L.wx[i0 + u];
                        racc = fmaf((L.jx[i0 + u] == ic) ? 1.f - w : w, row[u], racc);
                    }
                    acc = fmaf(wyv, racc, acc);
                }
                b = acc * dv;
            }
            cb[10 * planec + spc] = b;
        }
        __syncthreads();
    }
    // ---- coarsest: x = Ainv bh (one warp per row, as k_mg_dense_solve)
    {
        const int l = a.nlev - 1;
        const TailLevelArgs &L = a.L[l];
        const int ps = tail_ps(L.nx), plane = ps * (L.ny + 2), n = a.ncoarse;
        float *base = s_tail + L.off;
        for (int q = tid; q < n; q += T) s_b[q] = base[10 * plane + (q / L.nx + 1) * ps + q % L.nx + 1];
        __syncthreads();
        const int lane = tid & 31, w = tid >> 5, nw = T >> 5;
        for (int r = w; r < n; r += nw) {
            const float *row = a.ainv + (size_t)r * n;
            float acc = 0.f;
            for (int c = lane; c < n; c += 32) acc = fmaf(row[c], s_b[c], acc);
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(KB_FULL, acc, o);
            if (lane == 0) base[12 * plane + (r / L.nx + 1) * ps + r % L.nx + 1] = acc;
        }
        cur[l] = 12;
        __syncthreads();
    }
    // ---- up
    for (int l = a.nlev - 2; l >= 0; --l) {
        const TailLevelArgs &L = a.L[l], &C = a.L[l + 1];
        const int ps = tail_ps(L.nx), plane = ps * (L.ny + 2);
        const int psc = tail_ps(C.nx), planec = psc * (C.ny + 2);
        float *xf = s_tail + L.off + cur[l] * plane;
        const float *xc = s_tail + C.off + cur[l + 1] * planec;
        for (int q = tid; q < L.nx * L.ny; q += T) {
            const int i = q % L.nx, j = q / L.nx;
            const float *r0 = xc + (L.jy[j] + 1) * psc + L.jx[i] + 1, *r1 = r0 + psc;
            xf[(j + 1) * ps + i + 1] += prolong_value(r0[0], r0[1], r1[0], r1[1], L.wx[i], L.wy[j]);
        }
        __syncthreads();
        int c = cur[l], o = (c == 11) ? 12 : 11;
        for (int k = 0; k < a.nu; ++k) { sweep(l, c, o, a.omega[k], false); const int t = c; c = o; o = t; }
        cur[l] = c;
    }
    // ---- out
    {
        const TailLevelArgs &L = a.L[0];
        const int ps = tail_ps(L.nx), plane = ps * (L.ny + 2);
        const float *x = s_tail + L.off + cur[0] * plane;
        for (int q = tid; q < L.nx * L.ny; q += T) {
            const int i = q % L.nx, j = q / L.nx;
            a.x_out[(int64_t)j * L.pitch + i] = x[(j + 1) * ps + i + 1];
        }
    }
}

dim3 grid2d(int width, int block, int ny) {
    const int gx = (int)kb_cdiv(width, block);
    int gy = (int)((int64_t)kb_num_sms() * 16 / gx);
    if (gy < 1) gy = 1;
    if (gy > ny) gy = ny;
    if (gy > 65535) gy = 65535;
    return dim3(gx, gy);
}

template <int MODE>
int launch_sweep(const MgLevel &L, const float *xi, float *out, float omega, const int *done, cudaStream_t s, bool pdl) {
    KB_CUDA(kb_launch_pdl(k_mg_sweep<MODE>, dim3(kb_grid(L.np, 256, 8)), dim3(256), 0, s, pdl, L.np, L.pitch,
                          (const float *)L.cf, (const float *)L.d, (const float *)L.bh, xi, out, omega, done));
    KB_LAUNCH_CHECK();
    return 0;
}

int g_mg_upwind = 0;          // 1: restriction with the coarse Petrov-Galerkin test functions (kb_mg_set_upwind_restriction)
float g_mg_overweight = 1.f;  // residual over-weighting of the restriction (kb_mg_set_overweight)

int launch_restrict(const MgLevel &L, const MgLevel &Cn, float omega, const int *done, cudaStream_t s, bool pdl) {
    // one coarse node per thread (no row loop: the gather is a chain of dependent table and residual loads)
    const dim3 grid((unsigned)kb_cdiv(Cn.pitch, 128), (unsigned)(Cn.ny < 65535 ? Cn.ny : 65535));
    KB_CUDA(kb_launch_pdl(k_mg_restrict, grid, dim3(128), 0, s, pdl, Cn.nx, Cn.ny, Cn.pitch,
                          L.pitch, (const float *)Cn.dinv, (const float *)(g_mg_upwind ? Cn.cbx : nullptr),
                          (const float *)(g_mg_upwind ? Cn.cby : nullptr), (const float *)L.r, (const int *)L.jx, (const float *)L.wx,
                          (const int *)L.sx, (const int *)L.jy, (const float *)L.wy, (const int *)L.sy, Cn.bh, Cn.xa,
                          omega, g_mg_overweight, done));
    KB_LAUNCH_CHECK();
    return 0;
}

int g_mg_aggressive = 0;      // levels with <= this many nodes per direction are coarsened by 4 instead of 2 (kb_mg_set_aggressive)
int g_mg_fused = 1;           // 1: fused legs (k_mg_leg) where nu <= 3 ; 0: one kernel per sweep (kb_mg_set_fused)
int g_mg_rows_per_task = 0;   // node rows per warp task of the fused legs; 0 = chosen per level (kb_mg_set_rows_per_task)
int g_mg_pfd = 6;             // L2 prefetch distance of the fused legs, in rows (kb_mg_set_prefetch_rows)
int g_mg_pf1 = 2;             // L1 prefetch distance (kb_mg_set_prefetch_rows_l1)
int g_mg_min_rows = 2;        // shortest task of the fused legs (kb_mg_set_min_rows)

template <int NS, bool DOWN, bool L0>
int launch_leg(LegArgs a, cudaStream_t s, bool pdl) {
    constexpr int VW = 32 - 2 * NS;
    a.nstrips = (int)kb_cdiv(a.nx, VW);
    if (g_mg_rows_per_task > 0) a.rows_per_task = g_mg_rows_per_task;
    else {      // enough warps to fill the machine a few times over, tasks no longer than 64 rows.  On the small levels a
                // leg is a chain of rows + 2 NS dependent steps (~0.6 us each) and nothing else: keep it short there
                // (the recomputed halo rows cost nothing on a level that does not fill the machine)
        int64_t r = (int64_t)a.nstrips * a.ny / ((int64_t)kb_num_sms() * 48);
        a.rows_per_task = (int)(r < g_mg_min_rows ? g_mg_min_rows : (r > 64 ? 64 : r));
    }
    a.pfd = g_mg_pfd;
    a.pf1 = g_mg_pf1;
    a.ntasks = a.nstrips * (int)kb_cdiv(a.ny, a.rows_per_task);
    KB_CUDA(kb_launch_pdl(k_mg_leg<NS, DOWN, L0, 2>, dim3((unsigned)kb_cdiv(a.ntasks, 8)), dim3(256), 0, s, pdl, a));
    KB_LAUNCH_CHECK();
    return 0;
}
template <bool DOWN, bool L0>
int launch_leg_ns(int ns, const LegArgs &a, cudaStream_t s, bool pdl) {
    switch (ns) {
        case 1: return launch_leg<1, DOWN, L0>(a, s, pdl);
        case 2: return launch_leg<2, DOWN, L0>(a, s, pdl);
        case 3: return launch_leg<3, DOWN, L0>(a, s, pdl);
        default: return KB_EINVAL;
    }
}

LegArgs leg_args(const kb_mg *mg, const MgLevel &L, const int *done) {
    LegArgs a = {};
    a.nx = L.nx; a.ny = L.ny; a.pitch = L.pitch; a.cf = L.cf; a.d = L.d; a.dinv = L.dinv; a.done = done;
    for (int k = 0; k < 4; ++k) a.omega[k] = mg->omega[k];
    a.nfree = mg->nfree;
    return a;
}

int launch_prolong(const MgLevel &L, const MgLevel &Cn, const float *xc, float *xf, const int *done, cudaStream_t s, bool pdl) {
    KB_CUDA(kb_launch_pdl(k_mg_prolong, grid2d(L.nx, 128, L.ny), dim3(128), 0, s, pdl, L.nx, L.ny, L.pitch, Cn.pitch, xc,
                          (const int *)L.jx, (const float *)L.wx, (const int *)L.jy, (const float *)L.wy, xf, done));
    KB_LAUNCH_CHECK();
    return 0;
}

int g_mg_tail = 1;            // 1: the coarsest levels run as one shared-memory kernel (kb_mg_set_tail)
int g_mg_assemble_tiled = 1;  // 1: coarse operators by k_mg_assemble_dia_tile (each element evaluated once per tile)

size_t tail_smem_bytes(const kb_mg *mg, int first) {
    size_t f = 0;
    for (int l = first; l < mg->nlev; ++l) f += (size_t)tail_block_floats(mg->L[l].nx, mg->L[l].ny);
    return f * sizeof(float);
}
bool tail_usable(const kb_mg *mg) {
    if (!g_mg_tail || mg->tail_first < 0 || g_mg_upwind || g_mg_overweight != 1.f || mg->nu > 4) return false;
    for (int l = mg->tail_first; l < mg->nlev; ++l) if (mg->reps[l] != 1) return false;
    return true;
}
int launch_tail(kb_mg *mg, const int *done, cudaStream_t s, bool pdl) {
    TailArgs a = {};
    const int first = mg->tail_first;
    a.nlev = mg->nlev - first; a.nu = mg->nu; a.ncoarse = mg->ncoarse; a.ainv = mg->ainv; a.done = done;
    for (int k = 0; k < 4; ++k) a.omega[k] = mg->omega[k];
    int off = 0;
    for (int k = 0; k < a.nlev; ++k) {
        const MgLevel &L = mg->L[first + k];
        TailLevelArgs &t = a.L[k];
        t.nx = L.nx; t.ny = L.ny; t.pitch = L.pitch; t.cf = L.cf; t.d = L.d; t.dinv = L.dinv;
        t.jx = L.jx; t.jy = L.jy; t.sx = L.sx; t.sy = L.sy; t.wx = L.wx; t.wy = L.wy;
        t.off = off;
        off += tail_block_floats(L.nx, L.ny);
    }
    a.bh_in = mg->L[first].bh; a.x_out = mg->L[first].xb;
    const size_t bytes = (size_t)off * sizeof(float);
    static KbOncePerDevice configured;
    if (configured.need()) KB_CUDA(cudaFuncSetAttribute(k_mg_tail, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    KB_CUDA(kb_launch_pdl(k_mg_tail, dim3(1), dim3(1024), bytes, s, pdl, a));
    KB_LAUNCH_CHECK();
    return 0;
}

// The cycle with fused legs, recursive: down-leg (pre-smoothing + residual) -> restriction -> cycle on the coarser levels
// -> up-leg (prolongation + post-smoothing).  mg->reps[l] > 1 repeats the coarse-grid correction of level l from the
// corrected iterate (prolongation and residual by the one-kernel-per-node forms, no smoothing in between) before the
// up-leg: the coarser levels are cycled reps[l] times per visit of level l.
struct FusedCtx {
    kb_mg *mg; const double *d_v; double *d_s, *d_z; const double *d_alpha, *d_q; const int *done; cudaStream_t s; bool pdl;
};
int cycle_fused(FusedCtx &c, int l) {
    kb_mg *mg = c.mg;
    const int nl = mg->nlev, ns = mg->nu;
    const MgLevel &L = mg->L[l];
    int rc;
    if (l == mg->tail_first && l > 0 && tail_usable(mg)) return launch_tail(mg, c.done, c.s, c.pdl);   // bh of l -> xb of l
    if (l == nl - 1) {
        KB_CUDA(kb_launch_pdl(k_mg_dense_solve, dim3(1), dim3(1024), 0, c.s, c.pdl, mg->ncoarse, L.nx, L.pitch,
                              (const float *)mg->ainv, (const float *)L.bh, L.xb, c.done));
        KB_LAUNCH_CHECK();
        return 0;
    }
    const MgLevel &Cn = mg->L[l + 1];
    {                                                        // down: x -> xa, residual -> r
        LegArgs a = leg_args(mg, L, c.done);
        a.x_out = L.xa; a.r_out = L.r;
        if (l == 0) {
            a.bh_out = L.bh; a.v = c.d_v; a.s_out = c.d_s; a.alpha = c.d_alpha; a.q = c.d_q; a.n2f = mg->n2f;
            rc = launch_leg_ns<true, true>(ns, a, c.s, c.pdl);
        } else {
            a.bh_in = L.bh;
            rc = launch_leg_ns<true, false>(ns, a, c.s, c.pdl);
        }
        if (rc) return rc;
    }
    for (int rep = 0;; ++rep) {
        if ((rc = launch_restrict(L, Cn, mg->omega[0], c.done, c.s, c.pdl))) return rc;      // rhs (and omega rhs) of l+1
        if ((rc = cycle_fused(c, l + 1))) return rc;                                         // -> Cn.xb
        if (rep + 1 >= mg->reps[l]) break;
        if ((rc = launch_prolong(L, Cn, Cn.xb, L.xa, c.done, c.s, c.pdl))) return rc;
        if ((rc = launch_sweep<1>(L, L.xa, L.r, 0.f, c.done, c.s, c.pdl))) return rc;
    }
    {                                                        // up: xa + P xb(l+1) -> xb ; level 0 -> z
        LegArgs a = leg_args(mg, L, c.done);
        a.bh_in = L.bh; a.x_in = L.xa; a.xc = Cn.xb; a.pitch_c = Cn.pitch; a.jx = L.jx; a.jy = L.jy; a.wx = L.wx; a.wy = L.wy;
        a.x_out = L.xb;
        if (l == 0) { a.z = c.d_z; a.n2f = mg->n2f; rc = launch_leg_ns<false, true>(ns, a, c.s, c.pdl); }
        else rc = launch_leg_ns<false, false>(ns, a, c.s, c.pdl);
        if (rc) return rc;
    }
    return 0;
}

// One-kernel-per-sweep form of the cycle, recursive: the same arithmetic as cycle_fused, one kernel per sweep (fallback for
// nu > 3 and the cross-check).  On entry L.bh holds the level's (unit-diagonal) right-hand side and c.cur[l] = omega_1 bh
// (the first sweep from zero).
struct CycleCtx {
    kb_mg *mg; const int *done; cudaStream_t s; bool pdl;
    float *cur[MG_MAXLEV], *oth[MG_MAXLEV];
};
int cycle_unfused(CycleCtx &c, int l) {
    kb_mg *mg = c.mg;
    const int nl = mg->nlev;
    const float *om = mg->omega;
    const MgLevel &L = mg->L[l];
    int rc;
    if (l == nl - 1) {
        KB_CUDA(kb_launch_pdl(k_mg_dense_solve, dim3(1), dim3(1024), 0, c.s, c.pdl, mg->ncoarse, L.nx, L.pitch,
                              (const float *)mg->ainv, (const float *)L.bh, c.cur[l], c.done));
        KB_LAUNCH_CHECK();
        return 0;
    }
    const MgLevel &Cn = mg->L[l + 1];
    for (int k = 1; k < mg->nu; ++k) {
        if ((rc = launch_sweep<0>(L, c.cur[l], c.oth[l], om[k], c.done, c.s, c.pdl))) return rc;
        float *t = c.cur[l]; c.cur[l] = c.oth[l]; c.oth[l] = t;
    }
    for (int rep = 0; rep < mg->reps[l]; ++rep) {
        if ((rc = launch_sweep<1>(L, c.cur[l], L.r, 0.f, c.done, c.s, c.pdl))) return rc;
        if ((rc = launch_restrict(L, Cn, om[0], c.done, c.s, c.pdl))) return rc;
        c.cur[l + 1] = Cn.xa; c.oth[l + 1] = Cn.xb;
        if ((rc = cycle_unfused(c, l + 1))) return rc;
        if ((rc = launch_prolong(L, Cn, c.cur[l + 1], c.cur[l], c.done, c.s, c.pdl))) return rc;
    }
    for (int k = 0; k < mg->nu; ++k) {
        if ((rc = launch_sweep<0>(L, c.cur[l], c.oth[l], om[k], c.done, c.s, c.pdl))) return rc;
        float *t = c.cur[l]; c.cur[l] = c.oth[l]; c.oth[l] = t;
    }
    return 0;
}

}  // namespace

int64_t kbmg_nfree(const kb_mg *mg) { return mg->nfree; }
KbMgSolveCache *kbmg_solve_cache(kb_mg *mg) { return &mg->solve_cache; }
bool kbmg_ready(const kb_mg *mg) { return mg->ready; }

int kbmg_apply(kb_mg *mg, const double *d_v, double *d_s, double *d_z, const double *d_alpha, const double *d_q,
               const int *done, cudaStream_t s, bool pdl) {
    if (!mg || !mg->ready || !d_v || !d_z || (d_q != nullptr && (!d_s || !d_alpha || d_s == d_v))) return KB_EINVAL;
    const int nl = mg->nlev;
    if (g_mg_fused && mg->nu <= 3 && nl >= 2) {
        FusedCtx f = {mg, d_v, d_s, d_z, d_alpha, d_q, done, s, pdl};
        return cycle_fused(f, 0);
    }
    CycleCtx c;
    c.mg = mg; c.done = done; c.s = s; c.pdl = pdl;
    const MgLevel &L0 = mg->L[0];
    KB_CUDA(kb_launch_pdl(k_mg_entry, grid2d(L0.pitch, 256, L0.ny), dim3(256), 0, s, pdl, L0.ny, L0.pitch,
                          (const int *)mg->n2f, (const float *)L0.dinv, d_v, d_s, d_alpha, d_q, L0.bh, L0.xa, mg->omega[0], done));
    KB_LAUNCH_CHECK();
    c.cur[0] = L0.xa; c.oth[0] = L0.xb;
    int rc;
    if ((rc = cycle_unfused(c, 0))) return rc;
    KB_CUDA(kb_launch_pdl(k_mg_exit, dim3(kb_grid(mg->nfree, 256, 8)), dim3(256), 0, s, pdl, mg->nfree,
                          (const int *)mg->free2pp, (const float *)c.cur[0], d_z, done));
    KB_LAUNCH_CHECK();
    return 0;
}

// ------------------------------------------------------------------------------------------------ C ABI
extern "C" int kb_mg_transfer_table(int64_t nfine, int64_t ncoarse, int32_t *h_j, float *h_w, int32_t *h_start) {
    if (!h_j || !h_w || !h_start || nfine < 2 || ncoarse < 2 || nfine > INT32_MAX || ncoarse > nfine) return KB_EINVAL;
    transfer_table((int)nfine, (int)ncoarse, h_j, h_w, h_start);
    return 0;
}

extern "C" int kb_mg_create(int64_t nxn, int64_t nyn, int64_t coarsest_nodes, kb_mg **out) {
    if (!out || nxn < 2 || nyn < 2) return KB_EINVAL;
    // 32-bit indices into the interleaved coefficient array (8 per padded node, guard rows included)
    if ((nyn + 2 * MG_GUARD) * ((nxn + 32) / 32 * 32) >= ((int64_t)1 << 28) - 64) return KB_ETOOBIG;
    if (coarsest_nodes < 4) coarsest_nodes = 4;
    if (coarsest_nodes > MG_DENSE_MAX) coarsest_nodes = MG_DENSE_MAX;
    kb_mg *mg = new (std::nothrow) kb_mg();
    if (!mg) return KB_EINVAL;
    for (int l = 0; l < MG_MAXLEV; ++l) mg->reps[l] = 1;
    int nx = (int)nxn, ny = (int)nyn;
    for (;;) {
        MgLevel &L = mg->L[mg->nlev++];
        L.nx = nx; L.ny = ny;
        L.pitch = (nx + 1 + 31) / 32 * 32;
        L.np = (int64_t)ny * L.pitch;
        if ((int64_t)nx * ny <= coarsest_nodes || mg->nlev == MG_MAXLEV) break;
        // below g_mg_aggressive nodes per direction a level is coarsened by 4 (development switch, default off)
        const int fx = nx <= g_mg_aggressive ? 4 : 2, fy = ny <= g_mg_aggressive ? 4 : 2;
        const int nxc = nx > 3 ? (nx - 1 + fx - 1) / fx + 1 : nx, nyc = ny > 3 ? (ny - 1 + fy - 1) / fy + 1 : ny;
        if (nxc == nx && nyc == ny) break;
        nx = nxc; ny = nyc;
    }
    const MgLevel &Lc = mg->L[mg->nlev - 1];
    if ((int64_t)Lc.nx * Lc.ny > MG_DENSE_MAX) { delete mg; return KB_ETOOBIG; }
    mg->ncoarse = Lc.nx * Lc.ny;
    for (int l = 1; l + 1 < mg->nlev; ++l)                  // shared-memory tail: >= 2 levels of <= TAIL_NODES nodes that fit
        if ((int64_t)mg->L[l].nx * mg->L[l].ny <= TAIL_NODES && mg->nlev - l <= TAIL_MAXLEV &&
            tail_smem_bytes(mg, l) <= 200 * 1024) { mg->tail_first = l; break; }
    for (int l = 0; l + 1 < mg->nlev; ++l) {
        MgLevel &L = mg->L[l];
        const MgLevel &Cn = mg->L[l + 1];
        L.h_jx.resize(L.nx); L.h_wx.resize(L.nx); L.h_sx.resize(Cn.nx + 1);
        L.h_jy.resize(L.ny); L.h_wy.resize(L.ny); L.h_sy.resize(Cn.ny + 1);
        transfer_table(L.nx, Cn.nx, L.h_jx.data(), L.h_wx.data(), L.h_sx.data());
        transfer_table(L.ny, Cn.ny, L.h_jy.data(), L.h_wy.data(), L.h_sy.data());
    }
    *out = mg;
    return 0;
}

extern "C" void kb_mg_destroy(kb_mg *mg) { delete mg; }

extern "C" int kb_mg_num_levels(const kb_mg *mg) { return mg ? mg->nlev : 0; }

extern "C" int kb_mg_level_shape(const kb_mg *mg, int level, int64_t *nxn, int64_t *nyn) {
    if (!mg || level < 0 || level >= mg->nlev || !nxn || !nyn) return KB_EINVAL;
    *nxn = mg->L[level].nx; *nyn = mg->L[level].ny;
    return 0;
}

extern "C" size_t kb_mg_workspace_bytes(const kb_mg *mg) {
    if (!mg) return 0;
    return carve(const_cast<kb_mg *>(mg), nullptr, false);
}

extern "C" int kb_mg_setup(kb_mg *mg, const double *d_points, const double *d_velocity, const int32_t *d_renum,
                           const kb_material *material, int mode, int64_t nfree, int64_t nnz, const int32_t *d_indptr,
                           const int32_t *d_indices, const double *d_vals, int nu, double omega, void *d_workspace,
                           size_t workspace_bytes, void *stream) {
    if (!mg || !d_points || !d_renum || !material || !d_indptr || !d_indices || !d_vals || !d_workspace || nfree < 1 ||
        nnz < 1 || nu < 1 || !(omega < 2.0))
        return KB_EINVAL;
    if (d_velocity != nullptr && mode != KB_MODE_CONSISTENT) return KB_EINVAL;   // coarse levels: W (x) u.gradN + PG weights
    if (workspace_bytes < carve(mg, nullptr, false)) return KB_EWORKSPACE;
    cudaStream_t s = (cudaStream_t)stream;
    mg->ready = false;
    const bool first = mg->ws_base != d_workspace;      // new workspace: vectors zeroed, transfer tables uploaded
    carve(mg, (char *)d_workspace, true);
    mg->ws_base = d_workspace;
    if (nu > 4 && !(omega > 0.0)) return KB_EINVAL;
    mg->nu = nu; mg->nfree = nfree;
    if (omega > 0.0) for (int k = 0; k < 4; ++k) mg->omega[k] = (float)omega;      // omega <= 0: keep kb_mg_set_smoother's factors
    MgLevel &L0 = mg->L[0];
    const int64_t nn0 = (int64_t)L0.nx * L0.ny;
    if (nfree > nn0) return KB_EINVAL;
    // coefficient arrays and vectors start from zero (padding, guard rows, absent neighbours)
    for (int l = 0; l < mg->nlev; ++l) {
        MgLevel &L = mg->L[l];
        KB_CUDA(cudaMemsetAsync(L.cf - padded_origin(L, 8), 0, sizeof(float) * padded_elems(L, 8), s));
        KB_CUDA(cudaMemsetAsync(L.d - padded_origin(L, 1), 0, sizeof(float) * padded_elems(L, 1), s));
        KB_CUDA(cudaMemsetAsync(L.dinv - padded_origin(L, 1), 0, sizeof(float) * padded_elems(L, 1), s));
        if (g_mg_upwind) {                                  // only the Petrov-Galerkin restriction switch reads them
            KB_CUDA(cudaMemsetAsync(L.cbx - padded_origin(L, 1), 0, sizeof(float) * padded_elems(L, 1), s));
            KB_CUDA(cudaMemsetAsync(L.cby - padded_origin(L, 1), 0, sizeof(float) * padded_elems(L, 1), s));
        }
        if (first) {
            float *v[4] = {L.bh, L.xa, L.xb, L.r};
            for (int k = 0; k < 4; ++k) KB_CUDA(cudaMemsetAsync(v[k] - padded_origin(L, 1), 0, sizeof(float) * padded_elems(L, 1), s));
            if (l + 1 < mg->nlev) {
                const MgLevel &Cn = mg->L[l + 1];
                KB_CUDA(cudaMemcpyAsync(L.jx, L.h_jx.data(), sizeof(int) * L.nx, cudaMemcpyHostToDevice, s));
                KB_CUDA(cudaMemcpyAsync(L.jy, L.h_jy.data(), sizeof(int) * L.ny, cudaMemcpyHostToDevice, s));
                KB_CUDA(cudaMemcpyAsync(L.wx, L.h_wx.data(), sizeof(float) * L.nx, cudaMemcpyHostToDevice, s));
                KB_CUDA(cudaMemcpyAsync(L.wy, L.h_wy.data(), sizeof(float) * L.ny, cudaMemcpyHostToDevice, s));
                KB_CUDA(cudaMemcpyAsync(L.sx, L.h_sx.data(), sizeof(int) * (Cn.nx + 1), cudaMemcpyHostToDevice, s));
                KB_CUDA(cudaMemcpyAsync(L.sy, L.h_sy.data(), sizeof(int) * (Cn.ny + 1), cudaMemcpyHostToDevice, s));
            }
        }
    }
    KB_CUDA(cudaMemsetAsync(mg->n2f - padded_origin(L0, 1), 0xff, sizeof(int) * padded_elems(L0, 1), s));   // -1
    KB_CUDA(cudaMemsetAsync(mg->flag, 0, sizeof(int) * 8, s));
    k_mg_free_map<<<kb_grid(nn0, 256, 8), 256, 0, s>>>(nn0, L0.nx, L0.pitch, d_renum, mg->free2pp, mg->n2f,
                                                       L0.fixed);
    KB_LAUNCH_CHECK();
    k_mg_dia_from_csr<<<kb_grid(nfree, 256, 8), 256, 0, s>>>(nfree, d_indptr, d_indices, d_vals, mg->free2pp, mg->n2f,
                                                            L0.pitch, L0.cf, L0.d, L0.dinv, mg->flag);
    KB_LAUNCH_CHECK();
    const ElemOpts o = kb_make_opts(material, KB_MODE_CONSISTENT, d_velocity != nullptr ? 1 : 0, 0);
    for (int l = 1; l < mg->nlev; ++l) {
        MgLevel &L = mg->L[l];
        const MgLevel &F = mg->L[l - 1];
        const int64_t nn = (int64_t)L.nx * L.ny;
        const double *pf = l == 1 ? d_points : F.pts, *vf = d_velocity == nullptr ? nullptr : (l == 1 ? d_velocity : F.vel);
        k_mg_sample<<<kb_grid(nn, 256, 8), 256, 0, s>>>(L.nx, L.ny, F.nx, F.ny, pf, vf, F.fixed, L.pts, L.vel, L.fixed);
        KB_LAUNCH_CHECK();
        if (d_velocity != nullptr && g_mg_upwind) {
            k_mg_pg_weights<<<kb_grid(nn, 256, 8), 256, 0, s>>>(L.nx, L.ny, L.pitch, L.pts, L.vel, o.k, L.cbx, L.cby);
            KB_LAUNCH_CHECK();
        }
        if (g_mg_assemble_tiled && kb_cdiv(L.ny, AT_Y) <= 65535) {
            const dim3 tg((unsigned)kb_cdiv(L.nx, AT_X), (unsigned)kb_cdiv(L.ny, AT_Y));
            const size_t sm = sizeof(double) * AT_EX * AT_EY * 16;
            static KbOncePerDevice configured;
            if (configured.need()) {
                KB_CUDA(cudaFuncSetAttribute(k_mg_assemble_dia_tile<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm));
                KB_CUDA(cudaFuncSetAttribute(k_mg_assemble_dia_tile<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm));
            }
            if (d_velocity != nullptr)
                k_mg_assemble_dia_tile<true><<<tg, AT_X * AT_Y, sm, s>>>(L.nx, L.ny, L.pitch, L.pts, L.vel, L.fixed, o, L.cf, L.d, L.dinv);
            else
                k_mg_assemble_dia_tile<false><<<tg, AT_X * AT_Y, sm, s>>>(L.nx, L.ny, L.pitch, L.pts, L.vel, L.fixed, o, L.cf, L.d, L.dinv);
        } else if (d_velocity != nullptr)
            k_mg_assemble_dia<true><<<kb_grid(nn, 128, 16), 128, 0, s>>>(L.nx, L.ny, L.pitch, L.pts, L.vel, L.fixed, o, L.cf, L.d, L.dinv);
        else
            k_mg_assemble_dia<false><<<kb_grid(nn, 128, 16), 128, 0, s>>>(L.nx, L.ny, L.pitch, L.pts, L.vel, L.fixed, o, L.cf, L.d, L.dinv);
        KB_LAUNCH_CHECK();
    }
    {
        const MgLevel &Lc = mg->L[mg->nlev - 1];
        const int n = mg->ncoarse;
        k_mg_dense_build<<<n, 256, 0, s>>>(Lc.nx, Lc.ny, Lc.pitch, Lc.cf, Lc.d, mg->dense);
        KB_LAUNCH_CHECK();
        if (n <= MG_DENSE_SMEM) {
            const size_t bytes = sizeof(double) * (size_t)n * 2 * n;
            static KbOncePerDevice configured;
            if (configured.need()) {
                KB_CUDA(cudaFuncSetAttribute(k_mg_dense_invert<true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                             (int)(sizeof(double) * MG_DENSE_SMEM * 2 * MG_DENSE_SMEM)));
            }
            k_mg_dense_invert<true><<<1, 1024, bytes, s>>>(n, Lc.nx + 1, mg->dense, mg->ainv, mg->flag);
        } else {
            k_mg_dense_invert<false><<<1, 1024, 0, s>>>(n, Lc.nx + 1, mg->dense, mg->ainv, mg->flag);
        }
        KB_LAUNCH_CHECK();
    }
    int flag[2] = {0, 0};
    KB_CUDA(cudaMemcpyAsync(flag, mg->flag, sizeof(int) * 2, cudaMemcpyDeviceToHost, s));
    KB_CUDA(cudaStreamSynchronize(s));
    if (flag[0]) return KB_EINVAL;          // the matrix has entries outside the 9-point neighbourhood of the grid
    if (flag[1]) return KB_EBREAKDOWN;      // singular coarsest operator
    mg->ready = true;
    return 0;
}

extern "C" int kb_mg_apply(kb_mg *mg, const double *d_v, double *d_z, void *stream) {
    return kbmg_apply(mg, d_v, nullptr, d_z, nullptr, nullptr, nullptr, (cudaStream_t)stream, false);
}

/* development switches: fused legs on / off (results are bit-identical), node rows per warp task of the fused legs
 * (0 = chosen per level), their L2 prefetch distance */
extern "C" void kb_mg_set_fused(int on) { g_mg_fused = on; }
extern "C" void kb_mg_set_tail(int on) { g_mg_tail = on; }
extern "C" void kb_mg_set_assemble_tiled(int on) { g_mg_assemble_tiled = on; }
/* cycle schedule: h_reps[l] = coarse-grid corrections per visit of level l (levels beyond n keep 1 = V-cycle) */
extern "C" int kb_mg_set_cycle(kb_mg *mg, int n, const int32_t *h_reps) {
    if (!mg || !h_reps || n < 0 || n > MG_MAXLEV) return KB_EINVAL;
    for (int l = 0; l < MG_MAXLEV; ++l) mg->reps[l] = (l < n && h_reps[l] >= 1 && h_reps[l] <= 8) ? h_reps[l] : 1;
    return 0;
}
extern "C" void kb_mg_set_aggressive(int nodes) { g_mg_aggressive = nodes; }
extern "C" void kb_mg_set_upwind_restriction(int on) { g_mg_upwind = on; }
extern "C" void kb_mg_set_overweight(double eta) { if (eta > 0.0 && eta < 4.0) g_mg_overweight = (float)eta; }
extern "C" void kb_mg_set_rows_per_task(int rows) { if (rows >= 0) g_mg_rows_per_task = rows; }
extern "C" void kb_mg_set_min_rows(int rows) { if (rows >= 1) g_mg_min_rows = rows; }
extern "C" void kb_mg_set_prefetch_rows(int rows) { if (rows >= 0) g_mg_pfd = rows; }
extern "C" void kb_mg_set_prefetch_rows_l1(int rows) { if (rows >= 0 && rows <= 2) g_mg_pf1 = rows; }

/* per-sweep damping factors (a Chebyshev / Richardson sequence instead of one omega): sweep k of the nu pre- and of the nu
 * post-smoothing sweeps uses h_omega[k]; nu <= 4.  Pass omega <= 0 to kb_mg_setup to keep them. */
extern "C" int kb_mg_set_smoother(kb_mg *mg, int nu, const double *h_omega) {
    if (!mg || !h_omega || nu < 1 || nu > 4) return KB_EINVAL;
    for (int k = 0; k < 4; ++k) mg->omega[k] = (float)h_omega[k < nu ? k : nu - 1];
    mg->nu = nu;
    return 0;
}
