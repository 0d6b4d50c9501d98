// Opt-in geometric-multigrid preconditioner for quad-4 meshes numbered like Mesh.Rectangle (kiltro/MeshGeneration/Mesh.py:33-58).
//
// The reference solves K_b x = F_b with a sparse direct method (kiltro/FEMSolver.py:350-364); the prescribed replacement is
// Jacobi-preconditioned BiCGSTAB (krylov.cu), whose iteration count grows like 1/h.  This file provides ONE V-cycle as a
// fixed linear operator z = M^-1 v for the right-preconditioned BiCGSTAB of krylov_mg.cuh (kb_mg_bicgstab_solve):
//
//   * hierarchy: level l+1 has about half the elements of level l per direction (nx/2+1 nodes; non-nested when the element
//     count is odd), down to <= `coarsest` nodes; coordinates, velocities and Dirichlet flags of a coarse node are those of
//     the nearest finer node;
//   * operators are held as NINE-POINT STENCILS in fp32, one array per neighbour (DIA layout, rows padded to a multiple of
//     32 floats, one zero guard row above and below, so a sweep needs no index arithmetic and no boundary branch).  Rows
//     are scaled to a unit diagonal: only the 8 off-diagonal coefficients c_k = a_k / a_diag and the diagonal d are stored
//     (d = 0 marks a node that is no unknown: Dirichlet node or padding).  Level 0 is converted from the CSR matrix the
//     solver was handed (kb_apply_dirichlet_reduce's K_b); the coarser levels are REDISCRETISED on the coarser grid by the
//     path's own element routine (element.cuh) with the Petrov-Galerkin weights of TemperaturePG
//     (kiltro/VariationalPrinciple.py:216-261,311-342) -- Galerkin products R A P of the central operator diverge once
//     the coarse cell Peclet number passes 1;
//   * smoother: nu damped-Jacobi sweeps x += omega (bhat - Ahat x) before and after; 36 B per node cross HBM per sweep
//     (8 coefficients + rhs, fp32) instead of the 101 B/row of the fp64 CSR SpMV.  fp32 is enough for a preconditioner:
//     the outer recurrence, its residual and the verdict stay fp64 (right preconditioning: x and r remain consistent
//     whatever M is);
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
constexpr int MG_DENSE_MAX = 1024;        // rows of the coarsest dense system (shared-memory row / factor buffers)

struct Coef8 { float *c[8]; };            // stencil slots: (dy,dx) = (-1,-1) (-1,0) (-1,1) (0,-1) (0,1) (1,-1) (1,0) (1,1)

struct MgLevel {
    int nx = 0, ny = 0, pitch = 0;
    int64_t np = 0;                       // ny * pitch (padded nodes)
    Coef8 C = {};
    float *d = nullptr;                   // diagonal (0: not an unknown)
    float *bh = nullptr, *xa = nullptr, *xb = nullptr, *r = nullptr;    // fp32 vectors (guarded)
    double *pts = nullptr, *vel = nullptr;                              // levels >= 1
    uint8_t *fixed = nullptr;
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
    float omega = 0.67f;
    int64_t nfree = 0;
    int *free2pp = nullptr, *free2node = nullptr, *node2free = nullptr;
    int ncoarse = 0;
    double *dense = nullptr;
    float *ainv = nullptr;
    int *flag = nullptr;
    bool has_velocity = false;
    bool ready = false;
    void *ws_base = nullptr;        // workspace the device pointers above were carved from
};

namespace {

size_t a256(size_t x) { return (x + 255) & ~(size_t)255; }
size_t vec_floats(const MgLevel &l) { return (size_t)(l.ny + 2) * l.pitch + 64; }

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
        for (int k = 0; k < 8; ++k) { float *p = (float *)take(sizeof(float) * L.np); if (assign) L.C.c[k] = p; }
        { float *p = (float *)take(sizeof(float) * L.np); if (assign) L.d = p; }
        float *v[4];
        for (int k = 0; k < 4; ++k) v[k] = (float *)take(sizeof(float) * vec_floats(L)) + L.pitch + 32;
        if (assign) { L.bh = v[0]; L.xa = v[1]; L.xb = v[2]; L.r = v[3]; }
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
    const size_t nn0 = (size_t)mg->L[0].nx * mg->L[0].ny;
    {
        int *p = (int *)take(sizeof(int) * nn0), *q = (int *)take(sizeof(int) * nn0), *r = (int *)take(sizeof(int) * nn0);
        if (assign) { mg->free2pp = p; mg->free2node = q; mg->node2free = r; }
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
// reduced DOF -> compact node / padded index (level 0), and the fixed-node mask, from the renumbering of kb_dirichlet_split
__global__ void __launch_bounds__(256) k_mg_free_map(int64_t nn, int nx, int pitch, const int32_t *__restrict__ renum,
                                                     int *__restrict__ free2node, int *__restrict__ free2pp,
                                                     int *__restrict__ node2free, uint8_t *__restrict__ fixed) {
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; p < nn; p += stride) {
        const int f = renum[p];
        fixed[p] = f < 0;
        node2free[p] = f >= 0 ? f : -1;
        if (f >= 0) {
            free2node[f] = (int)p;
            free2pp[f] = (int)(p / nx) * pitch + (int)(p % nx);
        }
    }
}

// level 0: CSR rows of the reduced system -> unit-diagonal stencil coefficients.  An entry outside the 9-point
// neighbourhood of its row raises flag[0].
__global__ void __launch_bounds__(256) k_mg_dia_from_csr(int64_t nfree, const int32_t *__restrict__ indptr,
                                                         const int32_t *__restrict__ indices, const double *__restrict__ vals,
                                                         const int *__restrict__ free2node, int nx, int pitch, Coef8 C,
                                                         float *__restrict__ d, int *__restrict__ flag) {
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    bool bad = false;
    for (int64_t f = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; f < nfree; f += stride) {
        const int p = free2node[f], i = p % nx, j = p / nx;
        const int64_t pp = (int64_t)j * pitch + i;
        const int e0 = indptr[f], e1 = indptr[f + 1];
        double dg = 0.0;
        for (int e = e0; e < e1; ++e)
            if (indices[e] == (int)f) dg += vals[e];
        if (!(dg != 0.0) || !isfinite(dg)) dg = 1.0;
        d[pp] = (float)dg;
        const double inv = 1.0 / dg;
        for (int e = e0; e < e1; ++e) {
            const int c = indices[e];
            if (c == (int)f) continue;
            const int q = free2node[c], dx = q % nx - i, dy = q / nx - j;
            if (dx < -1 || dx > 1 || dy < -1 || dy > 1) { bad = true; continue; }
            const int k = (dy + 1) * 3 + (dx + 1);
            C.c[k > 4 ? k - 1 : k][pp] = (float)(vals[e] * inv);
        }
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
                                                         ElemOpts o, Coef8 C, float *__restrict__ d) {
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
        d[pp] = (float)dg;
        const double inv = 1.0 / dg;
#pragma unroll
        for (int k = 0; k < 9; ++k) {
            if (k == 4) continue;
            const int qi = i + (k % 3) - 1, qj = j + (k / 3) - 1;
            float cv = 0.f;
            if (qi >= 0 && qj >= 0 && qi < nx && qj < ny && !fixed[(int64_t)qj * nx + qi]) cv = (float)(st[k] * inv);
            C.c[k > 4 ? k - 1 : k][pp] = cv;
        }
    }
}

// coarsest level: dense [Ahat | I] (fp64, n x 2n) from the stencil; rows of non-unknowns are identity rows
__global__ void __launch_bounds__(256) k_mg_dense_build(int nx, int ny, int pitch, Coef8 C, const float *__restrict__ d,
                                                        double *__restrict__ W) {
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
        if (qi >= 0 && qj >= 0 && qi < nx && qj < ny) row[qj * nx + qi] = (double)C.c[k8][pp];
    }
}

// Gauss-Jordan without pivoting (rows are scaled to a unit diagonal and the rediscretised operators are diagonally
// dominant); a vanishing / non-finite pivot raises flag[1].  One CTA; step k touches the window [k, n+k] of every row.
__global__ void __launch_bounds__(1024) k_mg_dense_invert(int n, double *__restrict__ W, float *__restrict__ ainv,
                                                          int *__restrict__ flag) {
    __shared__ double s_row[MG_DENSE_MAX + 1];
    __shared__ double s_fac[MG_DENSE_MAX];
    const int m = 2 * n, tid = threadIdx.x, T = blockDim.x;
    for (int k = 0; k < n; ++k) {
        const double piv = W[(size_t)k * m + k];
        if (!(fabs(piv) > 1.0e-12) || !isfinite(piv)) { if (tid == 0) flag[1] = 1; return; }   // uniform over the CTA
        const double inv = 1.0 / piv;
        for (int c = tid; c <= n; c += T) s_row[c] = W[(size_t)k * m + k + c] * inv;
        for (int r = tid; r < n; r += T) s_fac[r] = (r == k) ? 0.0 : W[(size_t)r * m + k];
        __syncthreads();
        for (int c = tid; c <= n; c += T) W[(size_t)k * m + k + c] = s_row[c];
        const int width = n + 1;
        for (int64_t e = tid; e < (int64_t)n * width; e += T) {
            const int r = (int)(e / width), c = (int)(e % width);
            const double f = s_fac[r];
            if (f != 0.0) W[(size_t)r * m + k + c] -= f * s_row[c];
        }
        __syncthreads();
    }
    for (int64_t e = tid; e < (int64_t)n * n; e += T) ainv[e] = (float)W[(size_t)(e / n) * m + n + (e % n)];
}

// ------------------------------------------------------------------------------------------------ V-cycle kernels
// stage the fp64 reduced vector on the level-0 grid: bh = v / d (0 at non-unknowns), first sweep from zero x = omega bh.
// With q != NULL the vector is first updated in place, v -= alpha[0] q (BiCGSTAB's s = r - alpha v).
__global__ void __launch_bounds__(256) k_mg_entry(int nx, int ny, int pitch, const int *__restrict__ node2free,
                                                  const float *__restrict__ d, double *__restrict__ v,
                                                  const double *__restrict__ alpha, const double *__restrict__ q,
                                                  float *__restrict__ bh, float *__restrict__ xa, float omega,
                                                  const int *__restrict__ done) {
    kb_pdl_wait();
    if (done != nullptr && *done != 0) return;
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= pitch) return;
    const double a = (q != nullptr) ? alpha[0] : 0.0;
    for (int j = blockIdx.y; j < ny; j += gridDim.y) {
        const int64_t pp = (int64_t)j * pitch + i;
        float b = 0.f;
        if (i < nx) {
            const int f = node2free[(int64_t)j * nx + i];
            if (f >= 0) {
                double vv = v[f];
                if (q != nullptr) { vv -= a * __ldcs(q + f); v[f] = vv; }
                b = (float)(vv / (double)d[pp]);
            }
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

// MODE 0: damped-Jacobi sweep  out = x + omega (bh - x - sum_k c_k x_k)     MODE 1: residual  out = d (bh - x - sum_k c_k x_k)
// Flat over the padded grid, four nodes per thread (rows are padded to 32 floats, so every float4 is aligned; the guard
// rows and the zero coefficients of padding / boundary nodes make index arithmetic and branches unnecessary).
template <int MODE>
__global__ void __launch_bounds__(256) k_mg_sweep(int64_t np4, int pitch, Coef8 C, const float *__restrict__ d,
                                                  const float *__restrict__ bh, const float *__restrict__ xi,
                                                  float *__restrict__ out, float omega, const int *__restrict__ done) {
    kb_pdl_wait();
    if (done != nullptr && *done != 0) return;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t q = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; q < np4; q += stride) {
        const int64_t p = 4 * q;
        const float *xc = xi + p, *xd = xc - pitch, *xu = xc + pitch;
        float m[3][6];
        {
            const float4 a = ld4(xd), b = ld4(xc), c = ld4(xu);
            m[0][1] = a.x; m[0][2] = a.y; m[0][3] = a.z; m[0][4] = a.w; m[0][0] = xd[-1]; m[0][5] = xd[4];
            m[1][1] = b.x; m[1][2] = b.y; m[1][3] = b.z; m[1][4] = b.w; m[1][0] = xc[-1]; m[1][5] = xc[4];
            m[2][1] = c.x; m[2][2] = c.y; m[2][3] = c.z; m[2][4] = c.w; m[2][0] = xu[-1]; m[2][5] = xu[4];
        }
        float cc[8][4];
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            const float4 v = ld4s(C.c[k] + p);
            cc[k][0] = v.x; cc[k][1] = v.y; cc[k][2] = v.z; cc[k][3] = v.w;
        }
        const float4 bv = ld4(bh + p);
        const float bb[4] = {bv.x, bv.y, bv.z, bv.w};
        float dd[4] = {0.f, 0.f, 0.f, 0.f};
        if (MODE == 1) { const float4 dv = ld4(d + p); dd[0] = dv.x; dd[1] = dv.y; dd[2] = dv.z; dd[3] = dv.w; }
        float o[4];
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            float acc = cc[0][e] * m[0][e];
            acc = fmaf(cc[1][e], m[0][e + 1], acc); acc = fmaf(cc[2][e], m[0][e + 2], acc);
            acc = fmaf(cc[3][e], m[1][e], acc); acc = fmaf(cc[4][e], m[1][e + 2], acc);
            acc = fmaf(cc[5][e], m[2][e], acc); acc = fmaf(cc[6][e], m[2][e + 1], acc); acc = fmaf(cc[7][e], m[2][e + 2], acc);
            const float res = bb[e] - m[1][e + 1] - acc;
            o[e] = MODE == 0 ? fmaf(omega, res, m[1][e + 1]) : dd[e] * res;
        }
        *reinterpret_cast<float4 *>(out + p) = make_float4(o[0], o[1], o[2], o[3]);
    }
}

// coarse rhs = P^T (fine residual), scaled to the unit-diagonal system, + the first sweep from zero
__global__ void __launch_bounds__(128) k_mg_restrict(int nxc, int nyc, int pitch_c, int pitch_f, const float *__restrict__ dc,
                                                     const float *__restrict__ rf, const int *__restrict__ jx,
                                                     const float *__restrict__ wx, const int *__restrict__ sx,
                                                     const int *__restrict__ jy, const float *__restrict__ wy,
                                                     const int *__restrict__ sy, float *__restrict__ bhc,
                                                     float *__restrict__ xac, float omega, const int *__restrict__ done) {
    kb_pdl_wait();
    if (done != nullptr && *done != 0) return;
    const int ic = blockIdx.x * blockDim.x + threadIdx.x;
    if (ic >= pitch_c) return;
    int i0 = 0, i1 = 0;
    if (ic < nxc) { i0 = sx[ic > 0 ? ic - 1 : 0]; i1 = sx[ic + 1 < nxc ? ic + 1 : nxc]; }
    for (int jc = blockIdx.y; jc < nyc; jc += gridDim.y) {
        const int64_t ppc = (int64_t)jc * pitch_c + ic;
        float b = 0.f;
        const float dcv = ic < nxc ? dc[ppc] : 0.f;
        if (dcv != 0.f) {
            const int j0 = sy[jc > 0 ? jc - 1 : 0], j1 = sy[jc + 1 < nyc ? jc + 1 : nyc];
            float acc = 0.f;
            for (int jf = j0; jf < j1; ++jf) {
                const float wyv = (jy[jf] == jc) ? 1.f - wy[jf] : wy[jf];
                const float *row = rf + (int64_t)jf * pitch_f;
                float racc = 0.f;
                for (int f = i0; f < i1; ++f) racc = fmaf((jx[f] == ic) ? 1.f - wx[f] : wx[f], row[f], racc);
                acc = fmaf(wyv, racc, acc);
            }
            b = acc / dcv;
        }
        bhc[ppc] = b;
        xac[ppc] = omega * b;
    }
}

// x_f += P x_c at the unknowns of the fine level
__global__ void __launch_bounds__(128) k_mg_prolong(int nxf, int nyf, int pitch_f, int pitch_c, const float *__restrict__ df,
                                                    const float *__restrict__ xc, const int *__restrict__ jx,
                                                    const float *__restrict__ wx, const int *__restrict__ jy,
                                                    const float *__restrict__ wy, float *__restrict__ xf,
                                                    const int *__restrict__ done) {
    kb_pdl_wait();
    if (done != nullptr && *done != 0) return;
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nxf) return;
    const int a = jx[i];
    const float wa = wx[i];
    for (int j = blockIdx.y; j < nyf; j += gridDim.y) {
        const int64_t pp = (int64_t)j * pitch_f + i;
        if (df[pp] == 0.f) continue;
        const float wb = wy[j];
        const float *r0 = xc + (int64_t)jy[j] * pitch_c + a, *r1 = r0 + pitch_c;
        const float lo = fmaf(wa, r0[1] - r0[0], r0[0]), hi = fmaf(wa, r1[1] - r1[0], r1[0]);
        xf[pp] += fmaf(wb, hi - lo, lo);
    }
}

// coarsest level: x = Ainv bh (one CTA, one warp per row)
__global__ void __launch_bounds__(1024) k_mg_dense_solve(int n, int nx, int pitch, const float *__restrict__ ainv,
                                                         const float *__restrict__ bh, float *__restrict__ x,
                                                         const int *__restrict__ done) {
    kb_pdl_wait();
    if (done != nullptr && *done != 0) return;
    __shared__ float s_b[MG_DENSE_MAX];
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
    const int64_t np4 = L.np / 4;
    KB_CUDA(kb_launch_pdl(k_mg_sweep<MODE>, dim3(kb_grid(np4, 256, 8)), dim3(256), 0, s, pdl, np4, L.pitch, L.C,
                          (const float *)L.d, (const float *)L.bh, xi, out, omega, done));
    KB_LAUNCH_CHECK();
    return 0;
}

}  // namespace

int64_t kbmg_nfree(const kb_mg *mg) { return mg->nfree; }
bool kbmg_ready(const kb_mg *mg) { return mg->ready; }

int kbmg_apply(kb_mg *mg, double *d_v, double *d_z, const double *d_alpha, const double *d_q, const int *done,
               cudaStream_t s, bool pdl) {
    if (!mg || !mg->ready || !d_v || !d_z) return KB_EINVAL;
    const int nl = mg->nlev;
    const float om = mg->omega;
    float *cur[MG_MAXLEV], *oth[MG_MAXLEV];
    for (int l = 0; l < nl; ++l) { cur[l] = mg->L[l].xa; oth[l] = mg->L[l].xb; }
    int rc;
    {
        const MgLevel &L = mg->L[0];
        KB_CUDA(kb_launch_pdl(k_mg_entry, grid2d(L.pitch, 256, L.ny), dim3(256), 0, s, pdl, L.nx, L.ny, L.pitch,
                              (const int *)mg->node2free, (const float *)L.d, d_v, d_alpha, d_q, L.bh, L.xa, om, done));
        KB_LAUNCH_CHECK();
    }
    for (int l = 0; l + 1 < nl; ++l) {
        const MgLevel &L = mg->L[l], &Cn = mg->L[l + 1];
        for (int k = 1; k < mg->nu; ++k) {
            if ((rc = launch_sweep<0>(L, cur[l], oth[l], om, done, s, pdl))) return rc;
            float *t = cur[l]; cur[l] = oth[l]; oth[l] = t;
        }
        if ((rc = launch_sweep<1>(L, cur[l], L.r, om, done, s, pdl))) return rc;
        KB_CUDA(kb_launch_pdl(k_mg_restrict, grid2d(Cn.pitch, 128, Cn.ny), dim3(128), 0, s, pdl, Cn.nx, Cn.ny, Cn.pitch,
                              L.pitch, (const float *)Cn.d, (const float *)L.r, (const int *)L.jx, (const float *)L.wx,
                              (const int *)L.sx, (const int *)L.jy, (const float *)L.wy, (const int *)L.sy, Cn.bh, Cn.xa,
                              om, done));
        KB_LAUNCH_CHECK();
    }
    {
        const MgLevel &L = mg->L[nl - 1];
        KB_CUDA(kb_launch_pdl(k_mg_dense_solve, dim3(1), dim3(1024), 0, s, pdl, mg->ncoarse, L.nx, L.pitch,
                              (const float *)mg->ainv, (const float *)L.bh, cur[nl - 1], done));
        KB_LAUNCH_CHECK();
    }
    for (int l = nl - 2; l >= 0; --l) {
        const MgLevel &L = mg->L[l], &Cn = mg->L[l + 1];
        KB_CUDA(kb_launch_pdl(k_mg_prolong, grid2d(L.nx, 128, L.ny), dim3(128), 0, s, pdl, L.nx, L.ny, L.pitch, Cn.pitch,
                              (const float *)L.d, (const float *)cur[l + 1], (const int *)L.jx, (const float *)L.wx,
                              (const int *)L.jy, (const float *)L.wy, cur[l], done));
        KB_LAUNCH_CHECK();
        for (int k = 0; k < mg->nu; ++k) {
            if ((rc = launch_sweep<0>(L, cur[l], oth[l], om, done, s, pdl))) return rc;
            float *t = cur[l]; cur[l] = oth[l]; oth[l] = t;
        }
    }
    KB_CUDA(kb_launch_pdl(k_mg_exit, dim3(kb_grid(mg->nfree, 256, 8)), dim3(256), 0, s, pdl, mg->nfree,
                          (const int *)mg->free2pp, (const float *)cur[0], d_z, done));
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
    if (nxn * nyn > (int64_t)INT32_MAX / 2) return KB_ETOOBIG;
    if (coarsest_nodes < 4) coarsest_nodes = 4;
    if (coarsest_nodes > MG_DENSE_MAX) coarsest_nodes = MG_DENSE_MAX;
    kb_mg *mg = new (std::nothrow) kb_mg();
    if (!mg) return KB_EINVAL;
    int nx = (int)nxn, ny = (int)nyn;
    for (;;) {
        MgLevel &L = mg->L[mg->nlev++];
        L.nx = nx; L.ny = ny;
        L.pitch = (nx + 1 + 31) / 32 * 32;
        L.np = (int64_t)ny * L.pitch;
        if ((int64_t)nx * ny <= coarsest_nodes || mg->nlev == MG_MAXLEV) break;
        const int nxc = nx > 3 ? nx / 2 + 1 : nx, nyc = ny > 3 ? ny / 2 + 1 : ny;
        if (nxc == nx && nyc == ny) break;
        nx = nxc; ny = nyc;
    }
    const MgLevel &Lc = mg->L[mg->nlev - 1];
    if ((int64_t)Lc.nx * Lc.ny > MG_DENSE_MAX) { delete mg; return KB_ETOOBIG; }
    mg->ncoarse = Lc.nx * Lc.ny;
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
        nnz < 1 || nu < 1 || !(omega > 0.0 && omega < 2.0))
        return KB_EINVAL;
    if (d_velocity != nullptr && mode != KB_MODE_CONSISTENT) return KB_EINVAL;   // coarse levels: W (x) u.gradN + PG weights
    if (workspace_bytes < carve(mg, nullptr, false)) return KB_EWORKSPACE;
    cudaStream_t s = (cudaStream_t)stream;
    mg->ready = false;
    const bool first = mg->ws_base != d_workspace;      // new workspace: vectors zeroed, transfer tables uploaded
    carve(mg, (char *)d_workspace, true);
    mg->ws_base = d_workspace;
    mg->nu = nu; mg->omega = (float)omega; mg->nfree = nfree; mg->has_velocity = d_velocity != nullptr;
    MgLevel &L0 = mg->L[0];
    const int64_t nn0 = (int64_t)L0.nx * L0.ny;
    if (nfree > nn0) return KB_EINVAL;
    // coefficient arrays and vectors start from zero (padding, guard rows, absent neighbours)
    for (int l = 0; l < mg->nlev; ++l) {
        MgLevel &L = mg->L[l];
        for (int k = 0; k < 8; ++k) KB_CUDA(cudaMemsetAsync(L.C.c[k], 0, sizeof(float) * L.np, s));
        KB_CUDA(cudaMemsetAsync(L.d, 0, sizeof(float) * L.np, s));
        if (first) {
            float *v[4] = {L.bh, L.xa, L.xb, L.r};
            for (int k = 0; k < 4; ++k) KB_CUDA(cudaMemsetAsync(v[k] - L.pitch - 32, 0, sizeof(float) * vec_floats(L), s));
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
    KB_CUDA(cudaMemsetAsync(mg->flag, 0, sizeof(int) * 8, s));
    k_mg_free_map<<<kb_grid(nn0, 256, 8), 256, 0, s>>>(nn0, L0.nx, L0.pitch, d_renum, mg->free2node, mg->free2pp,
                                                       mg->node2free, L0.fixed);
    KB_LAUNCH_CHECK();
    k_mg_dia_from_csr<<<kb_grid(nfree, 256, 8), 256, 0, s>>>(nfree, d_indptr, d_indices, d_vals, mg->free2node, L0.nx,
                                                            L0.pitch, L0.C, L0.d, mg->flag);
    KB_LAUNCH_CHECK();
    const ElemOpts o = kb_make_opts(material, KB_MODE_CONSISTENT, d_velocity != nullptr ? 1 : 0, 0);
    for (int l = 1; l < mg->nlev; ++l) {
        MgLevel &L = mg->L[l];
        const MgLevel &F = mg->L[l - 1];
        const int64_t nn = (int64_t)L.nx * L.ny;
        const double *pf = l == 1 ? d_points : F.pts, *vf = d_velocity == nullptr ? nullptr : (l == 1 ? d_velocity : F.vel);
        k_mg_sample<<<kb_grid(nn, 256, 8), 256, 0, s>>>(L.nx, L.ny, F.nx, F.ny, pf, vf, F.fixed, L.pts, L.vel, L.fixed);
        KB_LAUNCH_CHECK();
        if (d_velocity != nullptr)
            k_mg_assemble_dia<true><<<kb_grid(nn, 128, 16), 128, 0, s>>>(L.nx, L.ny, L.pitch, L.pts, L.vel, L.fixed, o, L.C, L.d);
        else
            k_mg_assemble_dia<false><<<kb_grid(nn, 128, 16), 128, 0, s>>>(L.nx, L.ny, L.pitch, L.pts, L.vel, L.fixed, o, L.C, L.d);
        KB_LAUNCH_CHECK();
    }
    {
        const MgLevel &Lc = mg->L[mg->nlev - 1];
        k_mg_dense_build<<<mg->ncoarse, 256, 0, s>>>(Lc.nx, Lc.ny, Lc.pitch, Lc.C, Lc.d, mg->dense);
        KB_LAUNCH_CHECK();
        k_mg_dense_invert<<<1, 1024, 0, s>>>(mg->ncoarse, mg->dense, mg->ainv, mg->flag);
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
    return kbmg_apply(mg, const_cast<double *>(d_v), d_z, nullptr, nullptr, nullptr, (cudaStream_t)stream, false);
}
