// K9 (native form): the row-block partitioned Krylov solvers and the partitioned theta-method loop with EVERY exchange done
// by the kernels themselves over NVLink / NVSwitch peer memory (peer.cuh) -- no collective library call and no host
// round trip inside an iteration.  Included at the end of krylov.cu (same translation unit: it reuses the SpMV launcher,
// the breakdown bounds and the state-readback helpers of the single-GPU driver).
//
// New capability (the reference is single-process): it replaces nothing in /root/reference, it extends
// LinearSolver.Solve (kiltro/FEMSolver.py:350-364) and TransientSolver (:253-297) to a mesh split into horizontal strips.
//
// One BiCGSTAB iteration = 4 kernels, exactly the single-GPU loop's:
//   SpMV   v = A p        waits (edge CTAs only) for the neighbours' halo rows of p; its last CTA publishes
//                         {<rhat,v>, <v,v>, local <r,r> of the previous update} to every rank          (dot event e1)
//   vec    s = r - a v    every CTA collects e1 (rank-ordered sum: bit-identical on all ranks), tests convergence /
//                         breakdown of the previous iterate ON DEVICE, derives alpha; the first CTAs update the two
//                         boundary mesh rows first, store them straight into the neighbours' halo slots and raise
//                         the neighbours' flags -- a whole kernel before the next SpMV needs them
//   SpMV   t = A s        same, publishes {<s,t>, <t,t>, <rhat,t>, <rhat,s>}                           (dot event e2)
//   vec    x, r, p        collects e2 -> omega, rho, beta; merged update; boundary rows of p pushed early; local <r,r>
// CG = SpMV + 2 vec kernels (2 dot events, 1 halo push).  All launches are programmatic dependent launches.  Because
// every rank adds the same partial sums in the same order, every rank takes the same decisions (done flag, restarts)
// without talking: the host only reads its own device's state every `check_every` iterations.
#pragma once

namespace {

using kbpeer::Ctx;

constexpr int S_RHOB = 12;                // second buffer of rho (the update kernel reads one and writes the other)
constexpr int S_BAD = 13;                 // omega / beta / rho breakdown noticed by the update kernel
constexpr int PV_HALO_CTAS = 8;           // CTAs that own the boundary mesh rows of a pushed vector

struct HaloPush {                          // where the first / last owned mesh row of a produced vector goes
    double *dn_dst, *up_dst;               // halo-above slot of rank-1 / halo-below slot of rank+1 (null at the domain ends)
    unsigned long long *dn_flag, *up_flag; // their flags in the neighbours' control blocks
    unsigned long long seq;
    int64_t nxn, n;
    __device__ __forceinline__ int64_t lo() const { return dn_dst != nullptr ? nxn : 0; }
    __device__ __forceinline__ int64_t hi() const { return up_dst != nullptr ? n - nxn : n; }
    __device__ __forceinline__ void store(int64_t i, double v) const {
        if (dn_dst != nullptr && i < nxn) dn_dst[i] = v;
        if (up_dst != nullptr && i >= n - nxn) up_dst[i - (n - nxn)] = v;
    }
    __device__ __forceinline__ void raise() const {        // the caller has fenced (system scope) after the pushed rows:
        if (dn_flag != nullptr) kbpeer::st_relaxed_sys(dn_flag, seq);   // fence + relaxed stores = releases, without the second
        if (up_flag != nullptr) kbpeer::st_relaxed_sys(up_flag, seq);   // store waiting for the first one's round trip
    }
};

// Fused vector kernel of the partitioned loops.  Body interface:
//   struct S { ...; int stop; }   S prepare() const          thread 0 of every CTA (may wait for the peers' partial sums)
//   struct L;  void load(i, L&) const ;  double apply(i, L, S, double (&acc)[4]) const   -> the value to push (PUSH)
//   void finish(const double (&a)[4]) const                   last CTA, thread 0 (NACC > 0)
//   static constexpr int NACC (0, 2 or 4) ; static constexpr bool PUSH
template <class Body>
__global__ void __launch_bounds__(256) k_pvec(int64_t n, Body body, HaloPush h, double *__restrict__ partials,
                                              unsigned int *__restrict__ tickets, const int *__restrict__ done) {
    kb_pdl_wait();
    if (done != nullptr && *done != 0) return;
    __shared__ double red[32];
    __shared__ int s_last;
    __shared__ typename Body::S s_sc;
    if (threadIdx.x == 0) s_sc = body.prepare();
    __syncthreads();
    const typename Body::S sc = s_sc;
    if (sc.stop) return;                                       // same decision in every CTA (identical sums)
    double acc[4] = {0.0, 0.0, 0.0, 0.0};
    int64_t lo = 0, hi = n;
    if (Body::PUSH) {
        lo = h.lo(); hi = h.hi();
        const int64_t nb = lo + (n - hi);                      // boundary entries: [0, lo) and [hi, n)
        const int hb = (int)gridDim.x < PV_HALO_CTAS ? (int)gridDim.x : PV_HALO_CTAS;
        if (nb > 0 && (int)blockIdx.x < hb) {
            for (int64_t j = (int64_t)blockIdx.x * 256 + threadIdx.x; j < nb; j += (int64_t)hb * 256) {
                const int64_t i = j < lo ? j : hi + (j - lo);
                typename Body::L l;
                body.load(i, l);
                h.store(i, body.apply(i, l, sc, acc));
            }
            __threadfence_system();                            // this thread's peer stores are out ...
            __syncthreads();
            if (threadIdx.x == 0) {                            // ... before the CTA reports; the last of the hb CTAs raises the flags
                const unsigned int t = atomicInc(tickets + 1, (unsigned int)hb - 1u);
                if (t == (unsigned int)hb - 1u) { __threadfence_system(); h.raise(); }
            }
        }
    }
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    int64_t i = lo + (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    for (; i + 3 * stride < hi; i += 4 * stride) {
        typename Body::L l0, l1, l2, l3;
        body.load(i, l0); body.load(i + stride, l1); body.load(i + 2 * stride, l2); body.load(i + 3 * stride, l3);
        body.apply(i, l0, sc, acc); body.apply(i + stride, l1, sc, acc);
        body.apply(i + 2 * stride, l2, sc, acc); body.apply(i + 3 * stride, l3, sc, acc);
    }
    for (; i < hi; i += stride) {
        typename Body::L l0;
        body.load(i, l0);
        body.apply(i, l0, sc, acc);
    }
    if (Body::NACC > 0) {
#pragma unroll
        for (int k = 0; k < Body::NACC; ++k) acc[k] = block_sum(acc[k], red);
        if (threadIdx.x == 0) {
#pragma unroll
            for (int k = 0; k < Body::NACC; ++k) partials[k * gridDim.x + blockIdx.x] = acc[k];
            __threadfence();
            const unsigned int t = atomicInc(tickets, gridDim.x - 1);
            s_last = (t == gridDim.x - 1);
        }
        __syncthreads();
        if (s_last) {
            __threadfence();
            double a[4] = {0.0, 0.0, 0.0, 0.0};
            for (int k = threadIdx.x; k < (int)gridDim.x; k += 256) {
#pragma unroll
                for (int q = 0; q < Body::NACC; ++q) a[q] += __ldcg(partials + q * gridDim.x + k);
            }
#pragma unroll
            for (int q = 0; q < Body::NACC; ++q) a[q] = block_sum(a[q], red);
            if (threadIdx.x == 0) body.finish(a);
        }
    }
}

struct PStopOnly { int stop; };

// ---- restart / verification sequence
struct PCopyPush {          // dst = src ; boundary rows -> neighbours           (x -> p before the true-residual SpMV)
    const double *src; double *dst;
    typedef PStopOnly S; struct L { double v; };
    static constexpr int NACC = 0; static constexpr bool PUSH = true;
    __device__ S prepare() const { return S{0}; }
    __device__ void load(int64_t i, L &l) const { l.v = src[i]; }
    __device__ double apply(int64_t i, const L &l, const S &, double (&)[4]) const { dst[i] = l.v; return l.v; }
    __device__ void finish(const double (&)[4]) const {}
};
struct PResidInit {         // r = b - q ; p = r ; rhat = r ; publishes {<r,r>, <b,b>, <r,r>_unscaled, <b,b>_unscaled}
    Ctx c; unsigned long long ev; const double *b, *q, *scale; double *r, *p, *rhat, *dr;
    typedef PStopOnly S; struct L { double b, q, s; };
    static constexpr int NACC = 4; static constexpr bool PUSH = true;
    __device__ S prepare() const { return S{0}; }
    __device__ void load(int64_t i, L &l) const { l.b = b[i]; l.q = q[i]; l.s = scale != nullptr ? scale[i] : 1.0; }
    __device__ double apply(int64_t i, const L &l, const S &, double (&a)[4]) const {
        const double ri = l.b - l.q;
        r[i] = ri; p[i] = ri;
        if (rhat != nullptr) rhat[i] = ri;
        a[0] += ri * ri; a[1] += l.b * l.b;
        const double ru = ri / l.s, bu = l.b / l.s;
        a[2] += ru * ru; a[3] += bu * bu;
        return ri;
    }
    __device__ void finish(const double (&a)[4]) const {
        dr[0] = a[0];
        __threadfence_system();
        kbpeer::post_dots(c, ev, a[0], a[1], a[2], a[3]);
    }
};
// one thread: sums the ranks' partial norms and initialises the scalars of a (re)started cycle
__global__ void k_peer_init(Ctx c, unsigned long long ev, double *sc, int *done, double rtol2, int unscaled, double slack) {
    kb_pdl_wait();
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    double d[4];
    const bool ok = kbpeer::collect_dots(c, ev, d);
    const double rr = d[0], bb = d[1];
    sc[S_RHO] = rr; sc[S_RHOB] = rr; sc[S_RES2] = rr; sc[S_AUX0] = bb; sc[S_R0N2] = rr; sc[S_BAD] = 0.0;
    sc[S_ALPHA] = 0.0; sc[S_OMEGA] = 1.0; sc[S_BETA] = 0.0;
    sc[S_URES2] = d[2]; sc[S_UB2] = d[3];
    double tol2 = rtol2 * bb;
    int dn = 0;
    if (!ok || !isfinite(rr)) dn = 2;
    else if (unscaled) {                                       // CG: the verdict is taken on the unscaled residual (see krylov_solve)
        const double ru2 = d[2], tolu2 = rtol2 * d[3];
        if (!isfinite(ru2)) dn = 2;
        else if (!(ru2 > tolu2)) dn = 1;
        else { const double t = slack * rr * (tolu2 / ru2); if (t < tol2) tol2 = t; }
    } else if (!(rr > tol2)) dn = 1;
    sc[S_TOL2] = tol2;
    done[0] = dn; done[1] = 0; done[2] = dn;                  // done[2]: verdict of this (re)start alone (the loop never writes it)
}
__global__ void __launch_bounds__(256) k_pfill(double *__restrict__ p, int64_t n, double v) {
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) p[i] = v;
}
// a rank leaves this kernel only after every rank entered it (used between cycles / solves; one thread)
__global__ void k_peer_barrier(Ctx c, unsigned long long ev) {
    kb_pdl_wait();
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    __threadfence_system();
    kbpeer::post_dots(c, ev, 0.0, 0.0, 0.0, 0.0);
    double d[4];
    kbpeer::collect_dots(c, ev, d);
}

// ---- BiCGSTAB
struct FinPostDots {        // SpMV epilogue: publish this rank's partial sums (+ the local <r,r> of the last update) to every rank
    Ctx c; unsigned long long ev; const double *extra;
    __device__ void operator()(double a0, double a1, double a2 = 0.0, double a3 = 0.0) const {
        if (extra != nullptr) a2 = extra[0];
        kbpeer::post_dots(c, ev, a0, a1, a2, a3);
    }
};
struct PBiS {               // after v = A p: convergence / breakdown of the previous iterate, alpha, s = r - alpha v (pushed)
    Ctx c; unsigned long long ev; const double *rho, *v; double *r, *sc; int *done; const int *iter;
    struct S { double a; int stop; }; struct L { double r, v; };
    static constexpr int NACC = 0; static constexpr bool PUSH = true;
    __device__ S prepare() const {
        double d[4];
        const bool ok = kbpeer::collect_dots(c, ev, d);        // {<rhat,v>, <v,v>, <r,r> of the iterate, -}
        const double rr = d[2];
        int dn = 0; double why = 0.0, a = 0.0;
        if (!ok) { dn = 2; why = 9.0; }
        else if (!(rr > sc[S_TOL2]) && isfinite(rr)) dn = 1;
        else if (!isfinite(rr) || sc[S_BAD] != 0.0 || rr > KB_GROWTH2_MAX * sc[S_R0N2]) { dn = 2; why = sc[S_BAD] != 0.0 ? 3.0 : (isfinite(rr) ? 5.0 : 6.0); }
        else {
            a = rho[0] / d[0];
            if (!isfinite(a) || d[0] == 0.0) { dn = 2; why = 1.0; }
            else if (a * a * d[1] > KB_GROWTH2_MAX * rr) { dn = 2; why = 2.0; }     // spike guard, before the update
        }
        if (blockIdx.x == 0) {
            sc[S_RES2] = rr;
            if (dn != 0) { done[0] = dn; done[1] = iter[0]; sc[S_WHY] = why; }
            else sc[S_ALPHA] = a;
        }
        return S{a, dn != 0 ? 1 : 0};
    }
    __device__ void load(int64_t i, L &l) const { l.r = r[i]; l.v = __ldcs(v + i); }
    __device__ double apply(int64_t i, const L &l, const S &s, double (&)[4]) const {
        const double sn = l.r - s.a * l.v;
        r[i] = sn;
        return sn;
    }
    __device__ void finish(const double (&)[4]) const {}
};
struct PBiUpdateP {         // after t = A s: omega, rho, beta ; x += a p + om s ; r = s - om t ; p = r + beta (p - om v) (pushed)
    Ctx c; unsigned long long ev; const double *rho_in; double *rho_out; const double *t, *v; double *x, *r, *p, *sc, *dr; int *iter;
    struct S { double a, om, b; int stop; }; struct L { double x, p, s, t, v; };
    static constexpr int NACC = 2; static constexpr bool PUSH = true;
    __device__ S prepare() const {
        double d[4];
        const bool ok = kbpeer::collect_dots(c, ev, d);        // {<s,t>, <t,t>, <rhat,t>, <rhat,s>}
        const double om = d[1] > 0.0 ? d[0] / d[1] : 0.0;
        const double rho_new = d[3] - om * d[2];
        const double a = sc[S_ALPHA];
        const double b = (rho_new / rho_in[0]) * (a / om);
        if (blockIdx.x == 0) {
            rho_out[0] = rho_new; sc[S_OMEGA] = om; sc[S_BETA] = b;
            sc[S_BAD] = (ok && isfinite(b) && rho_new != 0.0 && om != 0.0) ? 0.0 : 1.0;   // judged by the next kernel, unless converged
        }
        return S{a, om, b, 0};
    }
    __device__ void load(int64_t i, L &l) const { l.x = x[i]; l.p = p[i]; l.s = r[i]; l.t = __ldcs(t + i); l.v = __ldcs(v + i); }
    __device__ double apply(int64_t i, const L &l, const S &s, double (&acc)[4]) const {
        x[i] = l.x + (s.a * l.p + s.om * l.s);
        const double ri = l.s - s.om * l.t;
        r[i] = ri;
        const double pn = ri + s.b * (l.p - s.om * l.v);
        p[i] = pn;
        acc[0] += ri * ri;
        return pn;
    }
    __device__ void finish(const double (&a)[4]) const { dr[0] = a[0]; ++iter[0]; }
};

// ---- CG
struct PCgUpdate {          // after q = A p: alpha = rho / <p,q> ; x += a p ; r -= a q ; publishes the local <r,r>
    Ctx c; unsigned long long ev_in, ev_out; const double *rho, *p, *q; double *x, *r, *sc; int *done, *iter;
    struct S { double a; int stop; }; struct L { double x, p, r, q; };
    static constexpr int NACC = 2; static constexpr bool PUSH = false;
    __device__ S prepare() const {
        double d[4];
        const bool ok = kbpeer::collect_dots(c, ev_in, d);
        const double a = rho[0] / d[0];
        const int bad = (!ok || !isfinite(a) || d[0] == 0.0) ? 1 : 0;
        if (blockIdx.x == 0) {
            if (bad) { done[0] = 2; done[1] = iter[0]; sc[S_WHY] = ok ? 1.0 : 9.0; }
            else sc[S_ALPHA] = a;
        }
        return S{a, bad};
    }
    __device__ void load(int64_t i, L &l) const { l.x = x[i]; l.p = p[i]; l.r = r[i]; l.q = __ldcs(q + i); }
    __device__ double apply(int64_t i, const L &l, const S &s, double (&acc)[4]) const {
        x[i] = l.x + s.a * l.p;
        const double ri = l.r - s.a * l.q;
        r[i] = ri;
        acc[0] += ri * ri;
        return ri;
    }
    __device__ void finish(const double (&a)[4]) const {
        ++iter[0];
        __threadfence_system();
        kbpeer::post_dots(c, ev_out, a[0], 0.0, 0.0, 0.0);
    }
};
struct PCgP {               // summed <r,r>: convergence test, beta ; p = r + beta p (pushed)
    Ctx c; unsigned long long ev; const double *rho_in; double *rho_out; const double *r; double *p, *sc; int *done; const int *iter;
    struct S { double b; int stop; }; struct L { double r, p; };
    static constexpr int NACC = 0; static constexpr bool PUSH = true;
    __device__ S prepare() const {
        double d[4];
        const bool ok = kbpeer::collect_dots(c, ev, d);
        const double rr = d[0];
        int dn = 0;
        if (!ok || !isfinite(rr)) dn = 2;
        else if (!(rr > sc[S_TOL2])) dn = 1;
        const double b = rr / rho_in[0];
        if (blockIdx.x == 0) {
            sc[S_RES2] = rr;
            if (dn != 0) { done[0] = dn; done[1] = iter[0]; }
            else { rho_out[0] = rr; sc[S_BETA] = b; }
        }
        return S{b, dn != 0 ? 1 : 0};
    }
    __device__ void load(int64_t i, L &l) const { l.r = r[i]; l.p = p[i]; }
    __device__ double apply(int64_t i, const L &l, const S &s, double (&)[4]) const {
        const double pn = l.r + s.b * l.p;
        p[i] = pn;
        return pn;
    }
    __device__ void finish(const double (&)[4]) const {}
};

// ---- transient step
struct PRhs {               // rhs = (K T + Fm) [* scale]      (scale: symmetric Jacobi scaling of the inner system)
    const double *kt, *fm, *scale; double *rhs;
    typedef PStopOnly S; struct L { double k, f, s; };
    static constexpr int NACC = 0; static constexpr bool PUSH = false;
    __device__ S prepare() const { return S{0}; }
    __device__ void load(int64_t i, L &l) const { l.k = kt[i]; l.f = fm[i]; l.s = scale != nullptr ? scale[i] : 1.0; }
    __device__ double apply(int64_t i, const L &l, const S &, double (&)[4]) const { const double v = (l.k + l.f) * l.s; rhs[i] = v; return v; }
    __device__ void finish(const double (&)[4]) const {}
};
struct PStep {              // y = yh * scale ; T = fixed ? T_D : T - dt y   (boundary rows of T pushed)
    const int32_t *renum; const double *td, *yh, *scale; double *T; double dt;
    typedef PStopOnly S; struct L { double T, y, s; int r; };
    static constexpr int NACC = 0; static constexpr bool PUSH = true;
    __device__ S prepare() const { return S{0}; }
    __device__ void load(int64_t i, L &l) const { l.T = T[i]; l.y = yh[i]; l.s = scale[i]; l.r = renum[i]; }
    __device__ double apply(int64_t i, const L &l, const S &, double (&)[4]) const {
        const double v = l.r >= 0 ? l.T - dt * (l.y * l.s) : td[-1 - l.r];
        T[i] = v;
        return v;
    }
    __device__ void finish(const double (&)[4]) const {}
};

// ------------------------------------------------------------------------------------------------ host side
struct PeerWork {
    double *rhat, *v, *t, *partials, *sc, *dr;
    unsigned int *tickets;
    int *done, *iter;
};

size_t peer_carve(PeerWork *w, char *base, int64_t n, int64_t nnz) {
    const int64_t nblk = num_blocks(nnz);
    const int64_t npart = 4 * (int64_t)kb_num_sms() * 8 + 4 * nblk + 64;
    size_t off = 0;
#define TAKE(field, type, count)                                  \
    do {                                                          \
        if (w) w->field = (type *)(base + off);                   \
        off += align256(sizeof(type) * (size_t)(count));          \
    } while (0)
    TAKE(rhat, double, n + 1); TAKE(v, double, n + 1); TAKE(t, double, n + 1);
    TAKE(partials, double, npart);
    TAKE(sc, double, S_COUNT);
    TAKE(dr, double, 4);
    TAKE(tickets, unsigned int, 4);
    TAKE(done, int, 4);
    TAKE(iter, int, 4);
#undef TAKE
    return off;
}

struct PeerEnv {            // everything a launch needs, resolved once per call
    kb_peer_part *pp;
    kb_peer_op op;
    Ctx c;
    int64_t n, nxn, xoff;   // owned rows, rows per mesh row, offset of the SpMV's x argument relative to the vector base
    cudaStream_t s;
    PeerWork w;
};

Ctx make_ctx(const kb_peer_ctx *h) {
    Ctx c;
    c.P = h->nranks; c.rank = h->rank;
    for (int r = 0; r < kbpeer::MAXP; ++r) c.ctrl[r] = (unsigned long long *)h->ctrl[r];
    return c;
}

// With lazy module loading (the CUDA 12 default) the FIRST launch of a kernel may have to wait for running kernels -- fatal
// when the running kernel is itself waiting for a peer whose next kernel has not been loaded yet (ranks sharing one
// process / device in the tests).  Every kernel of the partitioned loops is therefore loaded before the first launch.
template <class F> void preload_one(F *f) { cudaFuncAttributes a; (void)cudaFuncGetAttributes(&a, f); }
void peer_preload() {
    static bool done = false;
    if (done) return;
    preload_one(k_pvec<PCopyPush>); preload_one(k_pvec<PResidInit>); preload_one(k_pvec<PBiS>); preload_one(k_pvec<PBiUpdateP>);
    preload_one(k_pvec<PCgUpdate>); preload_one(k_pvec<PCgP>); preload_one(k_pvec<PRhs>); preload_one(k_pvec<PStep>);
    preload_one(k_peer_init); preload_one(k_peer_barrier); preload_one(k_reset_state); preload_one(k_pfill);
    preload_one(k_spmv_pat<0, NoFinal>); preload_one(k_spmv_pat<1, FinPostDots>); preload_one(k_spmv_pat<2, FinPostDots>);
    preload_one(k_spmv_tma<0, NoFinal, short>); preload_one(k_spmv_tma<1, FinPostDots, short>); preload_one(k_spmv_tma<2, FinPostDots, short>);
    preload_one(k_spmv_tma<0, NoFinal, int>); preload_one(k_spmv_tma<1, FinPostDots, int>); preload_one(k_spmv_tma<2, FinPostDots, int>);
    preload_one(k_spmv<0, NoFinal>); preload_one(k_spmv<1, FinPostDots>); preload_one(k_spmv<2, FinPostDots>);
    // the launcher's one-off shared-memory opt-ins, for the same reason
    cudaFuncSetAttribute(k_spmv_pat<0, NoFinal>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)pat_smem_bytes());
    cudaFuncSetAttribute(k_spmv_pat<1, FinPostDots>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)pat_smem_bytes());
    cudaFuncSetAttribute(k_spmv_pat<2, FinPostDots>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)pat_smem_bytes());
    cudaFuncSetAttribute(k_spmv_tma<0, NoFinal, short>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)tma_smem_bytes<short>());
    cudaFuncSetAttribute(k_spmv_tma<1, FinPostDots, short>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)tma_smem_bytes<short>());
    cudaFuncSetAttribute(k_spmv_tma<2, FinPostDots, short>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)tma_smem_bytes<short>());
    cudaFuncSetAttribute(k_spmv_tma<0, NoFinal, int>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)tma_smem_bytes<int>());
    cudaFuncSetAttribute(k_spmv_tma<1, FinPostDots, int>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)tma_smem_bytes<int>());
    cudaFuncSetAttribute(k_spmv_tma<2, FinPostDots, int>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)tma_smem_bytes<int>());
    (void)cudaGetLastError();
    done = true;
}

int peer_check_args(const kb_peer_part *pp, const kb_peer_op *op) {
    peer_preload();
    if (!pp || !op) return KB_EINVAL;
    const kb_peer_ctx &x = pp->ctx;
    if (x.nranks < 1 || x.nranks > kbpeer::MAXP || x.rank < 0 || x.rank >= x.nranks) return KB_EINVAL;
    for (int r = 0; r < x.nranks; ++r) if (!x.ctrl[r]) return KB_EINVAL;
    if (pp->n_own <= 0 || pp->nxn <= 0 || pp->own_off < 0 || pp->n_loc < pp->own_off + pp->n_own) return KB_EINVAL;
    if (pp->n_own >= INT32_MAX || op->nnz_own >= INT32_MAX) return KB_ETOOBIG;
    const bool dn = x.rank > 0, up = x.rank < x.nranks - 1;
    if ((dn || up) && pp->n_own < 2 * pp->nxn) return KB_EINVAL;          // at least two mesh rows per rank
    for (int k = 0; k < kbpeer::NVEC; ++k) {
        if (!pp->d_vec[k]) return KB_EINVAL;
        if (dn != (pp->d_dn_dst[k] != nullptr) || up != (pp->d_up_dst[k] != nullptr)) return KB_EINVAL;
    }
    if (!op->d_indptr_own || !op->d_indices || !op->d_vals || !op->d_rowblk || op->nnz_own < 0) return KB_EINVAL;
    return 0;
}

PeerEnv make_env(kb_peer_part *pp, const kb_peer_op *op, cudaStream_t s) {
    PeerEnv e;
    e.pp = pp; e.op = *op; e.c = make_ctx(&pp->ctx);
    e.n = pp->n_own; e.nxn = pp->nxn; e.s = s;
    // the 16-bit / row-pattern SpMV forms address x relative to the first row of the view
    e.xoff = (op->d_cols16 != nullptr || op->d_row_pat_own != nullptr) ? pp->own_off : 0;
    return e;
}

HaloPush make_push(const PeerEnv &e, int vec, unsigned long long seq) {
    HaloPush h;
    const kb_peer_ctx &x = e.pp->ctx;
    h.dn_dst = (double *)e.pp->d_dn_dst[vec]; h.up_dst = (double *)e.pp->d_up_dst[vec];
    h.seq = seq; h.nxn = e.nxn; h.n = e.n;
    // rank-1 sees us as its upper neighbour (side 1); rank+1 sees us as its lower neighbour (side 0)
    h.dn_flag = h.dn_dst ? (unsigned long long *)x.ctrl[x.rank - 1] + kbpeer::W_HALOFLAG + vec * 2 + 1 : nullptr;
    h.up_flag = h.up_dst ? (unsigned long long *)x.ctrl[x.rank + 1] + kbpeer::W_HALOFLAG + vec * 2 + 0 : nullptr;
    return h;
}
HaloPush no_push(const PeerEnv &e) {
    HaloPush h;
    h.dn_dst = h.up_dst = nullptr; h.dn_flag = h.up_flag = nullptr; h.seq = 0; h.nxn = e.nxn; h.n = e.n;
    return h;
}

double *own_of(const PeerEnv &e, int vec) { return (double *)e.pp->d_vec[vec] + e.pp->own_off; }

template <class Body>
int launch_pvec(const PeerEnv &e, Body body, HaloPush h, const int *done) {
    const int grid = kb_grid(kb_cdiv(e.n, 4), 256, 8);
    KB_CUDA(kb_launch_pdl(k_pvec<Body>, dim3(grid), dim3(256), 0, e.s, true, e.n, body, h, e.w.partials, e.w.tickets, done));
    KB_LAUNCH_CHECK();
    return 0;
}

// y = A x(vec) on the owned rows; waits for the halo rows of vector `vec` at sequence `seq` (edge CTAs only)
template <int DOTS, class Final>
int peer_spmv(const PeerEnv &e, const double *vals, int vec, unsigned long long seq, double *y, const double *w, const double *w2,
              const int *done, Final fin) {
    kbpeer::Wait wt{nullptr, nullptr, seq, e.c.ctrl[e.c.rank] + kbpeer::W_ERROR, (int)e.nxn};
    if (e.c.rank > 0) wt.f0 = e.c.ctrl[e.c.rank] + kbpeer::W_HALOFLAG + vec * 2 + 0;
    if (e.c.rank < e.c.P - 1) wt.f1 = e.c.ctrl[e.c.rank] + kbpeer::W_HALOFLAG + vec * 2 + 1;
    const double *x = (const double *)e.pp->d_vec[vec] + e.xoff;
    return launch_spmv_raw<DOTS, Final>(e.n, e.op.nnz_own, e.op.d_indptr_own, e.op.d_indices, vals, e.op.d_rowblk, e.w.partials,
                                        e.w.tickets, x, y, w, done, fin, e.s, e.op.d_cols16, w2, wt, e.op.d_row_pat_own,
                                        e.op.d_pat_tab, true);
}

int peer_read_state(const PeerEnv &e, HostState *h, unsigned long long *err) {
    KB_CUDA(cudaMemcpyAsync(h->sc, e.w.sc, sizeof(double) * S_COUNT, cudaMemcpyDeviceToHost, e.s));
    KB_CUDA(cudaMemcpyAsync(h->done, e.w.done, sizeof(int) * 4, cudaMemcpyDeviceToHost, e.s));
    KB_CUDA(cudaMemcpyAsync(h->iter, e.w.iter, sizeof(int) * 4, cudaMemcpyDeviceToHost, e.s));
    KB_CUDA(cudaMemcpyAsync(err, e.c.ctrl[e.c.rank] + kbpeer::W_ERROR, 8, cudaMemcpyDeviceToHost, e.s));
    KB_CUDA(cudaStreamSynchronize(e.s));
    return 0;
}

// (re)start: q = A x (halo of x through vector 0), r = b - q, p = rhat = r, global norms -> scalars / done flag
template <int METHOD>
int peer_restart(PeerEnv &e, const double *vals, const double *b, const double *x, const double *scale, double rtol2,
                 double slack) {
    kb_peer_part *pp = e.pp;
    int rc;
    k_peer_barrier<<<1, 32, 0, e.s>>>(e.c, ++pp->ev);
    KB_LAUNCH_CHECK();
    double *p = own_of(e, 0), *r = own_of(e, 1);
    // x travels through vector 1 (r), NOT vector 0: the residual kernel below pushes the new p into the neighbours' halo
    // slots of vector 0 at once, and nothing orders that push after the neighbours' SpMV of this sequence -- a faster
    // rank would overwrite halo entries its neighbour is still reading.  (Inside the iteration two pushes of the same
    // vector are always separated by a dot event, which every rank posts only after its own SpMV.)
    if ((rc = launch_pvec(e, PCopyPush{x, r}, make_push(e, 1, ++pp->hseq[1]), nullptr))) return rc;
    if ((rc = peer_spmv<0>(e, vals, 1, pp->hseq[1], e.w.v, nullptr, nullptr, nullptr, NoFinal{}))) return rc;
    const unsigned long long ev = ++pp->ev;
    if ((rc = launch_pvec(e, PResidInit{e.c, ev, b, e.w.v, scale, r, p, METHOD == 1 ? e.w.rhat : nullptr, e.w.dr},
                          make_push(e, 0, ++pp->hseq[0]), nullptr))) return rc;
    k_peer_init<<<1, 32, 0, e.s>>>(e.c, ev, e.w.sc, e.w.done, rtol2, METHOD == 0 ? 1 : 0, slack);
    KB_LAUNCH_CHECK();
    return 0;
}

// Partitioned Krylov driver on the (already Jacobi-scaled) operator `vals`: METHOD 0 = CG, 1 = BiCGSTAB.  b, x: owned
// parts (x: initial guess in, solution out, both of the scaled system).  scale (CG): owned part of D^-1/2 for the
// unscaled verdict.  first_batch > 0: iterations enqueued before the first look at the state (time loops know how many
// the previous step took).  Same restart / verdict rules as krylov_solve.
template <int METHOD>
int peer_krylov(PeerEnv &e, const double *vals, const double *b, double *x, const double *scale, double rtol, int maxiter,
                int check_every, int first_batch, kb_solve_info *info) {
    kb_peer_part *pp = e.pp;
    const int64_t launches0 = g_kb_launches;
    memset(info, 0, sizeof(*info));
    if (check_every < 1) check_every = 1;
    if (maxiter < 1) maxiter = 1;
    int rc;
    k_reset_state<<<1, 32, 0, e.s>>>(e.w.tickets, e.w.done, e.w.iter, e.w.sc);
    KB_LAUNCH_CHECK();
    double *p = own_of(e, 0), *r = own_of(e, 1);
    double *q = e.w.v, *t = e.w.t, *rhat = e.w.rhat;
    HostState hs;
    unsigned long long perr = 0;
    int total_iters = 0, status = KB_ENOTCONV, restarts = 0;
    const double rtol2 = rtol * rtol;
    double prev_true = INFINITY;
    bool recurrence_converged = false, attainable = false, verified = false;
    auto res2_of = [&](const HostState &h) { return METHOD == 0 ? h.sc[S_URES2] : h.sc[S_RES2]; };
    auto b2_of = [&](const HostState &h) { return METHOD == 0 ? h.sc[S_UB2] : h.sc[S_AUX0]; };
    for (int restart = 0; restart < 16 && total_iters < maxiter; ++restart) {
        restarts = restart;
        if ((rc = peer_restart<METHOD>(e, vals, b, x, scale, rtol2, restart == 0 ? 1.0 : 0.25))) return rc;
        // time loops (first_batch > 0) enqueue the first batch of the first cycle right behind the (re)start and look at
        // both together: one host round trip less per step (the batch's kernels are no-ops when the start already converged)
        const bool fused = first_batch > 0 && restart == 0;
        if (!fused) {
            if ((rc = peer_read_state(e, &hs, &perr))) return rc;
            if (perr) return KB_EPEER;
            verified = true;                                   // hs holds the true residual of the current x
            // S_RES2 at a (re)start is the true residual (the unscaled one for CG sits in S_URES2)
            const double true_rr = res2_of(hs), tol2 = rtol2 * b2_of(hs);
            if (hs.done[0] == 1) { status = 0; break; }
            if (hs.done[0] == 2) { status = KB_EBREAKDOWN; break; }
            if (restart > 0 && !(true_rr < 0.25 * prev_true) && true_rr <= 100.0 * tol2) { attainable = true; break; }
            if (recurrence_converged && !(true_rr < 0.25 * prev_true) && true_rr <= 1.0e6 * tol2) { attainable = true; break; }
            prev_true = true_rr;
        }
        recurrence_converged = false;
        bool stop = false;
        int start_verdict = 0;
        int cur = 0;                                           // rho lives in sc[S_RHO] (cur = 0) or sc[S_RHOB] (cur = 1)
        int batch_want = (first_batch > 0 && restart == 0) ? first_batch : check_every;
        while (!stop && total_iters < maxiter) {
            const int batch = (maxiter - total_iters) < batch_want ? (maxiter - total_iters) : batch_want;
            batch_want = check_every;
            for (int it = 0; it < batch; ++it) {
                double *rho_in = e.w.sc + (cur ? S_RHOB : S_RHO), *rho_out = e.w.sc + (cur ? S_RHO : S_RHOB);
                if (METHOD == 0) {
                    const unsigned long long e1 = ++pp->ev, e2 = ++pp->ev;
                    if ((rc = peer_spmv<1>(e, vals, 0, pp->hseq[0], q, p, nullptr, e.w.done, FinPostDots{e.c, e1, nullptr}))) return rc;
                    if ((rc = launch_pvec(e, PCgUpdate{e.c, e1, e2, rho_in, p, q, x, r, e.w.sc, e.w.done, e.w.iter}, no_push(e), e.w.done))) return rc;
                    if ((rc = launch_pvec(e, PCgP{e.c, e2, rho_in, rho_out, r, p, e.w.sc, e.w.done, e.w.iter},
                                          make_push(e, 0, ++pp->hseq[0]), e.w.done))) return rc;
                } else {
                    const unsigned long long e1 = ++pp->ev;
                    if ((rc = peer_spmv<1>(e, vals, 0, pp->hseq[0], q, rhat, nullptr, e.w.done, FinPostDots{e.c, e1, e.w.dr}))) return rc;
                    if ((rc = launch_pvec(e, PBiS{e.c, e1, rho_in, q, r, e.w.sc, e.w.done, e.w.iter}, make_push(e, 1, ++pp->hseq[1]), e.w.done))) return rc;
                    const unsigned long long e2 = ++pp->ev;
                    if ((rc = peer_spmv<2>(e, vals, 1, pp->hseq[1], t, r, rhat, e.w.done, FinPostDots{e.c, e2, nullptr}))) return rc;
                    if ((rc = launch_pvec(e, PBiUpdateP{e.c, e2, rho_in, rho_out, t, q, x, r, p, e.w.sc, e.w.dr, e.w.iter},
                                          make_push(e, 0, ++pp->hseq[0]), e.w.done))) return rc;
                }
                cur ^= 1;
            }
            verified = false;
            if ((rc = peer_read_state(e, &hs, &perr))) return rc;
            if (perr) return KB_EPEER;
            total_iters = hs.iter[0];
            if (hs.done[0] != 0) stop = true;
            if (fused && hs.done[2] != 0) { start_verdict = hs.done[2]; break; }   // the start itself converged / broke: x untouched
            if (fused) prev_true = METHOD == 0 ? hs.sc[S_URES2] : hs.sc[S_R0N2];
        }
        if (start_verdict != 0) {
            // norms of the start: S_R0N2 / S_URES2 were written by k_peer_init only
            hs.sc[S_RES2] = hs.sc[S_R0N2];
            verified = true;
            status = start_verdict == 1 ? 0 : KB_EBREAKDOWN;
            break;
        }
        if (getenv("KB_DEBUG_SOLVER"))
            fprintf(stderr, "[kb peer solver r%d] cycle %d: iters %d done %d why %.0f rho %.3e alpha %.3e omega %.3e beta %.3e rr %.3e r0n2 %.3e tol2 %.3e\n",
                    e.c.rank, restart, total_iters, hs.done[0], hs.sc[S_WHY], hs.sc[S_RHO], hs.sc[S_ALPHA], hs.sc[S_OMEGA],
                    hs.sc[S_BETA], hs.sc[S_RES2], hs.sc[S_R0N2], hs.sc[S_TOL2]);
        if (hs.done[0] == 1) recurrence_converged = true;
        else if (hs.done[0] == 2) status = KB_EBREAKDOWN;
        else break;                                            // maxiter exhausted
    }
    if (!verified) {                                           // true residual of the final iterate (not yet known)
        if ((rc = peer_restart<METHOD>(e, vals, b, x, scale, rtol2, 1.0))) return rc;
        if ((rc = peer_read_state(e, &hs, &perr))) return rc;
        if (perr) return KB_EPEER;
    }
    const double res2 = res2_of(hs), b2 = b2_of(hs);
    info->iterations = total_iters;
    info->residual = sqrt(res2);
    info->rhs_norm = sqrt(b2);
    info->restarts = restarts;
    if (b2 == 0.0 && res2 == 0.0) status = 0;
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

// ------------------------------------------------------------------------------------------------ C ABI
// loads every kernel of the partitioned loops now (call it before ranks that share a device start waiting for each other)
extern "C" int kb_peer_preload(void) { peer_preload(); KB_CUDA(cudaDeviceSynchronize()); return 0; }

extern "C" int kb_peer_alloc(size_t bytes, void **d_ptr, unsigned char *h_handle64) {
    if (!d_ptr || !h_handle64 || bytes == 0) return KB_EINVAL;
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
    KB_CUDA(cudaMalloc(d_ptr, bytes));
    KB_CUDA(cudaMemset(*d_ptr, 0, bytes));
    cudaIpcMemHandle_t hd;
    KB_CUDA(cudaIpcGetMemHandle(&hd, *d_ptr));
    memcpy(h_handle64, &hd, 64);
    return 0;
}
extern "C" int kb_peer_open(const unsigned char *h_handle64, void **d_ptr) {
    if (!d_ptr || !h_handle64) return KB_EINVAL;
    cudaIpcMemHandle_t hd;
    memcpy(&hd, h_handle64, 64);
    KB_CUDA(cudaIpcOpenMemHandle(d_ptr, hd, cudaIpcMemLazyEnablePeerAccess));
    return 0;
}
extern "C" int kb_peer_close(void *d_ptr) { if (d_ptr) KB_CUDA(cudaIpcCloseMemHandle(d_ptr)); return 0; }
extern "C" int kb_peer_free(void *d_ptr) { if (d_ptr) KB_CUDA(cudaFree(d_ptr)); return 0; }
extern "C" int kb_peer_error(const kb_peer_ctx *ctx, int64_t *h_err, void *stream) {
    if (!ctx || !h_err) return KB_EINVAL;
    unsigned long long e = 0;
    KB_CUDA(cudaMemcpyAsync(&e, (unsigned long long *)ctx->ctrl[ctx->rank] + kbpeer::W_ERROR, 8, cudaMemcpyDeviceToHost,
                            (cudaStream_t)stream));
    KB_CUDA(cudaStreamSynchronize((cudaStream_t)stream));
    *h_err = (int64_t)e;
    return 0;
}
extern "C" int kb_peer_clear_error(const kb_peer_ctx *ctx, void *stream) {
    if (!ctx) return KB_EINVAL;
    KB_CUDA(cudaMemsetAsync((unsigned long long *)ctx->ctrl[ctx->rank] + kbpeer::W_ERROR, 0, 8, (cudaStream_t)stream));
    return 0;
}

extern "C" size_t kb_peer_krylov_workspace_bytes(int64_t n_own, int64_t nnz_own) { return peer_carve(nullptr, nullptr, n_own, nnz_own); }

static int peer_solve_entry(int method, kb_peer_part *part, const kb_peer_op *op, const double *d_b_own, double *d_x_own,
                            const double *d_scale_own, double rtol, int maxiter, int check_every, void *d_workspace,
                            size_t workspace_bytes, kb_solve_info *h_info, void *stream) {
    int rc = peer_check_args(part, op);
    if (rc) return rc;
    if (!d_b_own || !d_x_own || !d_workspace || !h_info) return KB_EINVAL;
    if (workspace_bytes < peer_carve(nullptr, nullptr, part->n_own, op->nnz_own)) return KB_EWORKSPACE;
    PeerEnv e = make_env(part, op, (cudaStream_t)stream);
    peer_carve(&e.w, (char *)d_workspace, part->n_own, op->nnz_own);
    if (method == 0)
        return peer_krylov<0>(e, op->d_vals, d_b_own, d_x_own, d_scale_own, rtol, maxiter, check_every, 0, h_info);
    return peer_krylov<1>(e, op->d_vals, d_b_own, d_x_own, nullptr, rtol, maxiter, check_every, 0, h_info);
}
extern "C" int kb_peer_cg_solve(kb_peer_part *part, const kb_peer_op *op, const double *d_b_own, double *d_x_own,
                                const double *d_scale_own, double rtol, int maxiter, int check_every, void *d_workspace,
                                size_t workspace_bytes, kb_solve_info *h_info, void *stream) {
    return peer_solve_entry(0, part, op, d_b_own, d_x_own, d_scale_own, rtol, maxiter, check_every, d_workspace,
                            workspace_bytes, h_info, stream);
}
extern "C" int kb_peer_bicgstab_solve(kb_peer_part *part, const kb_peer_op *op, const double *d_b_own, double *d_x_own,
                                      double rtol, int maxiter, int check_every, void *d_workspace, size_t workspace_bytes,
                                      kb_solve_info *h_info, void *stream) {
    return peer_solve_entry(1, part, op, d_b_own, d_x_own, nullptr, rtol, maxiter, check_every, d_workspace, workspace_bytes,
                            h_info, stream);
}

// Partitioned theta-method loop (same scheme as kb_transient_run).  T lives in the region's vector 2 (full local vector;
// its halo rows are pushed by the update kernel of every step).  d_Kff_vals: Kff on the op's pattern (unscaled);
// op->d_vals: the Jacobi-scaled inner operator  (M + theta dt Kff) D^-1  or  D^-1/2 (.) D^-1/2 ; d_scale_own: D^-1 or D^-1/2.
extern "C" size_t kb_peer_transient_workspace_bytes(int64_t n_own, int64_t nnz_own) {
    return peer_carve(nullptr, nullptr, n_own, nnz_own) + 3 * align256(sizeof(double) * (size_t)(n_own + 1));
}
extern "C" int kb_peer_transient_run(kb_peer_part *part, const kb_peer_op *op, const double *d_Kff_vals, int a_symmetric,
                                     const double *d_scale_own, const double *d_Fm_own, const int32_t *d_renum_own,
                                     const double *d_applied, double dt, int nsteps, double rtol, int maxiter,
                                     void *d_workspace, size_t workspace_bytes, kb_solve_info *h_info, void *stream) {
    int rc = peer_check_args(part, op);
    if (rc) return rc;
    if (!d_Kff_vals || !d_scale_own || !d_Fm_own || !d_renum_own || !d_workspace || !h_info || nsteps < 0) return KB_EINVAL;
    if (workspace_bytes < kb_peer_transient_workspace_bytes(part->n_own, op->nnz_own)) return KB_EWORKSPACE;
    PeerEnv e = make_env(part, op, (cudaStream_t)stream);
    const size_t kws = peer_carve(&e.w, (char *)d_workspace, part->n_own, op->nnz_own);
    const size_t vb = align256(sizeof(double) * (size_t)(part->n_own + 1));
    char *base = (char *)d_workspace + kws;
    double *kt = (double *)base, *rhs = (double *)(base + vb), *yh = (double *)(base + 2 * vb);
    const int64_t launches0 = g_kb_launches;
    const int64_t n = part->n_own;
    double *T = own_of(e, 2);
    k_pfill<<<kb_grid(n, 256, 8), 256, 0, e.s>>>(yh, n, 0.0);     // (not cudaMemset: see peer_preload)
    KB_LAUNCH_CHECK();
    k_reset_state<<<1, 32, 0, e.s>>>(e.w.tickets, e.w.done, e.w.iter, e.w.sc);   // the halo ticket of the first push below
    KB_LAUNCH_CHECK();
    k_peer_barrier<<<1, 32, 0, e.s>>>(e.c, ++part->ev);
    KB_LAUNCH_CHECK();
    // halo rows of T^0
    if ((rc = launch_pvec(e, PCopyPush{T, T}, make_push(e, 2, ++part->hseq[2]), nullptr))) return rc;
    int64_t total_iters = 0;
    double worst = 0.0;
    int status = 0, failed_step = 0, steps_done = 0, guess = 0;
    for (int step = 1; step <= nsteps; ++step) {
        if ((rc = peer_spmv<0>(e, d_Kff_vals, 2, part->hseq[2], kt, nullptr, nullptr, nullptr, NoFinal{}))) return rc;
        if ((rc = launch_pvec(e, PRhs{kt, d_Fm_own, a_symmetric ? d_scale_own : nullptr, rhs}, no_push(e), nullptr))) return rc;
        kb_solve_info info;
        if (a_symmetric) rc = peer_krylov<0>(e, op->d_vals, rhs, yh, d_scale_own, rtol, maxiter, 4, guess, &info);
        else rc = peer_krylov<1>(e, op->d_vals, rhs, yh, nullptr, rtol, maxiter, 4, guess, &info);
        if (rc) return rc;
        guess = info.iterations + 1;                           // the next step needs about as many
        total_iters += info.iterations;
        if (info.rhs_norm > 0 && info.residual / info.rhs_norm > worst) worst = info.residual / info.rhs_norm;
        if (info.status != 0 || !isfinite(info.residual)) {
            status = info.status != 0 ? info.status : KB_EBREAKDOWN;
            failed_step = step;
            if (!isfinite(info.residual)) worst = NAN;
            break;
        }
        steps_done = step;
        if ((rc = launch_pvec(e, PStep{d_renum_own, d_applied, yh, d_scale_own, T, dt}, make_push(e, 2, ++part->hseq[2]), nullptr))) return rc;
    }
    KB_CUDA(cudaStreamSynchronize(e.s));
    memset(h_info, 0, sizeof(*h_info));
    h_info->iterations = (int32_t)(total_iters > INT32_MAX ? INT32_MAX : total_iters);
    h_info->status = status;
    h_info->residual = worst;
    h_info->rhs_norm = 1.0;
    h_info->failed_step = failed_step;
    h_info->steps_done = steps_done;
    h_info->kernel_launches = g_kb_launches - launches0;
    return 0;
}
