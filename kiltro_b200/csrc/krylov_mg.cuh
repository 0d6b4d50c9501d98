// Opt-in solver: BiCGSTAB right-preconditioned by one multigrid V-cycle (multigrid.cu), beside the prescribed Jacobi
// method of krylov.cu.  Replaces the same call, LinearSolver.Solve -> spsolve (kiltro/FEMSolver.py:350-364).
//
//   r = b - A x ; rhat = r ; p = r ; rho = <rhat,r>
//   loop:  ph = M^-1 p ; v = A ph ; alpha = rho / <rhat,v>            (SpMV with fused dots: krylov.cu's kernels)
//          s = r - alpha v                                             (rides in the V-cycle's staging kernel)
//          sh = M^-1 s ; t = A sh ; omega = <s,t>/<t,t> ; rho' = <rhat,s> - omega <rhat,t> ; beta = (rho'/rho)(alpha/omega)
//          x += alpha ph + omega sh ; r = s - omega t ; p = r + beta (p - omega v) ; <r,r>     (one vector kernel)
//
// The matrix is used as handed over (no Jacobi scaling: M takes its place), all scalars live on device, the host looks at
// the `done` flag every `check_every` iterations, the verdict is taken on the TRUE residual b - A x (restart cycles as in
// krylov_solve).  The V-cycle runs in fp32; with right preconditioning x and r stay consistent in fp64 whatever M is.
// Included at the end of krylov.cu (same translation unit: it uses the launchers and functors defined there).
#pragma once
#include "multigrid.cuh"

namespace {

struct MgWork {
    Work w;                          // row partition, row-pattern dictionary, partials, scalars, flags (vals_hat/scale/cols16 unused)
    double *r, *s, *p, *v, *t, *rhat, *ph, *sh;
};

size_t carve_mg(MgWork *m, char *base, int64_t n, int64_t nnz) {
    const int64_t nblk = num_blocks(nnz);
    const int64_t nvec_blocks = 4 * (int64_t)kb_num_sms() * 8 + 4 * nblk + 64;
    size_t off = 0;
#define TAKE(lhs, type, count)                                    \
    do {                                                          \
        if (m) lhs = (type *)(base + off);                        \
        off += align256(sizeof(type) * (size_t)(count));          \
    } while (0)
    TAKE(m->r, double, n + 1); TAKE(m->s, double, n + 1); TAKE(m->p, double, n + 1); TAKE(m->v, double, n + 1); TAKE(m->t, double, n + 1);
    TAKE(m->rhat, double, n + 1); TAKE(m->ph, double, n + 1); TAKE(m->sh, double, n + 1);
    TAKE(m->w.bandflag, int, 4);
    TAKE(m->w.row_pat, uint8_t, n + 48);
    TAKE(m->w.pat_keys, unsigned long long, P_SLOTS);
    TAKE(m->w.pat_tab, pat_off_t, P_SLOTS * P_MAXLEN);
    TAKE(m->w.partials, double, nvec_blocks);
    TAKE(m->w.sc, double, S_COUNT);
    TAKE(m->w.rowblk, int32_t, nblk + 1);
    TAKE(m->w.ticket, unsigned int, 4);
    TAKE(m->w.done, int, 4);
    TAKE(m->w.iter, int, 4);
#undef TAKE
    if (m) { m->w.vals_hat = nullptr; m->w.scale = nullptr; m->w.cols16 = nullptr; m->w.use16 = false; }
    return off;
}

struct HMgUpdateP {         // x += alpha ph + omega sh ; r = s - omega t ; p = r + beta (p - omega v) ; d0 = <r,r>
    const double *alpha, *omega, *beta, *ph, *sh, *t, *v, *s; double *x, *r, *p;
    struct S { double a, om, b; }; struct L { double x, ph, sh, s, t, p, v; };
    __device__ S scalars() const { return S{alpha[0], omega[0], beta[0]}; }
    __device__ void load(int64_t i, L &l) const {
        l.x = x[i]; l.ph = __ldcs(ph + i); l.sh = __ldcs(sh + i); l.s = __ldcs(s + i); l.t = __ldcs(t + i); l.p = p[i]; l.v = __ldcs(v + i);
    }
    __device__ void apply(int64_t i, const L &l, const S &s, double &d0, double &) const {
        x[i] = l.x + (s.a * l.ph + s.om * l.sh);
        const double ri = l.s - s.om * l.t;
        r[i] = ri;
        p[i] = ri + s.b * (l.p - s.om * l.v);
        d0 += ri * ri;
    }
};

int mg_bicgstab(kb_mg *mg, int64_t n, int64_t nnz, const int32_t *indptr, const int32_t *indices, const double *vals,
                const double *b, double *x, double rtol, int maxiter, int check_every, void *workspace,
                size_t workspace_bytes, kb_solve_info *info, cudaStream_t s) {
    if (!mg || !indptr || !indices || !vals || !b || !x || !workspace || !info || n < 1 || nnz < 1) return KB_EINVAL;
    if (!kbmg_ready(mg) || kbmg_nfree(mg) != n) return KB_EINVAL;
    if (n >= INT32_MAX || nnz >= INT32_MAX) return KB_ETOOBIG;
    if (workspace_bytes < carve_mg(nullptr, nullptr, n, nnz)) return KB_EWORKSPACE;
    memset(info, 0, sizeof(*info));
    if (check_every < 1) check_every = 1;
    if (maxiter < 1) maxiter = 1;
    const int64_t launches0 = g_kb_launches;
    MgWork m;
    carve_mg(&m, (char *)workspace, n, nnz);
    Work &w = m.w;
    k_reset_state<<<1, 32, 0, s>>>(w.ticket, w.done, w.iter, w.sc);
    KB_LAUNCH_CHECK();
    // row partition + row-pattern dictionary of the matrix as handed over: kept in the workspace and reused while the same
    // pattern arrays and the same workspace come back (a FEM pattern is built once per mesh; the values may change)
    KbMgSolveCache *cache = kbmg_solve_cache(mg);
    if (cache->indptr == indptr && cache->indices == indices && cache->workspace == workspace && cache->n == n &&
        cache->nnz == nnz) {
        w.usepat = cache->usepat != 0 && !g_no_patterns && !g_force_idx32 && !g_force_simple_spmv && tma_eligible(indptr, w.row_pat, vals);
    } else {
        const int64_t nblk = num_blocks(nnz);
        k_partition<<<(unsigned)kb_cdiv(nblk + 1, 256), 256, 0, s>>>(n, nblk, indptr, w.rowblk);
        KB_LAUNCH_CHECK();
        KB_CUDA(cudaMemsetAsync(w.bandflag, 0, sizeof(int) * 4, s));
        w.usepat = false;
        int have = 0;
        if (!g_force_idx32 && !g_no_patterns) {
            const int g = kb_grid(n, 256, 8);
            KB_CUDA(cudaMemsetAsync(w.pat_keys, 0, sizeof(unsigned long long) * P_SLOTS, s));
            KB_CUDA(cudaMemsetAsync(w.row_pat, 0, (size_t)n + 48, s));
            k_pattern_insert<<<g, 256, 0, s>>>(n, indptr, indices, w.pat_keys, w.pat_tab, w.row_pat, w.bandflag);
            KB_LAUNCH_CHECK();
            k_pattern_verify<<<g, 256, 0, s>>>(n, indptr, indices, w.pat_tab, w.row_pat, w.bandflag);
            KB_LAUNCH_CHECK();
            int flag[2] = {1, 1};
            KB_CUDA(cudaMemcpyAsync(flag, w.bandflag, sizeof(int) * 2, cudaMemcpyDeviceToHost, s));
            KB_CUDA(cudaStreamSynchronize(s));
            have = flag[1] == 0;
            w.usepat = have && !g_force_simple_spmv && tma_eligible(indptr, w.row_pat, vals);
            *cache = KbMgSolveCache{indptr, indices, workspace, n, nnz, have};
        }
    }
    int rc;
    HostState hs;
    int total_iters = 0, status = KB_ENOTCONV, restarts = 0;
    const double rtol2 = rtol * rtol;
    double prev_true = INFINITY;
    bool recurrence_converged = false, attainable = false, verdict_known = false;
    for (int restart = 0; restart < 8 && total_iters < maxiter; ++restart) {
        restarts = restart;
        // (re)start from the TRUE residual: r = b - A x ; p = r ; rhat = r ; rho = <r,r>
        if ((rc = launch_spmv<0>(n, nnz, indptr, indices, vals, w, x, m.v, nullptr, nullptr, NoFinal{}, s))) return rc;
        if ((rc = launch_vec<1>(n, BodyResidual{b, m.v, m.r, m.p, m.rhat}, FinInit{w.sc, w.done, rtol2}, w, nullptr, s))) return rc;
        if ((rc = read_state(w, &hs, s))) return rc;
        const double true_rr = hs.sc[S_RES2], tol2 = rtol2 * hs.sc[S_AUX0];
        if (hs.done[0] == 1) { status = 0; verdict_known = true; break; }     // hs holds the true residual of the returned x
        if (hs.done[0] == 2) { status = KB_EBREAKDOWN; break; }
        if (restart > 0 && !(true_rr < 0.25 * prev_true) && true_rr <= 100.0 * tol2) { attainable = true; break; }
        if (recurrence_converged && !(true_rr < 0.25 * prev_true) && true_rr <= 1.0e6 * tol2) { attainable = true; break; }
        prev_true = true_rr;
        recurrence_converged = false;
        bool stop = false;
        while (!stop && total_iters < maxiter) {
            const int batch = (maxiter - total_iters) < check_every ? (maxiter - total_iters) : check_every;
            for (int it = 0; it < batch; ++it) {
                if ((rc = kbmg_apply(mg, m.p, nullptr, m.ph, nullptr, nullptr, w.done, s, true))) return rc;
                if ((rc = launch_spmv<1>(n, nnz, indptr, indices, vals, w, m.ph, m.v, m.rhat, w.done, FinBiAlpha{w.sc, w.done}, s, nullptr, true))) return rc;
                // s = r - alpha v is formed by the staging kernel of the V-cycle
                if ((rc = kbmg_apply(mg, m.r, m.s, m.sh, w.sc + S_ALPHA, m.v, w.done, s, true))) return rc;
                if ((rc = launch_spmv<2>(n, nnz, indptr, indices, vals, w, m.sh, m.t, m.s, w.done, FinBiOmega2{w.sc, w.done, w.iter}, s, m.rhat, true))) return rc;
                if ((rc = launch_vec4<1>(n, HMgUpdateP{w.sc + S_ALPHA, w.sc + S_OMEGA, w.sc + S_BETA, m.ph, m.sh, m.t, m.v, m.s, x, m.r, m.p},
                                         FinBiRes{w.sc, w.done, w.iter}, w.partials, w.ticket, w.done, s, true))) return rc;
            }
            if ((rc = read_state(w, &hs, s))) return rc;
            total_iters = hs.iter[0];
            if (hs.done[0] != 0) stop = true;
        }
        if (getenv("KB_DEBUG_SOLVER"))
            fprintf(stderr, "[kb mg solver] cycle %d: iters %d done %d why %.0f rho %.3e alpha %.3e omega %.3e beta %.3e rr %.3e tol2 %.3e\n",
                    restart, total_iters, hs.done[0], hs.sc[S_WHY], hs.sc[S_RHO], hs.sc[S_ALPHA], hs.sc[S_OMEGA], hs.sc[S_BETA],
                    hs.sc[S_RES2], hs.sc[S_TOL2]);
        if (hs.done[0] == 1) recurrence_converged = true;
        else if (hs.done[0] == 2) status = KB_EBREAKDOWN;
        else break;
        k_clear_done<<<1, 1, 0, s>>>(w.done);
        KB_LAUNCH_CHECK();
    }
    // verdict on the true residual of the returned x (already at hand when the loop ended on its own true-residual test)
    if (!verdict_known) {
        if ((rc = launch_spmv<0>(n, nnz, indptr, indices, vals, w, x, m.v, nullptr, nullptr, NoFinal{}, s))) return rc;
        if ((rc = launch_vec<1>(n, BodyResidual{b, m.v, m.r, m.p, nullptr}, FinInit{w.sc, w.iter + 2, rtol2}, w, nullptr, s))) return rc;
        if ((rc = read_state(w, &hs, s))) return rc;
    }
    const double res2 = hs.sc[S_RES2], b2 = hs.sc[S_AUX0];
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

extern "C" size_t kb_mg_bicgstab_workspace_bytes(int64_t nrows, int64_t nnz) { return carve_mg(nullptr, nullptr, nrows, nnz); }

extern "C" int kb_mg_bicgstab_solve(kb_mg *mg, int64_t nrows, int64_t nnz, const int32_t *d_indptr, const int32_t *d_indices,
                                    const double *d_vals, const double *d_b, double *d_x, double rtol, int maxiter,
                                    int check_every, void *d_workspace, size_t workspace_bytes, kb_solve_info *h_info,
                                    void *stream) {
    return mg_bicgstab(mg, nrows, nnz, d_indptr, d_indices, d_vals, d_b, d_x, rtol, maxiter, check_every, d_workspace,
                       workspace_bytes, h_info, (cudaStream_t)stream);
}
