// Geometric multigrid V-cycle on Mesh.Rectangle-numbered grids (multigrid.cu): the opt-in preconditioner of
// kb_mg_bicgstab_solve (krylov_mg.cuh).  Internal interface between the two translation units.
#pragma once
#include "common.cuh"

struct kb_mg;      // host-side plan + device pointers into the caller's workspace (include/kiltro_b200.h: opaque)

// z = V-cycle(v): v, z fp64 vectors of the REDUCED system (length nfree); stream-ordered, no host synchronisation.
// With d_alpha / d_q the vector that is preconditioned is  s = v - alpha[0] * q, and the kernel that stages it also writes
// it to d_s (!= d_v): BiCGSTAB's s = r - alpha v rides along.
// `done` (may be NULL): device flag, a non-zero value turns every kernel of the cycle into an early exit.
int kbmg_apply(kb_mg *mg, const double *d_v, double *d_s, double *d_z, const double *d_alpha, const double *d_q,
               const int *done, cudaStream_t s, bool pdl);
int64_t kbmg_nfree(const kb_mg *mg);
bool kbmg_ready(const kb_mg *mg);

// What kb_mg_bicgstab_solve keeps between calls in the caller's solve workspace (row partition + row-pattern dictionary of
// the matrix): valid while the same pattern arrays and the same workspace are passed again.
struct KbMgSolveCache { const void *indptr, *indices, *workspace; int64_t n, nnz; int usepat; };
KbMgSolveCache *kbmg_solve_cache(kb_mg *mg);
