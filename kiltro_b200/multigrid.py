"""EXPERIMENTAL, opt-in: geometric-multigrid preconditioner for quad meshes numbered like Mesh.Rectangle.

The reference solves with a direct method (FEMSolver.py:335-364); the prescribed replacement is Jacobi-preconditioned
CG / BiCGSTAB (LinearSolver's default, csrc/krylov.cu), whose iteration count grows like 1/h.  This module offers
BiCGSTAB right-preconditioned by ONE V-cycle instead:

    mg = GridMultigrid(mesh, formulation, material, boundary_condition)
    FEMSolver().Solve(..., solver=LinearSolver(preconditioner=mg, tol=1e-10))

  * hierarchy: every level is its own Mesh.Rectangle on the same box with about half the elements per direction
    (non-nested when the count is odd), nodal velocities / Dirichlet flags sampled at the nearest finer node;
  * operators: level 0 is the matrix handed to LinearSolver.Solve; the coarser ones are REDISCRETISED with the ordinary
    assembly kernels and -- because their cell Peclet number doubles per level -- with the Petrov-Galerkin weights of
    TemperaturePG / HeatAdvectionDiffusionPG (VariationalPrinciple.py:216-261,311-342).  Galerkin products R A P of the
    central operator diverge (tests/research/mg_prototype.py);
  * transfer: bilinear interpolation P between the two rectangles, restricted to the free DOFs, R = P^T -- stored as
    CSR matrices and applied with the SpMV kernel (round 2: stencil kernels; building P uses SciPy on the host, which
    is set-up of index structure, not solve-path arithmetic);
  * smoother: nu damped-Jacobi sweeps x += omega D^-1 (b - A x) before and after (kb_spmv, kb_axpby, kb_vec_mul);
  * coarsest level: Jacobi-BiCGSTAB on device (LinearSolver).

All solve-path arithmetic runs in libkiltro_b200's kernels; the Krylov recurrence is driven from Python with a handful of
scalar read-backs per iteration (an iteration is milliseconds of V-cycle work)."""
import math

import numpy as np
import torch

from . import _lib, device as D
from .sparse import DeviceCSR


def _interp_1d(nf, nc):
    """Linear interpolation from nc uniformly spaced nodes to nf uniformly spaced nodes on the same interval (CSR)."""
    import scipy.sparse as sp
    i = np.arange(nf, dtype=np.float64)
    s = i * (nc - 1) / (nf - 1)
    j = np.minimum(np.floor(s).astype(np.int64), nc - 2)
    w = s - j
    rows = np.concatenate([np.arange(nf), np.arange(nf)])
    cols = np.concatenate([j, j + 1])
    vals = np.concatenate([1.0 - w, w])
    keep = vals != 0.0
    return sp.csr_matrix((vals[keep], (rows[keep], cols[keep])), shape=(nf, nc))


def _device_csr(M):
    M = M.tocsr()
    M.sort_indices()
    return DeviceCSR(D.to_device(M.data.astype(np.float64), torch.float64), D.to_device(M.indices.astype(np.int32), torch.int32),
                     D.to_device(M.indptr.astype(np.int32), torch.int32), M.shape)


class _Level(object):
    pass


class GridMultigrid(object):

    def __init__(self, mesh, formulation, material, boundary_condition, nu=2, omega=0.67, coarsest=400, max_levels=12):
        from .FiniteElements.fused import grid_shape
        if formulation.ndim != 2 or formulation.nvar != 1:
            raise ValueError("GridMultigrid needs a 2-D scalar problem")
        if formulation.mode != _lib.KB_MODE_CONSISTENT and formulation.device_velocity() is not None:
            raise ValueError("GridMultigrid supports the consistent advection term (HeatAdvectionDiffusion) or pure diffusion")
        probe = _Level(); probe.grid, probe.nvar, probe.grid_scratch = None, 1, None      # (what grid_shape caches on)
        shape = grid_shape(probe, mesh)
        if shape is None:
            raise _lib.KiltroB200Error("GridMultigrid needs a quad-4 mesh numbered like Mesh.Rectangle")
        self.lib = _lib.load()
        self.nu, self.omega = int(nu), float(omega)
        self.material = material
        pts = mesh.device_points()
        self.box = (tuple(float(v) for v in pts[0].tolist()), tuple(float(v) for v in pts[-1].tolist()))
        flags = D.to_device(boundary_condition.dirichlet_flags, torch.float64).reshape(-1)
        vel = formulation.device_velocity()
        self.levels = []
        nxn, nyn = shape
        free = torch.isnan(flags).reshape(nyn, nxn)
        lvl = _Level()
        lvl.shape, lvl.free, lvl.A = (nxn, nyn), free.cpu().numpy(), None      # level 0: operator arrives with Solve()
        self.levels.append(lvl)
        while len(self.levels) < max_levels and int(self.levels[-1].free.sum()) > coarsest and min(nxn, nyn) > 5:
            nxc, nyc = nxn // 2 + 1, nyn // 2 + 1
            ix = torch.round(torch.arange(nxc, device=pts.device, dtype=torch.float64) * (nxn - 1) / (nxc - 1)).long()
            iy = torch.round(torch.arange(nyc, device=pts.device, dtype=torch.float64) * (nyn - 1) / (nyc - 1)).long()
            pick = (iy[:, None] * nxn + ix[None, :]).reshape(-1)             # nearest finer node of every coarse node
            flags = flags[pick].contiguous()
            vel = vel[pick].contiguous() if vel is not None else None
            lvl = self._coarse_level(nxc, nyc, vel, flags)
            fine = self.levels[-1]
            P = self._transfer(fine.shape, fine.free, (nxc, nyc), lvl.free)
            fine.P, fine.R = _device_csr(P), _device_csr(P.T)
            self.levels.append(lvl)
            nxn, nyn = nxc, nyc
        self.coarse_solver = None
        self.cycles = 0

    # ------------------------------------------------------------------------------------------------ set-up
    def _coarse_level(self, nxn, nyn, vel, flags):
        from .MeshGeneration import Mesh
        from .BoundaryCondition import BoundaryCondition
        from .FEMSolver import FEMSolver
        from .FiniteElements.Assembly import Assemble
        from .VariationalPrinciple import HeatDiffusion, HeatAdvectionDiffusionPG
        m = Mesh(element_type="quad")
        m.Rectangle(self.box[0], self.box[1], nxn - 1, nyn - 1)
        form = HeatAdvectionDiffusionPG(m, velocity=vel) if vel is not None else HeatDiffusion(m)
        fs = FEMSolver(verbose=False)
        bc = BoundaryCondition()
        f2 = flags.reshape(-1, 1)
        bc.SetDirichletCriteria(lambda: f2)
        K, _, _ = Assemble(fs, form.function_spaces, form, m, self.material, bc)
        bc.GetDirichletBoundaryConditions(form, m, self.material, fs)
        F = D.zeros((nxn * nyn, 1))
        K_b = bc.ApplyDirichletGetReducedMatrices(K, F, bc.applied_dirichlet_device)[0]
        lvl = _Level()
        lvl.shape, lvl.free = (nxn, nyn), torch.isnan(flags).reshape(nyn, nxn).cpu().numpy()
        self._attach(lvl, K_b)
        return lvl

    def _attach(self, lvl, A):
        n = A.shape[0]
        lvl.A, lvl.n = A, n
        lvl.dinv = D.empty(n)
        _lib.check(self.lib.kb_diag_scale(n, D.ptr(A.indptr), D.ptr(A.indices), D.ptr(A.data), 0, D.ptr(lvl.dinv),
                                          D.stream_ptr()))
        lvl.x, lvl.b, lvl.q, lvl.z = D.empty(n), D.empty(n), D.empty(n), D.empty(n)

    @staticmethod
    def _transfer(fshape, ffree, cshape, cfree):
        import scipy.sparse as sp
        P = sp.kron(_interp_1d(fshape[1], cshape[1]), _interp_1d(fshape[0], cshape[0]), format="csr")   # y-major, x fastest
        return P[ffree.ravel()][:, cfree.ravel()].tocsr()

    # ------------------------------------------------------------------------------------------------ kernels
    def _spmv(self, A, x, y):
        _lib.check(self.lib.kb_spmv(A.shape[0], A.nnz, D.ptr(A.indptr), D.ptr(A.indices), D.ptr(A.data),
                                    D.ptr(A.row_blocks()), D.ptr(x), D.ptr(y), D.stream_ptr()))

    def _axpby(self, a, x, b, y, out):
        _lib.check(self.lib.kb_axpby(out.shape[0], float(a), D.ptr(x), float(b), D.ptr(y), D.ptr(out), D.stream_ptr()))

    def _sweep(self, L, b):
        self._spmv(L.A, L.x, L.q)                                            # q = A x
        self._axpby(1.0, b, -1.0, L.q, L.q)                                  # q = b - A x
        _lib.check(self.lib.kb_vec_mul(L.n, D.ptr(L.dinv), D.ptr(L.q), D.ptr(L.z), D.stream_ptr()))
        self._axpby(1.0, L.x, self.omega, L.z, L.x)                          # x += omega D^-1 (b - A x)

    def _vcycle(self, l, b):
        L = self.levels[l]
        if l == len(self.levels) - 1:
            if self.coarse_solver is None:
                from .FEMSolver import LinearSolver
                self.coarse_solver = LinearSolver(linear_solver_type="bicgstab", tol=1.0e-10, check_every=16, maxiter=4000)
            return self.coarse_solver.Solve(L.A, b)
        _lib.check(self.lib.kb_vec_mul(L.n, D.ptr(L.dinv), D.ptr(b), D.ptr(L.z), D.stream_ptr()))
        self._axpby(self.omega, L.z, 0.0, None, L.x)                         # first sweep from x = 0
        for _ in range(self.nu - 1):
            self._sweep(L, b)
        self._spmv(L.A, L.x, L.q)
        self._axpby(1.0, b, -1.0, L.q, L.q)                                  # residual
        C = self.levels[l + 1]
        self._spmv(L.R, L.q, C.b)                                            # restrict
        xc = self._vcycle(l + 1, C.b)
        self._spmv(L.P, xc, L.z)                                             # prolongate
        self._axpby(1.0, L.x, 1.0, L.z, L.x)
        for _ in range(self.nu):
            self._sweep(L, b)
        return L.x

    def apply(self, A, r):
        """z ~ A^-1 r by one V-cycle (A: the level-0 operator of this Solve; returns a tensor owned by the hierarchy)."""
        L0 = self.levels[0]
        if L0.A is not A:
            if A.shape[0] != int(L0.free.sum()):
                raise ValueError("the matrix does not belong to the mesh / boundary conditions of this hierarchy")
            self._attach(L0, A)
        self.cycles += 1
        return self._vcycle(0, r)

    # ------------------------------------------------------------------------------------------------ Krylov driver
    def solve(self, A, b, rtol=1.0e-10, maxiter=200):
        """BiCGSTAB right-preconditioned by the V-cycle.  Returns (x, info) with info like LinearSolver.info."""
        lib = self.lib
        n = A.shape[0]
        scratch = D.zeros(int(lib.kb_blocks_scratch_bytes(A.nnz)) + 1024, torch.uint8)
        out2 = D.zeros(2)

        def dot2(a, c):                                                       # (<a,c>, <c,c>) -- one sync
            _lib.check(lib.kb_dot2(n, D.ptr(a), D.ptr(c), D.ptr(out2), D.ptr(scratch), D.stream_ptr()))
            v = out2.tolist()
            return v[0], v[1]

        launches0 = lib.kb_launch_count()
        x = D.zeros(n)
        r = b.clone(); rhat = b.clone(); p = b.clone()
        v, s, t, ph, sh = D.empty(n), D.empty(n), D.empty(n), D.empty(n), D.empty(n)
        rho, bb = dot2(rhat, r)
        status, it, rr = _lib.KB_ENOTCONV, 0, bb
        history = []
        tol2 = rtol * rtol * bb
        if bb == 0.0:
            status = 0
        while status != 0 and it < maxiter:
            ph.copy_(self.apply(A, p))
            self._spmv(A, ph, v)
            rv, _ = dot2(rhat, v)
            if rv == 0.0 or not math.isfinite(rv):
                status = _lib.KB_EBREAKDOWN
                break
            alpha = rho / rv
            self._axpby(1.0, r, -alpha, v, s)
            sh.copy_(self.apply(A, s))
            self._spmv(A, sh, t)
            st, tt = dot2(s, t)
            if tt == 0.0 or not math.isfinite(tt):
                self._axpby(1.0, x, alpha, ph, x)
                it += 1
                _, rr = dot2(s, s)
                status = 0 if rr <= tol2 else _lib.KB_EBREAKDOWN
                break
            omega = st / tt
            self._axpby(1.0, x, alpha, ph, x)
            self._axpby(1.0, x, omega, sh, x)
            self._axpby(1.0, s, -omega, t, r)
            it += 1
            rho_new, rr = dot2(rhat, r)
            history.append(math.sqrt(rr / bb) if rr >= 0.0 and bb > 0.0 else float("nan"))
            if not math.isfinite(rr):
                status = _lib.KB_EBREAKDOWN
                break
            if rr <= tol2:
                status = 0
                break
            if rho_new == 0.0 or omega == 0.0:
                status = _lib.KB_EBREAKDOWN
                break
            beta = (rho_new / rho) * (alpha / omega)
            self._axpby(1.0, p, -omega, v, p)
            self._axpby(1.0, r, beta, p, p)
            rho = rho_new
        # verdict by the TRUE residual, like the other drivers
        self._spmv(A, x, v)
        self._axpby(1.0, b, -1.0, v, v)
        _, rt = dot2(v, v)
        if status == 0 and bb > 0.0 and not (rt <= 1.0e6 * tol2):
            status = _lib.KB_ENOTCONV
        info = {"method": "bicgstab+multigrid", "iterations": it, "status": int(status), "vcycles": self.cycles,
                "levels": len(self.levels), "residual": math.sqrt(max(rt, 0.0)), "rhs_norm": math.sqrt(bb),
                "relative_residual": math.sqrt(max(rt, 0.0) / bb) if bb > 0 else 0.0,
                "kernel_launches": int(lib.kb_launch_count() - launches0), "history": history}
        return x, info
