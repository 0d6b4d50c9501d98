"""Opt-in geometric-multigrid preconditioner for quad meshes numbered like Mesh.Rectangle -- host mirror of
csrc/multigrid.cu + csrc/krylov_mg.cuh.

The reference solves with a direct method (FEMSolver.py:335-364); the prescribed replacement is Jacobi-preconditioned
CG / BiCGSTAB (LinearSolver's default, csrc/krylov.cu), whose iteration count grows like 1/h.  This module offers
BiCGSTAB right-preconditioned by ONE V-cycle instead, reported beside the prescribed method, never instead of it:

    mg = GridMultigrid(mesh, formulation, material, boundary_condition)
    FEMSolver().Solve(..., solver=LinearSolver(preconditioner=mg, tol=1e-10))

Everything runs natively (kb_mg_* entry points, include/kiltro_b200.h): the hierarchy of Mesh.Rectangle-like levels, the
rediscretised Petrov-Galerkin coarse operators (VariationalPrinciple.py:216-261,311-342) held as fp32 nine-point stencils,
damped-Jacobi sweeps, bilinear transfer kernels, a dense coarsest solve, and the BiCGSTAB recurrence with its scalars on
device.  This file only owns the workspaces and passes pointers."""
import ctypes

import numpy as np
import torch

from . import _lib, device as D


class GridMultigrid(object):
    # coarse-grid corrections per visit of level 0, 1, 2: the correction of level 2 (one sixteenth of the fine nodes) is
    # repeated three times, each from the corrected iterate.  With rediscretised upwind-type coarse operators a V-cycle's
    # asymptotic factor is ~0.5; this schedule costs 1.3x a V-cycle and halves the iteration count
    # (S16: 6 instead of 13; 1025^2 / 2001^2 nodes: 6 instead of 9 / 10).  () = V-cycle.
    DEFAULT_CYCLE = (1, 1, 3)

    def __init__(self, mesh, formulation, material, boundary_condition, nu=3, omega=0.67, coarsest=100, check_every=2,
                 cycle=None):
        from .FiniteElements.fused import grid_shape
        if formulation.ndim != 2 or formulation.nvar != 1:
            raise ValueError("GridMultigrid needs a 2-D scalar problem")
        if formulation.mode != _lib.KB_MODE_CONSISTENT and formulation.device_velocity() is not None:
            raise ValueError("GridMultigrid supports the consistent advection term (HeatAdvectionDiffusion) or pure diffusion")

        class _Probe(object):
            grid, nvar, grid_scratch = None, 1, None                          # (what grid_shape caches on)
        shape = grid_shape(_Probe(), mesh)
        if shape is None:
            raise _lib.KiltroB200Error("GridMultigrid needs a quad-4 mesh numbered like Mesh.Rectangle")
        self.lib = _lib.load()
        # omega: one damping factor, or a sequence of nu per-sweep factors (e.g. a Chebyshev sequence)
        self.omegas = [float(w) for w in omega] if hasattr(omega, "__len__") else None
        self.nu, self.omega, self.check_every = int(nu), (-1.0 if self.omegas else float(omega)), int(check_every)
        if self.omegas is not None and len(self.omegas) != self.nu:
            raise ValueError("omega must be one number or nu numbers")
        self.mesh, self.formulation, self.material, self.boundary_condition = mesh, formulation, material, boundary_condition
        handle = ctypes.c_void_p()
        _lib.check(self.lib.kb_mg_create(int(shape[0]), int(shape[1]), int(coarsest), ctypes.byref(handle)))
        self._h = handle
        self.levels = []
        for l in range(int(self.lib.kb_mg_num_levels(self._h))):
            nx, ny = _lib.I64(), _lib.I64()
            _lib.check(self.lib.kb_mg_level_shape(self._h, l, ctypes.byref(nx), ctypes.byref(ny)))
            self.levels.append((int(nx.value), int(ny.value)))
        if self.omegas is not None:
            _lib.check(self.lib.kb_mg_set_smoother(self._h, self.nu, (_lib.F64 * self.nu)(*self.omegas)))
        # cycle: coarse-grid corrections per visit of level 0, 1, 2, ... (missing levels: 1); None = the default schedule
        self.cycle = list(self.DEFAULT_CYCLE if cycle is None else cycle)
        _lib.check(self.lib.kb_mg_set_cycle(self._h, len(self.cycle), (ctypes.c_int32 * max(1, len(self.cycle)))(*self.cycle)))
        self._ws = D.empty(int(self.lib.kb_mg_workspace_bytes(self._h)), torch.uint8)
        self._solve_ws = None
        self.cycles = 0
        self.setups = 0

    def __del__(self):
        h, self._h = getattr(self, "_h", None), None
        if h is not None and getattr(self, "lib", None) is not None:
            self.lib.kb_mg_destroy(h)

    @staticmethod
    def transfer_table(nfine, ncoarse):
        """(j, w, start) of the 1-D linear interpolation between ncoarse and nfine uniformly spaced nodes (host-only
        arithmetic of kb_mg_transfer_table): x_f[i] = (1 - w[i]) x_c[j[i]] + w[i] x_c[j[i] + 1]."""
        j = np.empty(nfine, np.int32); w = np.empty(nfine, np.float32); start = np.empty(ncoarse + 1, np.int32)
        _lib.check(_lib.load().kb_mg_transfer_table(int(nfine), int(ncoarse), j.ctypes.data, w.ctypes.data, start.ctypes.data))
        return j, w, start

    # ------------------------------------------------------------------------------------------------ set-up
    def setup(self, A):
        """Numeric set-up for the reduced matrix A of this mesh / boundary conditions: level-0 stencils from A, coarse
        levels rediscretised from the current coordinates, velocities and Dirichlet flags."""
        bc = self.boundary_condition
        if bc.renum is None:
            bc.GetDirichletBoundaryConditions(self.formulation, self.mesh)
        if int(bc.n_free) != int(A.shape[0]):
            raise ValueError("the matrix does not belong to the mesh / boundary conditions of this hierarchy")
        pts, vel = self.mesh.device_points(), self.formulation.device_velocity()
        _lib.check(self.lib.kb_mg_setup(self._h, D.ptr(pts), D.ptr(vel), D.ptr(bc.renum), self.material.as_struct(),
                                        int(self.formulation.mode), int(A.shape[0]), int(A.nnz), D.ptr(A.indptr),
                                        D.ptr(A.indices), D.ptr(A.data), self.nu, self.omega, D.ptr(self._ws),
                                        self._ws.numel(), D.stream_ptr()))
        self.setups += 1

    def apply(self, r):
        """z ~ A^-1 r by one V-cycle (after setup(A)); fp64 in and out, the cycle itself runs in fp32."""
        z = D.empty(r.shape[0])
        _lib.check(self.lib.kb_mg_apply(self._h, D.ptr(r), D.ptr(z), D.stream_ptr()))
        self.cycles += 1
        return z

    # ------------------------------------------------------------------------------------------------ Krylov driver
    def solve(self, A, b, rtol=1.0e-10, maxiter=200, x0=None, reuse_setup=False):
        """BiCGSTAB right-preconditioned by the V-cycle (kb_mg_bicgstab_solve).  Returns (x, info) with info like
        LinearSolver.info.  The numeric set-up is redone on every call unless reuse_setup=True (same matrix values)."""
        lib = self.lib
        n = int(A.shape[0])
        if not (reuse_setup and self.setups > 0):
            self.setup(A)
        nbytes = int(lib.kb_mg_bicgstab_workspace_bytes(n, A.nnz))
        if self._solve_ws is None or self._solve_ws.numel() < nbytes:
            self._solve_ws = D.empty(nbytes, torch.uint8)
        x = D.zeros(n) if x0 is None else D.to_device(x0, torch.float64).reshape(-1).clone()
        info = _lib.kb_solve_info()
        # the library keeps the SpMV's row-pattern dictionary of (indptr, indices) in the solve workspace and reuses it while
        # the same arrays come back: holding the tensors keeps their addresses from being recycled for another pattern
        self._pattern_ref = (A.indptr, A.indices)
        _lib.check(lib.kb_mg_bicgstab_solve(self._h, n, A.nnz, D.ptr(A.indptr), D.ptr(A.indices), D.ptr(A.data), D.ptr(b),
                                            D.ptr(x), float(rtol), int(maxiter), self.check_every, D.ptr(self._solve_ws),
                                            self._solve_ws.numel(), info, D.stream_ptr()))
        self.cycles += 2 * int(info.iterations)
        out = {"method": "bicgstab+multigrid", "iterations": int(info.iterations), "status": int(info.status),
               "vcycles": 2 * int(info.iterations), "levels": len(self.levels), "residual": float(info.residual),
               "rhs_norm": float(info.rhs_norm),
               "relative_residual": float(info.residual) / float(info.rhs_norm) if info.rhs_norm > 0 else 0.0,
               "tol": float(rtol), "attainable_accuracy": bool(info.attainable_accuracy), "restarts": int(info.restarts),
               "kernel_launches": int(info.kernel_launches)}
        return x, out
