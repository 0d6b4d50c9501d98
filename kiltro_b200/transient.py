"""Transient driver (K7), the working counterpart of kiltro/FEMSolver.py:253-297 + :138-145 (unrunnable upstream:
SURVEY F2).  theta = 0 reproduces the reference's intent exactly (forward Euler, consistent mass, (M^-1)[in,in]);
theta in (0, 1] is the incremental theta-method on the same operators.  The whole time loop runs inside
libkiltro_b200.so (kb_transient_run); only the requested snapshots come back to the host."""
import ctypes

import numpy as np
import torch

from . import _lib, device as D
from .FiniteElements.Assembly import Assemble


def snapshot_indices(nincr):
    """Columns the reference returns: TotalSol[:,0,[0,1,int(N/3),int(2N/3),-1]] (FEMSolver.py:297)."""
    return [0, 1, int(nincr / 3), int(2 * nincr / 3), nincr - 1]


def run_transient(fem_solver, function_spaces, formulation, solver, mesh, material, boundary_condition, NeumannFluxes):
    lib = _lib.load()
    if getattr(solver, "preconditioner", None) is not None:
        from warnings import warn
        warn("LinearSolver(preconditioner=...) is ignored by the transient loop: its inner systems (M + theta dt K) are "
             "solved with the prescribed Jacobi-CG / Jacobi-BiCGSTAB")
    nincr = fem_solver.number_of_increments
    if nincr is None or int(nincr) < 2:
        raise ValueError("transient analysis needs FEMSolver(number_of_increments=N) with N >= 2")
    nincr = int(nincr)
    nvar = formulation.nvar
    ndof = mesh.nnodes * nvar
    bc = boundary_condition
    K, Flux, M = Assemble(fem_solver, function_spaces, formulation, mesh, material, bc)
    # time step: FEMSolver.py:138-144  (factor 0.1 upstream)
    if fem_solver.time_step is not None:
        dt = float(fem_solver.time_step)
    else:
        d_dt = D.empty(1)
        pts_d, el_d = mesh.device_points(), mesh.device_elements()      # locals: keep the buffers alive for the launch
        _lib.check(lib.kb_time_step_rule(formulation.ndim, D.ptr(pts_d), D.ptr(el_d), int(mesh.nelem),
                                         material.as_struct(), fem_solver.time_step_factor, D.ptr(d_dt), D.stream_ptr()))
        dt = float(d_dt.item())
    fem_solver.time_increment = dt
    theta = fem_solver.theta
    # F = K[in,out] T_D - F_neumann   (FEMSolver.py:269-272), zero at fixed DOFs
    F = D.zeros((ndof, 1))
    bc.ApplyDirichletGetReducedMatrices(K, F, bc.applied_dirichlet_device, only_residual=True)
    F = -F - NeumannFluxes
    if fem_solver.include_heat_source:
        F = F - Flux
    if getattr(fem_solver, "characteristic_galerkin", False):
        # T[in]^{n+1} = T[in]^n - dt (M^-1)[in,in] (K_b T + F_b) + dt^2/2 (M^-1)[in,in] (Ku_b T + Fu_b)   (FEMSolver.py:287-292,
        # commented out upstream)  ==  the explicit step on K - dt/2 K_u and F - dt/2 Flux_u; K_u is not lifted
        # (GetReducedMatrices, BoundaryCondition.py:156-167), so the Dirichlet term above came from K alone.
        if theta != 0.0:
            raise ValueError("characteristic_galerkin=True is the explicit scheme of the reference: use theta=0")
        from .FiniteElements.Assembly import AssembleCharacteristicGalerkin
        K_u, Flux_u = AssembleCharacteristicGalerkin(function_spaces, formulation, mesh, material, bc, fem_solver=fem_solver)
        Keff = D.empty(K.nnz)
        _lib.check(lib.kb_axpby(K.nnz, 1.0, D.ptr(K.data), -0.5 * dt, D.ptr(K_u.data), D.ptr(Keff), D.stream_ptr()))
        K = type(K)(Keff, K.indices, K.indptr, K.shape, symmetric=K.symmetric)
        F = F - 0.5 * dt * Flux_u
    Fm = D.empty(ndof)
    F = F.contiguous()
    _lib.check(lib.kb_mask_free(ndof, D.ptr(bc.renum), D.ptr(F), D.ptr(Fm), D.stream_ptr()))
    Kff = D.empty(K.nnz); A = D.empty(K.nnz)
    _lib.check(lib.kb_transient_matrix(ndof, K.nnz, D.ptr(K.indptr), D.ptr(K.indices), D.ptr(K.data), D.ptr(M.data),
                                       D.ptr(bc.renum), theta * dt, D.ptr(Kff), D.ptr(A), D.stream_ptr()))
    # T^0 = initial field (+ prescribed values on top of it at fixed DOFs, FEMSolver.py:274)
    T = D.to_device(bc.initial_field, torch.float64).reshape(-1).clone()
    if T.shape[0] != ndof:
        raise ValueError("initial_field has %d entries, expected %d" % (T.shape[0], ndof))
    T += bc.UpdateFixDoFs(bc.applied_dirichlet_device, ndof, nvar).reshape(-1)
    snaps = getattr(fem_solver, "snapshot_increments", None)
    if snaps is None:
        snaps = snapshot_indices(nincr)
    elif isinstance(snaps, str) and snaps == "all":
        snaps = list(range(nincr))
    order = sorted(range(len(snaps)), key=lambda i: snaps[i])
    steps_sorted = (ctypes.c_int32 * len(snaps))(*[int(snaps[i]) for i in order])
    out = D.empty((len(snaps), ndof))
    nbytes = lib.kb_transient_workspace_bytes(ndof, K.nnz)
    ws = D.empty(nbytes, torch.uint8)
    info = _lib.kb_solve_info()
    symmetric = 1 if (theta == 0.0 or formulation.is_symmetric) else 0
    tol = getattr(solver, "tol", 1e-12)
    maxiter = getattr(solver, "maxiter", None) or 2000
    _lib.check(lib.kb_transient_run(ndof, K.nnz, D.ptr(K.indptr), D.ptr(K.indices), D.ptr(Kff), D.ptr(A), symmetric,
                                    D.ptr(Fm), D.ptr(bc.renum), D.ptr(bc.applied_dirichlet_device), D.ptr(T), dt,
                                    nincr - 1, steps_sorted, len(snaps), D.ptr(out), tol, int(maxiter), D.ptr(ws), nbytes,
                                    info, D.stream_ptr()))
    fem_solver.solver_info = {"method": "cg" if symmetric else "bicgstab", "iterations": int(info.iterations),
                              "status": int(info.status), "worst_relative_residual": float(info.residual),
                              "kernel_launches": int(info.kernel_launches), "dt": dt, "theta": theta, "steps": nincr - 1,
                              "steps_done": int(info.steps_done), "failed_step": int(info.failed_step)}
    fem_solver.converged = info.status == 0
    if info.status != 0:
        # the loop stops at the first failed inner solve (T keeps the last good state, later snapshots are not written)
        msg = "transient loop stopped at step %d of %d: inner %s solve failed (status %d: %s)" % (
            info.failed_step, nincr - 1, "cg" if symmetric else "bicgstab", info.status,
            lib.kb_error_string(int(info.status)).decode())
        if getattr(solver, "raise_on_failure", False):
            raise _lib.KiltroB200Error(msg)
        from warnings import warn
        warn(msg)
        out[[k for k, i in enumerate(order) if int(snaps[i]) > int(info.steps_done)]] = float("nan")
    # undo the sort, return (nnodes*nvar, nsnap) like TotalSol[:,0,[...]]
    res = torch.empty_like(out)
    for k, i in enumerate(order):
        res[i] = out[k]
    fem_solver.transient_result_device = res.t().contiguous()
    if getattr(fem_solver, "return_device", False):
        return fem_solver.transient_result_device
    return D.to_numpy(fem_solver.transient_result_device)
