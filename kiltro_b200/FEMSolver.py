"""Driver + linear solver, mirroring kiltro/FEMSolver.py (reference file:line cited per method).

FEMSolver.Solve(formulation, mesh, material, boundary_condition) keeps the reference's signature and return type
(a PostProcess whose .sol is (nnodes, nvar, nincr) float64).  Every stage runs on the GPU through libkiltro_b200.so:
symbolic pattern (K2) -> numeric assembly (K1+K3) -> Dirichlet/Neumann (K4) -> Jacobi-preconditioned CG / BiCGSTAB
(K5+K6) -> solution scatter (K4); the transient branch, which the reference ships broken, runs the
theta-method loop (K7) on device with theta = 0 reproducing the reference's intended forward Euler."""
from copy import deepcopy
from time import time
from warnings import warn

import numpy as np
import torch

from . import _lib, device as D
from .MeshGeneration import Mesh
from .FiniteElements.Assembly import Assemble
from .PostProcess import PostProcess
from .sparse import DeviceCSR, issparse

_quiet = [False]


def _say(*a):
    if not _quiet[0]:
        print(*a)


class FEMSolver(object):
    # the reference stashes the two scatter maps on the solver (FEMSolver.py:321-322); here they are views of the cached
    # pattern, built on first access (they are not needed on Mesh.Rectangle-numbered meshes)
    data_local_indices = property(lambda self: None if self.sparsity is None else self.sparsity.data_local_indices)
    data_global_indices = property(lambda self: None if self.sparsity is None else self.sparsity.data_global_indices)

    def __init__(self, analysis_type="steady", analysis_nature="linear",
                 number_of_load_increments=1, print_incremental_log=False,
                 memory_store_frequency=1, recompute_sparsity_pattern=False,
                 squeeze_sparsity_pattern=False, number_of_increments=None, theta=0.0, time_step=None,
                 time_step_factor=0.1, include_heat_source=False, include_robin=False,
                 characteristic_galerkin=False, assembly="auto", verbose=True, partition=None, comm=None,
                 exchange="peer"):
        # FEMSolver.py:14-28 (+ the examples' number_of_increments=, examples/transient_2d.py:73-74)
        self.analysis_type = analysis_type
        self.analysis_nature = analysis_nature
        self.number_of_load_increments = number_of_load_increments
        self.save_frequency = int(memory_store_frequency)
        self.print_incremental_log = print_incremental_log
        self.recompute_sparsity_pattern = recompute_sparsity_pattern
        self.squeeze_sparsity_pattern = squeeze_sparsity_pattern
        self.is_sparsity_pattern_computed = False
        # extensions (all default to the reference's behaviour)
        self.number_of_increments = number_of_increments
        self.theta = float(theta)
        self.time_step = time_step
        self.time_step_factor = float(time_step_factor)          # FEMSolver.py:144 uses 0.1
        self.include_heat_source = include_heat_source           # the reference drops Flux from the RHS
        # code the reference ships but never reaches (Assembly.py:90 discards the Robin term,
        # FEMSolver.py:264-266,287-292 comments the characteristic-Galerkin correction out) -- opt-in here
        self.include_robin = include_robin
        self.characteristic_galerkin = characteristic_galerkin
        self.assembly = assembly                                  # "auto" | "fused" | "fused_generic" | "two_pass"
        self.verbose = verbose
        # SURVEY 8(e), new capability: partition="rows" -> the mesh handed to Solve() is ONE RANK'S strip of a row-block
        # partitioned mesh (Mesh.Rectangle(..., partition=StripPartition(...))); formulation / boundary_condition hold the
        # rank's local nodal data; Solve() returns the rank's OWNED part of the solution.  comm: a
        # kiltro_b200.distributed communicator (default: torch.distributed's world); exchange: "peer" (halo rows and dot
        # products through NVLink peer memory inside the kernels) or "nccl" (host-driven loops over the communicator).
        if partition not in (None, "rows"):
            raise ValueError("partition must be None or 'rows'")
        self.partition = partition
        self.comm = comm
        self.exchange = exchange
        self._psys = None
        self._psys_key = None
        self.sparsity = None
        self.timings = {}
        self.solver_info = None

    # ------------------------------------------------------------------------------------------ FEMSolver.py:31-62
    def __checkdata__(self, material, boundary_condition, formulation, mesh, function_spaces, solver):
        if mesh is None:
            raise ValueError("No mesh detected for the analysis")
        elif not isinstance(mesh, Mesh):
            raise ValueError("mesh has to be an instance of Kiltro.Mesh")
        if boundary_condition is None:
            raise ValueError("No boundary conditions detected for the analysis")
        if material is None:
            raise ValueError("No material model chosen for the analysis")
        if formulation is None:
            raise ValueError("No variational formulation specified")
        if function_spaces is None:
            if formulation.function_spaces is None:
                raise ValueError("No interpolation functions specified")
            else:
                function_spaces = formulation.function_spaces
        if solver is None:
            solver = LinearSolver(linear_solver="iterative", linear_solver_type="auto")
        if boundary_condition.initial_field is None and self.analysis_type == "transient":
            raise ValueError("The problem is TRANSIENT but there are not initial conditions.")
        return function_spaces, solver

    def __makeoutput__(self, mesh, TotalSol, formulation):       # FEMSolver.py:65-70
        post_process = PostProcess(formulation.ndim, formulation.nvar)
        post_process.SetMesh(mesh)
        post_process.SetSolution(TotalSol)
        post_process.solver_info = self.solver_info
        post_process.converged = getattr(self, "converged", True)
        return post_process

    # ------------------------------------------------------------------------------------------ FEMSolver.py:73-153
    def Solve(self, formulation=None, mesh=None, material=None,
              boundary_condition=None, function_spaces=None, solver=None):
        _quiet[0] = not self.verbose
        function_spaces, solver = self.__checkdata__(material, boundary_condition, formulation, mesh,
                                                     function_spaces, solver)
        self.PrintPreAnalysisInfo(mesh, formulation)
        if self.partition is not None:
            return self._solve_partitioned(function_spaces, formulation, mesh, material, boundary_condition, solver)
        self.ComputeSparsityFEM(mesh, formulation)

        _say('Assembling the system and acquiring neccessary information for the analysis...')
        boundary_condition.GetDirichletBoundaryConditions(formulation, mesh, material, self)
        NeumannFluxes = boundary_condition.ComputeNeumannFluxes(mesh, material)

        if self.analysis_nature == "nonlinear":
            raise ValueError("nonlinear analysis is not available (the reference's branch is unreachable too)")
        if self.analysis_type == "steady":
            if self.save_frequency != 1:
                raise ValueError("memory_store_frequency != 1 is not supported for steady analyses")
            TotalSol = self.SteadyLinearDiffusionSolver(function_spaces, formulation, mesh, material,
                                                        boundary_condition, solver, None, NeumannFluxes)
            return self.__makeoutput__(mesh, TotalSol, formulation)
        return self.TransientSolver(function_spaces, formulation, solver, mesh, material, boundary_condition,
                                    NeumannFluxes)

    # ------------------------------------------------------------------------------------------ FEMSolver.py:156-228
    def SteadyLinearDiffusionSolver(self, function_spaces, formulation, omesh, material,
                                    boundary_condition, solver, TotalSol, NeumannFluxes):
        mesh = omesh                                             # the reference deep-copies (:165) but never edits it
        LoadIncrement = int(self.number_of_load_increments)
        LoadFactor = 1. / LoadIncrement
        nnodes, nvar = mesh.nnodes, formulation.nvar
        if TotalSol is None:
            TotalSol = D.zeros((nnodes, nvar, LoadIncrement))
        self.converged = True
        NodalFluxes = -LoadFactor * NeumannFluxes                                            # :171
        AppliedDirichletInc = LoadFactor * boundary_condition.applied_dirichlet_device       # :172
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
        for Increment in range(LoadIncrement):
            ev[0].record()
            K, Flux, _ = Assemble(self, function_spaces, formulation, mesh, material, boundary_condition)   # :183
            Residual = NodalFluxes.clone()                                                   # :179,186
            if self.include_heat_source:
                Residual += LoadFactor * Flux          # opt-in: the reference computes Flux and drops it
            if self.include_robin:
                # opt-in Newton cooling h (T - T_inf) over the domain: (K + K_e) T = b + Flux_e.  The reference assembles
                # K_e, Flux_e here (Assembly.py:90) and drops them; its only consumer (BoundaryCondition.py:143-153) is
                # unreachable, so the sign convention is the physical one.
                from .FiniteElements.Assembly import AssemblyRobinForces
                K_e, Flux_e = AssemblyRobinForces(self, function_spaces[0], mesh, material, boundary_condition)
                _lib.check(_lib.load().kb_axpby(K.nnz, 1.0, D.ptr(K.data), 1.0, D.ptr(K_e.data), D.ptr(K.data),
                                                D.stream_ptr()))
                Residual += LoadFactor * Flux_e
            ev[1].record()
            K_b, F_b = boundary_condition.ApplyDirichletGetReducedMatrices(K, Residual, AppliedDirichletInc)[:2]   # :190
            ev[2].record()
            self.last_reduced_nnz = K_b.nnz
            sol = solver.Solve(K_b, F_b)                                                     # :193
            ev[3].record()
            dphi = boundary_condition.TotalComponentSol(sol, AppliedDirichletInc, 0, K.shape[0], nvar)   # :196
            TotalSol[:, :, Increment] += dphi                                                # :199
            torch.cuda.synchronize()
            self.timings = {"assemble_ms": ev[0].elapsed_time(ev[1]), "bc_ms": ev[1].elapsed_time(ev[2]),
                            "solve_ms": ev[2].elapsed_time(ev[3])}
            self.solver_info = getattr(solver, "info", None)
            # the reference's direct solve cannot fail silently; an iterative one can: the verdict travels with the result
            # (PostProcess.converged / .solver_info) and is announced even when the solver object swallowed it
            st = self.solver_info.get("status", 0) if isinstance(self.solver_info, dict) else 0
            if st != 0:
                self.converged = False
                warn("FEMSolver: the linear solve of load increment %d did not converge (status %d, relative residual %.3e)"
                     % (Increment, st, self.solver_info.get("relative_residual", float("nan"))))
            _say("Finished load increment " + str(Increment) + " for linear problem. Solver time is",
                 self.timings["solve_ms"] * 1e-3)
        if LoadIncrement > 1:                                                                # :225-226
            TotalSol = torch.cumsum(TotalSol, dim=2)
        return TotalSol

    # ------------------------------------------------------------------------------------------ new capability
    def _solve_partitioned(self, function_spaces, formulation, mesh, material, boundary_condition, solver):
        """Row-block partitioned Solve(): same stages as the single-GPU path, per rank -- numeric assembly of the strip
        (owned rows complete, no communication), Dirichlet / Neumann, right-hand side -- then the partitioned
        Jacobi-CG / BiCGSTAB (kiltro_b200.distributed.solve_steady) or theta-method loop (run_transient).  Returns a
        PostProcess whose .sol_device is (n_own, nvar, 1): the nodes this rank owns (.partition tells which)."""
        from . import distributed as KD
        part = getattr(mesh, "partition", None)
        if part is None:
            raise ValueError("FEMSolver(partition='rows') needs a mesh made with Mesh.Rectangle(..., partition=StripPartition(...))")
        if self.analysis_nature == "nonlinear":
            raise ValueError("nonlinear analysis is not available (the reference's branch is unreachable too)")
        if getattr(solver, "preconditioner", None) is not None:
            raise ValueError("LinearSolver(preconditioner=...) is a single-GPU option: the partitioned solvers are the "
                             "prescribed Jacobi-CG / Jacobi-BiCGSTAB")
        comm = self.comm
        if comm is None:
            comm = KD.DistComm() if part.world > 1 else KD.LocalComm(1)
            self.comm = comm
        if comm.world != part.world:
            raise ValueError("the communicator has %d ranks, the partition %d" % (comm.world, part.world))
        transient = self.analysis_type != "steady"
        key = (part.nxn, part.nyn, part.world, part.rank, mesh.nnodes, mesh.nelem, transient)
        if self._psys is None or self._psys_key != key:
            self._psys = KD.local_system_from(mesh, formulation, boundary_condition, self, material, transient)
            self._psys_key = key
        else:                                                    # pattern / index forms / peer mappings stay cached
            KD.attach_objects(self._psys, mesh, formulation, boundary_condition, self, material, transient)
        sysm = self._psys
        _say('Assembling the system and acquiring neccessary information for the analysis...')
        KD.reassemble(sysm)
        ev = sysm.events + [torch.cuda.Event(enable_timing=True)]
        method = getattr(solver, "solver_subtype", "auto")
        tol = getattr(solver, "tol", 1e-12)
        self.converged = True
        if not transient:
            n_glob = part.nxn * part.nyn
            maxiter = getattr(solver, "maxiter", None) or max(1000, 20 * int(np.sqrt(n_glob)) + 200)
            sols, info = KD.solve_steady([sysm], comm, rtol=tol, maxiter=int(maxiter),
                                         check_every=getattr(solver, "check_every", 50), method=method,
                                         exchange=self.exchange)
            ev[3].record()
            TotalSol = sols[0].reshape(part.n_own, formulation.nvar, 1)
        else:
            nincr = self.number_of_increments
            if nincr is None or int(nincr) < 2:
                raise ValueError("transient analysis needs FEMSolver(number_of_increments=N) with N >= 2")
            if self.time_step is None:
                raise ValueError("partitioned transient analysis needs FEMSolver(time_step=dt) (the reference's rule, "
                                 "FEMSolver.py:138-144, is a global minimum over the elements)")
            self.time_increment = float(self.time_step)
            sols, info = KD.run_transient([sysm], comm, self.time_increment, int(nincr) - 1, theta=self.theta, rtol=tol,
                                          maxiter=getattr(solver, "maxiter", None) or 2000, exchange=self.exchange)
            ev[3].record()
            TotalSol = sols[0].reshape(part.n_own, formulation.nvar, 1)       # the last increment only
        torch.cuda.synchronize()
        self.timings = {"assemble_ms": ev[0].elapsed_time(ev[1]), "bc_ms": ev[1].elapsed_time(ev[2]),
                        "solve_ms": ev[2].elapsed_time(ev[3])}
        info = dict(info)
        info.setdefault("tol", tol)
        self.solver_info = info
        if hasattr(solver, "info"):
            solver.info = info
        if info.get("status", 0) != 0:
            self.converged = False
            msg = "partitioned %s solve did not converge (status %d, relative residual %.3e)" % (
                info.get("method", "?"), info["status"], info.get("relative_residual", info.get("worst_relative_residual", float("nan"))))
            if getattr(solver, "raise_on_failure", False):
                raise _lib.KiltroB200Error(msg)
            warn(msg)
        post = self.__makeoutput__(mesh, TotalSol, formulation)
        post.partition = part
        return post

    # ------------------------------------------------------------------------------------------ FEMSolver.py:253-297
    def TransientSolver(self, function_spaces, formulation, solver, mesh, material, boundary_condition,
                        NeumannFluxes):
        from .transient import run_transient
        return run_transient(self, function_spaces, formulation, solver, mesh, material, boundary_condition,
                             NeumannFluxes)

    def PrintPreAnalysisInfo(self, mesh, formulation):           # FEMSolver.py:300-304
        _say('Pre-processing the information. Getting paths, solution parameters, mesh info, interpolation info etc...')
        _say('Number of nodes is', mesh.nnodes, 'number of DoFs is', mesh.nnodes * formulation.nvar)
        _say('Number of elements is', mesh.nelem)

    def ComputeSparsityFEM(self, mesh, formulation):             # FEMSolver.py:313-330
        if self.is_sparsity_pattern_computed is False:
            if self.recompute_sparsity_pattern is False:
                t_sp = time()
                from .FiniteElements.ComputeSparsityPattern import ComputeSparsityPatternDevice
                if self.squeeze_sparsity_pattern:
                    raise ValueError("Not implemented yet")
                sp = ComputeSparsityPatternDevice(mesh, formulation.nvar)
                self.sparsity = sp
                self.indices, self.indptr = sp.indices, sp.indptr
                self.is_sparsity_pattern_computed = True
                torch.cuda.synchronize()
                self.timings["pattern_s"] = time() - t_sp
                if self.verbose:
                    print("Computed sparsity pattern for the mesh. Time elapsed is {} seconds".format(time() - t_sp))
            else:
                raise ValueError("recompute_sparsity_pattern=True is not implemented (neither is it upstream: "
                                 "Assembly.py:54)")
        return self.sparsity


class LinearSolver(object):
    """Replaces kiltro/FEMSolver.py:335-364 (scipy spsolve) with device Krylov solvers (csrc/krylov.cu).

    linear_solver_type: "auto" (CG when the matrix is known symmetric, BiCGSTAB otherwise), "cg", "bicgstab".
    The reference's ("direct", "lapack") arguments are accepted and mapped to "auto" with the default tolerance.
    preconditioner (experimental, opt-in): a kiltro_b200.multigrid.GridMultigrid -> BiCGSTAB preconditioned by a V-cycle."""

    def __init__(self, linear_solver="iterative", linear_solver_type="auto", tol=1.0e-12, maxiter=None,
                 check_every=50, raise_on_failure=False, preconditioner=None):
        self.is_sparse = True
        self.solver_type = linear_solver
        self.solver_subtype = linear_solver_type if linear_solver_type in ("cg", "bicgstab") else "auto"
        self.has_umfpack = False
        self.tol = float(tol)
        self.maxiter = maxiter
        self.check_every = int(check_every)
        self.raise_on_failure = raise_on_failure
        self.info = None
        self.x0 = None
        # None: Jacobi (the prescribed method, csrc/krylov.cu).  EXPERIMENTAL opt-in: a kiltro_b200.multigrid.GridMultigrid
        self.preconditioner = preconditioner

    def Solve(self, A, b):
        if not issparse(A):
            raise ValueError("Linear system is not of sparse type")          # FEMSolver.py:355-356
        lib = _lib.load()
        b = D.to_device(b, torch.float64).reshape(-1)
        n = A.shape[0]
        if n == 0 and b.shape[0] == 0:
            warn("Empty linear system!!! Nothing to solve!!!")               # FEMSolver.py:358-360
            return b.clone()
        if b.shape[0] != n:
            raise ValueError("dimension mismatch between A and b")
        if self.preconditioner is not None:
            x, self.info = self.preconditioner.solve(A, b, rtol=self.tol, maxiter=int(self.maxiter or 200))
            if self.info["status"] != 0:
                msg = "%s stopped after %d iterations with relative residual %.3e (status %d)" % (
                    self.info["method"], self.info["iterations"], self.info["relative_residual"], self.info["status"])
                if self.raise_on_failure:
                    raise _lib.KiltroB200Error(msg)
                warn(msg)
            return x
        method = self.solver_subtype
        if method == "auto":
            method = "cg" if A.symmetric else "bicgstab"
        maxiter = int(self.maxiter) if self.maxiter is not None else max(1000, 20 * int(np.sqrt(n)) + 200)
        x = D.zeros(n) if self.x0 is None else D.to_device(self.x0, torch.float64).reshape(-1).clone()
        nbytes = lib.kb_krylov_workspace_bytes(n, A.nnz)
        ws = D.empty(nbytes, torch.uint8)
        info = _lib.kb_solve_info()
        fn = lib.kb_cg_solve if method == "cg" else lib.kb_bicgstab_solve
        _lib.check(fn(n, A.nnz, D.ptr(A.indptr), D.ptr(A.indices), D.ptr(A.data), D.ptr(b), D.ptr(x), self.tol,
                      maxiter, self.check_every, D.ptr(ws), nbytes, info, D.stream_ptr()))
        # residual / rhs_norm: TRUE ||b - A x||_2 and ||b||_2 of the system handed over (unscaled), recomputed from x
        self.info = {"method": method, "iterations": int(info.iterations), "status": int(info.status),
                     "residual": float(info.residual), "rhs_norm": float(info.rhs_norm),
                     "relative_residual": float(info.residual) / float(info.rhs_norm) if info.rhs_norm > 0 else 0.0,
                     "tol": self.tol, "attainable_accuracy": bool(info.attainable_accuracy),
                     "restarts": int(info.restarts), "kernel_launches": int(info.kernel_launches)}
        if info.status != 0:
            msg = "%s stopped after %d iterations with relative residual %.3e > tol %.1e (status %d: %s)" % (
                method, info.iterations, self.info["relative_residual"], self.tol, info.status,
                lib.kb_error_string(int(info.status)).decode())
            if self.raise_on_failure:
                raise _lib.KiltroB200Error(msg)
            warn(msg)
        elif info.attainable_accuracy:
            warn("%s reached the attainable accuracy of fp64: relative residual %.3e (requested %.1e, accepted within 10x)"
                 % (method, self.info["relative_residual"], self.tol))
        return x
