"""Development probe: the S16 solve with each index form of the solver's SpMV, same process, same box."""
import sys
import torch
sys.path.insert(0, ".")
import kiltro_b200 as K
from kiltro_b200 import _lib
import bench

nps = int(sys.argv[1]) if len(sys.argv) > 1 else 4000
lib = _lib.load()
mesh, vel, flags, material = bench.make_problem(K, torch, nps)
form = K.HeatAdvectionDiffusion(mesh, velocity=vel)
fs = K.FEMSolver(verbose=False)
for rep in range(2):
    for mode in ("pattern", "offset16", "index32"):
        lib.kb_spmv_force_no_patterns(0 if mode == "pattern" else 1)
        lib.kb_spmv_force_idx32(1 if mode == "index32" else 0)
        bc = K.BoundaryCondition(); bc.SetDirichletCriteria(lambda: flags)
        solver = K.LinearSolver(linear_solver_type="bicgstab", tol=1e-10, check_every=25, maxiter=20000)
        lib.kb_profile_spmv_begin(8)
        fs.Solve(formulation=form, mesh=mesh, material=material, boundary_condition=bc, solver=solver)
        ns, tot = _lib.I64(0), _lib.F64(0.0)
        lib.kb_profile_spmv_end(ns, tot)
        it = solver.info["iterations"]
        print("%-9s solve %.1f ms, %d iterations, %.4f ms/iteration, sampled SpMV %.4f ms" % (
            mode, fs.timings["solve_ms"], it, fs.timings["solve_ms"] / it, tot.value / max(ns.value, 1)), flush=True)
