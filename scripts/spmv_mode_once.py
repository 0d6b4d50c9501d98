"""Development probe for ncu launch lists: a few BiCGSTAB iterations at S16 with one index form.  usage: mode [iters]"""
import sys
import torch
sys.path.insert(0, ".")
import kiltro_b200 as K
from kiltro_b200 import _lib
import bench
mode = sys.argv[1]; iters = int(sys.argv[2]) if len(sys.argv) > 2 else 20
lib = _lib.load()
mesh, vel, flags, material = bench.make_problem(K, torch, 4000)
form = K.HeatAdvectionDiffusion(mesh, velocity=vel)
fs = K.FEMSolver(verbose=False)
lib.kb_spmv_force_no_patterns(0 if mode == "pattern" else 1)
bc = K.BoundaryCondition(); bc.SetDirichletCriteria(lambda: flags)
solver = K.LinearSolver(linear_solver_type="bicgstab", tol=1e-10, check_every=iters, maxiter=iters)
import warnings
with warnings.catch_warnings():
    warnings.simplefilter("ignore")
    fs.Solve(formulation=form, mesh=mesh, material=material, boundary_condition=bc, solver=solver)
print(solver.info)
