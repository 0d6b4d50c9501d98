"""Development probe: the S16 solve with and without programmatic dependent launch in the Krylov loop (kb_set_pdl),
same process, same box; solutions must be bit-identical (PDL only moves launch/set-up time)."""
import sys
import torch
sys.path.insert(0, ".")
import kiltro_b200 as K
from kiltro_b200 import _lib
import bench

nps = int(sys.argv[1]) if len(sys.argv) > 1 else 4000
maxiter = int(sys.argv[2]) if len(sys.argv) > 2 else 1500
lib = _lib.load()
mesh, vel, flags, material = bench.make_problem(K, torch, nps)
form = K.HeatAdvectionDiffusion(mesh, velocity=vel)
fs = K.FEMSolver(verbose=False)
import warnings
warnings.simplefilter("ignore")
sols = {}
for rep in range(3):
    for pdl in (1, 0):
        lib.kb_set_pdl(pdl)
        bc = K.BoundaryCondition(); bc.SetDirichletCriteria(lambda: flags)
        solver = K.LinearSolver(linear_solver_type="bicgstab", tol=1e-10, check_every=100, maxiter=maxiter)
        out = fs.Solve(formulation=form, mesh=mesh, material=material, boundary_condition=bc, solver=solver)
        it = solver.info["iterations"]
        sols[pdl] = out.sol_device.clone()
        print("pdl=%d solve %.1f ms, %d iterations, %.4f ms/iteration" % (pdl, fs.timings["solve_ms"], it,
                                                                        fs.timings["solve_ms"] / it), flush=True)
print("bit-identical:", bool(torch.equal(sols[0], sols[1])))
lib.kb_set_pdl(1)
