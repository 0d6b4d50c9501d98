"""Development probe: the bench problem at nps x nps nodes solved by the prescribed Jacobi-BiCGSTAB and by BiCGSTAB
preconditioned with the experimental multigrid V-cycle (kiltro_b200/multigrid.py); same process, same box."""
import sys
import time
import warnings

import torch

sys.path.insert(0, ".")
import kiltro_b200 as K
from kiltro_b200.multigrid import GridMultigrid
import bench

warnings.simplefilter("ignore")
nps = int(sys.argv[1]) if len(sys.argv) > 1 else 1025
mesh, vel, flags, material = bench.make_problem(K, torch, nps)
form = K.HeatAdvectionDiffusion(mesh, velocity=vel)
fs = K.FEMSolver(verbose=False)


def run(solver):
    bc = K.BoundaryCondition(); bc.SetDirichletCriteria(lambda: flags)
    out = fs.Solve(formulation=form, mesh=mesh, material=material, boundary_condition=bc, solver=solver)
    return out.sol_device.clone()


ref = None
if "nojac" not in sys.argv:
    jac = K.LinearSolver(linear_solver_type="bicgstab", tol=1e-10, check_every=100)
    run(jac); ref = run(jac)
    print("jacobi-bicgstab : solve %.1f ms, %d iterations" % (fs.timings["solve_ms"], jac.info["iterations"]), flush=True)
bc0 = K.BoundaryCondition(); bc0.SetDirichletCriteria(lambda: flags)
torch.cuda.synchronize(); t0 = time.perf_counter()
mg = GridMultigrid(mesh, form, material, bc0)
torch.cuda.synchronize(); t_setup = time.perf_counter() - t0
mgs = K.LinearSolver(tol=1e-10, preconditioner=mg)
run(mgs); got = run(mgs)
err = float((got - ref).norm() / ref.norm()) if ref is not None else float("nan")
print("multigrid       : solve %.1f ms, %d iterations, %d levels, set-up %.2f s, launches %d, rel. diff to jacobi %.2e, "
      "true rel. residual %.2e, status %d, DOF/s (assembly + BC + solve) %.3e"
      % (fs.timings["solve_ms"], mgs.info["iterations"], mgs.info["levels"], t_setup, mgs.info["kernel_launches"], err,
         mgs.info["relative_residual"], mgs.info["status"],
         mesh.nnodes / (1e-3 * (fs.timings["solve_ms"] + fs.timings["assemble_ms"] + fs.timings["bc_ms"]))), flush=True)

if "sweep" in sys.argv:                                   # residual histories for a few smoother settings
    for nu, omega in ((2, 0.8), (3, 0.8), (2, 0.67)):
        mg.nu, mg.omega = nu, omega
        run(mgs)
        h = mgs.info["history"]
        print("nu=%d omega=%.2f: %d iterations, solve %.1f ms, history %s" % (
            nu, omega, mgs.info["iterations"], fs.timings["solve_ms"], " ".join("%.1e" % v for v in h[:6] + h[-4:])), flush=True)
