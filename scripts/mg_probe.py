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
import os as _os
from kiltro_b200 import _lib as _LL
if "MG_OVERWEIGHT" in _os.environ:
    _LL.load().kb_mg_set_overweight(float(_os.environ["MG_OVERWEIGHT"]))
if "MG_CYCLE" in _os.environ:
    from kiltro_b200.multigrid import GridMultigrid as _GM
    _GM.DEFAULT_CYCLE = tuple(int(v) for v in _os.environ["MG_CYCLE"].split(","))
if "MG_UPWIND" in _os.environ:
    _LL.load().kb_mg_set_upwind_restriction(int(_os.environ["MG_UPWIND"]))
import os
from kiltro_b200 import _lib as _L
if "MG_FUSED" in os.environ:
    _L.load().kb_mg_set_fused(int(os.environ["MG_FUSED"]))
if "MG_ROWS" in os.environ:
    _L.load().kb_mg_set_rows_per_task(int(os.environ["MG_ROWS"]))
if "MG_PFD" in os.environ:
    _L.load().kb_mg_set_prefetch_rows(int(os.environ["MG_PFD"]))
if "MG_MINROWS" in os.environ:
    _L.load().kb_mg_set_min_rows(int(os.environ["MG_MINROWS"]))
if "MG_AGGR" in os.environ:
    _L.load().kb_mg_set_aggressive(int(os.environ["MG_AGGR"]))
if "MG_TAIL" in os.environ:
    _L.load().kb_mg_set_tail(int(os.environ["MG_TAIL"]))
if "MG_PF1" in os.environ:
    _L.load().kb_mg_set_prefetch_rows_l1(int(os.environ["MG_PF1"]))
NU = int(os.environ.get("MG_NU", "3"))
OMEGA = [float(v) for v in os.environ["MG_OMEGAS"].split(",")] if "MG_OMEGAS" in os.environ else 0.67
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
mg = GridMultigrid(mesh, form, material, bc0, nu=NU, omega=OMEGA)
torch.cuda.synchronize(); t_setup = time.perf_counter() - t0
mgs = K.LinearSolver(tol=1e-10, preconditioner=mg)
run(mgs); got = run(mgs)
err = float((got - ref).norm() / ref.norm()) if ref is not None else float("nan")
print("multigrid       : solve %.1f ms, %d iterations, %d levels, set-up %.2f s, launches %d, rel. diff to jacobi %.2e, "
      "true rel. residual %.2e, status %d, DOF/s (assembly + BC + solve) %.3e"
      % (fs.timings["solve_ms"], mgs.info["iterations"], mgs.info["levels"], t_setup, mgs.info["kernel_launches"], err,
         mgs.info["relative_residual"], mgs.info["status"],
         mesh.nnodes / (1e-3 * (fs.timings["solve_ms"] + fs.timings["assemble_ms"] + fs.timings["bc_ms"]))), flush=True)

# numeric set-up alone (part of every solve): CUDA events around kb_mg_setup
from kiltro_b200 import device as D
bc1 = K.BoundaryCondition(); bc1.SetDirichletCriteria(lambda: flags)
Kc = K.Assemble(fs, form.function_spaces, form, mesh, material, bc1)[0]
bc1.GetDirichletBoundaryConditions(form, mesh, material, fs)
F = torch.zeros((mesh.nnodes, 1), dtype=torch.float64, device="cuda")
Kb, Fb = bc1.ApplyDirichletGetReducedMatrices(Kc, F, bc1.applied_dirichlet_device)[:2]
mg1 = GridMultigrid(mesh, form, material, bc1, nu=NU, omega=OMEGA)
mg1.setup(Kb)
e = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
e[0].record(); mg1.setup(Kb); e[1].record()
r = Fb.reshape(-1).clone()
z = mg1.apply(r); torch.cuda.synchronize()
e[2].record()
for _ in range(10):
    z = mg1.apply(r)
e[3].record(); torch.cuda.synchronize()
red = float(torch.linalg.vector_norm(r - Kb.dot(z)) / torch.linalg.vector_norm(r))
print("set-up %.3f ms ; V-cycle %.3f ms ; residual reduction of one cycle %.3f ; levels %s" % (
    e[0].elapsed_time(e[1]), e[2].elapsed_time(e[3]) / 10, red, mg1.levels), flush=True)

if "sweep" in sys.argv:                                   # a few smoother settings
    for nu, omega in ((1, 0.67), (2, 0.8), (3, 0.67), (2, 0.6), (2, 0.67)):
        bcs = K.BoundaryCondition(); bcs.SetDirichletCriteria(lambda: flags)
        mgs2 = K.LinearSolver(tol=1e-10, preconditioner=GridMultigrid(mesh, form, material, bcs, nu=nu, omega=omega))
        run(mgs2)
        print("nu=%d omega=%.2f: %d iterations, solve %.1f ms" % (nu, omega, mgs2.info["iterations"], fs.timings["solve_ms"]), flush=True)
