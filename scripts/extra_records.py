#!/usr/bin/env python
"""Single-GPU records beside the headline bench (SURVEY 8(d) config 4 with Pe = 0 -> CG, configs 3/5 at S16): prints one JSON
line per case.  Usage: python scripts/extra_records.py [nodes_per_side]"""
import json
import os
import sys
import warnings

import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import kiltro_b200 as K
import bench

warnings.simplefilter("ignore")
nps = int(sys.argv[1]) if len(sys.argv) > 1 else 4000
mesh, vel, flags, material = bench.make_problem(K, torch, nps)
nn = mesh.nnodes


def timed(fn):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); e0.record()
    out = fn()
    e1.record(); torch.cuda.synchronize()
    return out, e0.elapsed_time(e1)


# ---- steady pure diffusion (cell Pe = 0): symmetric system -> Jacobi-CG
form = K.HeatDiffusion(mesh)
bc = K.BoundaryCondition(); bc.SetDirichletCriteria(lambda: flags)
fs = K.FEMSolver(verbose=False)
solver = K.LinearSolver(linear_solver_type="auto", tol=1e-10, check_every=100)
step = lambda: fs.Solve(formulation=form, mesh=mesh, material=material, boundary_condition=bc, solver=solver)
step()
_, ms = timed(step)
info = solver.info
print(json.dumps({"case": "steady diffusion (Pe 0), Jacobi-CG rtol 1e-10, pattern cached", "ndof": nn, "ms_per_step": ms,
                  "dof_per_s": nn / (ms * 1e-3), "method": info["method"], "iterations": info["iterations"],
                  "ms_per_iteration": fs.timings["solve_ms"] / max(info["iterations"], 1),
                  "relative_residual": info["relative_residual"], "status": info["status"],
                  "phases_ms": fs.timings}), flush=True)
del fs, solver

# ---- transient (configs 3/5 physics at S16): initial field 200, T = 0 on x = 1, explicit (theta 0) and Crank-Nicolson
T0 = torch.full((nn, 1), 200.0, dtype=torch.float64, device="cuda")
tflags = torch.full((nn, 1), float("nan"), dtype=torch.float64, device="cuda")
tflags[torch.arange(nps, device="cuda") * nps + nps - 1, 0] = 0.0
nsteps = 20
for theta, f in ((0.0, form), (0.5, form), (0.5, K.HeatAdvectionDiffusion(mesh, velocity=vel))):
    bc = K.BoundaryCondition()
    bc.SetInitialConditions(lambda: T0); bc.SetDirichletCriteria(lambda: tflags)
    fs = K.FEMSolver(analysis_type="transient", number_of_increments=nsteps + 1, theta=theta, verbose=False)
    fs.return_device = True
    fs.snapshot_increments = [nsteps]
    run = lambda: fs.Solve(formulation=f, mesh=mesh, material=material, boundary_condition=bc, solver=K.LinearSolver(tol=1e-10))
    run()
    _, ms = timed(run)
    si = fs.solver_info
    print(json.dumps({"case": "transient theta=%.1f, %s, %d steps (assembly of K and M + loop on device)" % (
                          theta, "diffusion" if f is form else "convection-diffusion", nsteps),
                      "ndof": nn, "ms_total": ms, "ms_per_step": ms / nsteps, "dof_steps_per_s": nn * nsteps / (ms * 1e-3),
                      "inner_iterations": si["iterations"], "status": si["status"], "dt": fs.time_increment}), flush=True)
