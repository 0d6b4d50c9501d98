"""Development probe: solution error vs oracle spsolve as a function of the Krylov tolerance (S16-style case)."""
import json
import sys
import time

import numpy as np
import torch

sys.path.insert(0, ".")
import kiltro_b200 as K
from oracle import kiltro_oracle as O

nps = int(sys.argv[1]) if len(sys.argv) > 1 else 1000
pts, el = O.mesh_rectangle((0.0, 0.0), (1.0, 1.0), nps - 1, nps - 1)
h = 1.0 / (nps - 1); U = 0.45 * 2.0 / h
rng = np.random.default_rng(0)
vel = np.zeros(pts.shape); vel[:, 0] = U; vel += 0.1 * U * rng.standard_normal(pts.shape)
d = np.full((pts.shape[0], 1), np.nan); d[::nps] = 100.0; d[nps - 1::nps] = 500.0
mat = (1.0, 1.0, 1.0, 1.0, 0.0)
t0 = time.time()
tm = {}
ref = O.steady_solve(pts, el, vel, mat, d, None, "consistent", False, timings=tm)
print("oracle %.1fs %s" % (time.time() - t0, tm), flush=True)
m = K.Mesh(element_type="quad"); m.set_arrays(pts, el)
bc = K.BoundaryCondition(); bc.SetDirichletCriteria(lambda: d)
material = K.FourierConduction(2, k=1.0, area=1.0, density=1.0, heat_capacity_v=1.0, q=0.0)
form = K.HeatAdvectionDiffusion(m, velocity=vel)
fs = K.FEMSolver(verbose=False)
res = []
for tol in (1e-6, 1e-8, 1e-10, 1e-12, 1e-13):
    solver = K.LinearSolver(linear_solver_type="bicgstab", tol=tol, maxiter=100000, check_every=50)
    sol = fs.Solve(formulation=form, mesh=m, material=material, boundary_condition=bc, solver=solver)
    err = np.linalg.norm(sol.sol.ravel() - ref) / np.linalg.norm(ref)
    res.append({"tol": tol, "rel_l2_vs_oracle": err, "info": solver.info, "solve_ms": fs.timings["solve_ms"]})
    print(res[-1], flush=True)
json.dump({"nps": nps, "oracle_timings": tm, "results": res}, open("gpurun_out/rtol_probe_%d.json" % nps, "w"), indent=1)
