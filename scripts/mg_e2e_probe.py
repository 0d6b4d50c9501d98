"""Development probe: where the host-buffer (e2e) step of the multigrid leg spends its time."""
import sys
import time
import warnings

import torch

sys.path.insert(0, ".")
import kiltro_b200 as K
from kiltro_b200.multigrid import GridMultigrid
import bench

warnings.simplefilter("ignore")
nps = 4000
mesh, vel, flags, material = bench.make_problem(K, torch, nps)
fs = K.FEMSolver(analysis_type="steady", verbose=False)
form0 = K.HeatAdvectionDiffusion(mesh, velocity=vel)
fs.ComputeSparsityFEM(mesh, form0)
h_pts = mesh._points_d.cpu().pin_memory()
h_el = mesh._elements_d.to(torch.int64).cpu().pin_memory()
h_vel = vel.cpu().pin_memory()
h_flags = flags.cpu().pin_memory()
h_sol = torch.empty((mesh.nnodes, 1, 1), dtype=torch.float64).pin_memory()


def tick(label, t0, acc):
    torch.cuda.synchronize()
    t = time.perf_counter()
    acc.append((label, (t - t0) * 1e3))
    return t


for rep in range(3):
    acc = []
    torch.cuda.synchronize(); t = time.perf_counter()
    m = K.Mesh(element_type="quad"); m.set_arrays(h_pts, h_el); m.device_points(); m.device_elements()
    t = tick("mesh upload", t, acc)
    f = K.HeatAdvectionDiffusion(m, velocity=h_vel); f.device_velocity()
    t = tick("velocity upload", t, acc)
    b = K.BoundaryCondition(); b.SetDirichletCriteria(lambda: h_flags)
    t = tick("bc object", t, acc)
    mg = GridMultigrid(m, f, material, b)
    t = tick("GridMultigrid()", t, acc)
    sv = K.LinearSolver(tol=1e-10, preconditioner=mg, maxiter=200)
    o = fs.Solve(formulation=f, mesh=m, material=material, boundary_condition=b, solver=sv)
    t = tick("fs.Solve", t, acc)
    h_sol.copy_(o.sol_device, non_blocking=True); torch.cuda.current_stream().synchronize()
    t = tick("D2H", t, acc)
    del mg, sv, o, m, f, b
    t = tick("free", t, acc)
    print("rep %d: " % rep + " | ".join("%s %.1f" % kv for kv in acc) + " | timings %s" % {k: round(v, 2) for k, v in fs.timings.items()}, flush=True)
