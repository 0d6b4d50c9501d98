"""Ad-hoc GPU probe (development aid): times the phases of the steady path at a given structured size."""
import sys
import time

import numpy as np
import torch

sys.path.insert(0, ".")
import kiltro_b200 as K
from kiltro_b200 import _lib, device as D


def run(n, mode="consistent", tol=1e-10, maxiter=20000, check_every=100):
    torch.cuda.synchronize()
    t0 = time.time()
    mesh = K.Mesh(element_type="quad")
    mesh.Rectangle((0, 0), (1, 1), n - 1, n - 1)
    nn = mesh.nnodes
    h = 1.0 / (n - 1)
    U = 0.45 * 2.0 / h
    g = torch.Generator(device="cuda"); g.manual_seed(0)
    vel = torch.zeros((nn, 2), dtype=torch.float64, device="cuda")
    vel[:, 0] = U
    vel += 0.1 * U * torch.randn((nn, 2), dtype=torch.float64, device="cuda", generator=g)
    flags = torch.full((nn, 1), float("nan"), dtype=torch.float64, device="cuda")
    idx = torch.arange(n, device="cuda") * n
    flags[idx, 0] = 100.0
    flags[idx + n - 1, 0] = 500.0
    bc = K.BoundaryCondition(); bc.SetDirichletCriteria(lambda: flags)
    mat = K.FourierConduction(2, k=1.0, area=1.0, density=1.0, heat_capacity_v=1.0, q=0.0)
    form = K.HeatAdvectionDiffusion(mesh, velocity=vel) if mode == "consistent" else K.HeatDiffusion(mesh)
    fs = K.FEMSolver(verbose=False)
    solver = K.LinearSolver(tol=tol, maxiter=maxiter, check_every=check_every)
    torch.cuda.synchronize(); t1 = time.time()
    sol = fs.Solve(formulation=form, mesh=mesh, material=mat, boundary_condition=bc, solver=solver)
    torch.cuda.synchronize(); t2 = time.time()
    print("n=%d ndof=%d mode=%s setup %.3fs solve-call %.3fs timings=%s info=%s" % (
        n, nn, mode, t1 - t0, t2 - t1, fs.timings, solver.info), flush=True)
    # second call: pattern cached
    torch.cuda.synchronize(); t1 = time.time()
    sol = fs.Solve(formulation=form, mesh=mesh, material=mat, boundary_condition=bc, solver=solver)
    torch.cuda.synchronize(); t2 = time.time()
    it = max(solver.info["iterations"], 1)
    print("   2nd call %.3fs timings=%s  ms/iter=%.4f  DOF/s=%.3e" % (
        t2 - t1, fs.timings, fs.timings["solve_ms"] / it,
        nn / ((fs.timings["assemble_ms"] + fs.timings["bc_ms"] + fs.timings["solve_ms"]) * 1e-3)), flush=True)
    s = sol.sol_device[:, 0, 0].reshape(n, n)
    print("   row n/2 samples:", s[n // 2, :: max(n // 8, 1)].cpu().numpy())
    # SpMV alone
    Kc = K.Assemble(fs, form.function_spaces, form, mesh, mat, bc)[0]
    x = torch.randn(nn, dtype=torch.float64, device="cuda")
    Kc.dot(x)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20):
        Kc.dot(x)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 20
    byt = 12 * Kc.nnz + 20 * nn
    print("   spmv %.4f ms -> %.1f GB/s algorithmic (incl. torch.empty per call)" % (ms, byt / ms * 1e-6), flush=True)


if __name__ == "__main__":
    for a in sys.argv[1:]:
        n, mode = a.split(":") if ":" in a else (a, "consistent")
        try:
            run(int(n), mode)
        except Exception as e:  # noqa
            import traceback; traceback.print_exc()
