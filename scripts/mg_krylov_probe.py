"""Development probe: how many V-cycles different outer iterations need on the bench problem (S16 by default) -- plain
defect correction (Richardson), truncated GCR(m) and the native BiCGSTAB -- all with the native V-cycle as the right
preconditioner.  Vector arithmetic of the prototypes runs in torch (this is a probe, not product code)."""
import sys
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
nps = int(sys.argv[1]) if len(sys.argv) > 1 else 4000
nu = int(sys.argv[2]) if len(sys.argv) > 2 else 3
import os
PE = float(os.environ.get("PROBE_PECLET", str(bench.PECLET)))
mesh, vel, flags, material = bench.make_problem(K, torch, nps)
vel = vel * (PE / bench.PECLET)
form = K.HeatAdvectionDiffusion(mesh, velocity=vel)
fs = K.FEMSolver(verbose=False)
bc = K.BoundaryCondition(); bc.SetDirichletCriteria(lambda: flags)
Kc = K.Assemble(fs, form.function_spaces, form, mesh, material, bc)[0]
bc.GetDirichletBoundaryConditions(form, mesh, material, fs)
F = torch.zeros((mesh.nnodes, 1), dtype=torch.float64, device="cuda")
A, b = bc.ApplyDirichletGetReducedMatrices(Kc, F, bc.applied_dirichlet_device)[:2]
b = b.reshape(-1).clone()
mg = GridMultigrid(mesh, form, material, bc, nu=nu)
mg.setup(A)
bn = float(b.norm())
tol = 1e-10


def richardson(maxit=60):
    x = torch.zeros_like(b); hist = []
    r = b.clone()
    for it in range(maxit):
        x += mg.apply(r)
        r = b - A.dot(x)
        hist.append(float(r.norm()) / bn)
        if hist[-1] <= tol:
            break
    return hist


def gcr(m, maxit=60):
    """right-preconditioned truncated GCR: one V-cycle + one SpMV per iteration, orthogonalisation against the last m
    directions"""
    x = torch.zeros_like(b); r = b.clone(); hist = []
    P, AP = [], []
    for it in range(maxit):
        z = mg.apply(r)
        az = A.dot(z)
        for p, ap in zip(P, AP):
            beta = torch.dot(ap, az)
            z = z - beta * p
            az = az - beta * ap
        nrm = az.norm()
        z = z / nrm; az = az / nrm
        alpha = torch.dot(az, r)
        x += alpha * z
        r -= alpha * az
        P.append(z); AP.append(az)
        if len(P) > m:
            P.pop(0); AP.pop(0)
        hist.append(float(r.norm()) / bn)
        if hist[-1] <= tol:
            break
    true = float((b - A.dot(x)).norm()) / bn
    return hist, true


h = richardson()
print("richardson: %d cycles, history %s" % (len(h), " ".join("%.1e" % v for v in h[:6] + h[-3:])), flush=True)
for m in ((1, 30) if "PROBE_SHORT" in os.environ else (1, 2, 4, 8, 30)):
    h, true = gcr(m)
    print("GCR(%d): %d cycles, true rel. residual %.2e, history %s" % (m, len(h), true, " ".join("%.1e" % v for v in h[:6] + h[-2:])), flush=True)
x, info = mg.solve(A, b, rtol=tol, reuse_setup=True)
print("native BiCGSTAB: %d iterations = %d cycles, rel. residual %.2e" % (info["iterations"], 2 * info["iterations"], info["relative_residual"]))
