"""Development probe: time of the Dirichlet reduction (numeric phase) at S16, CUDA events, 20 repetitions."""
import sys
import warnings

import torch

sys.path.insert(0, ".")
import kiltro_b200 as K
import bench

warnings.simplefilter("ignore")
mesh, vel, flags, material = bench.make_problem(K, torch, 4000)
form = K.HeatAdvectionDiffusion(mesh, velocity=vel)
fs = K.FEMSolver(verbose=False)
bc = K.BoundaryCondition(); bc.SetDirichletCriteria(lambda: flags)
Kc = K.Assemble(fs, form.function_spaces, form, mesh, material, bc)[0]
bc.GetDirichletBoundaryConditions(form, mesh, material, fs)
F = torch.zeros((mesh.nnodes, 1), dtype=torch.float64, device="cuda")
for _ in range(3):
    out = bc.ApplyDirichletGetReducedMatrices(Kc, F.clone(), bc.applied_dirichlet_device)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
Fs = [F.clone() for _ in range(20)]
torch.cuda.synchronize(); e0.record()
for k in range(20):
    out = bc.ApplyDirichletGetReducedMatrices(Kc, Fs[k], bc.applied_dirichlet_device)
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 20
nnz, n = Kc.nnz, mesh.nnodes
print("ApplyDirichletGetReducedMatrices: %.3f ms per call; algorithmic 12 nnz + 16 n = %.3f GB -> %.0f GB/s" % (
    ms, (12 * nnz + 16 * n) * 1e-9, (12 * nnz + 16 * n) / (ms * 1e-3) * 1e-9))
