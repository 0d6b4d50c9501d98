"""Development probe: run the numeric assembly a few times (for ncu captures).  usage: asm_only.py [nps] [method] [analysis]"""
import sys
import torch
sys.path.insert(0, ".")
import kiltro_b200 as K
import bench

nps = int(sys.argv[1]) if len(sys.argv) > 1 else 4000
method = sys.argv[2] if len(sys.argv) > 2 else "fused"
at = sys.argv[3] if len(sys.argv) > 3 else "steady"
mesh, vel, flags, material = bench.make_problem(K, torch, nps)
form = K.HeatAdvectionDiffusion(mesh, velocity=vel)
fs = K.FEMSolver(analysis_type=at, verbose=False, assembly=method)
bc = K.BoundaryCondition()
for _ in range(3):
    out = K.Assemble(fs, form.function_spaces, form, mesh, material, bc)
torch.cuda.synchronize()
print(fs.assembly_used)
