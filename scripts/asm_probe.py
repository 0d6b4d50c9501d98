"""Development probe: numeric assembly time (CUDA events), fused vs two-pass, at a structured size."""
import sys
import torch
sys.path.insert(0, ".")
import kiltro_b200 as K
import bench

nps = int(sys.argv[1]) if len(sys.argv) > 1 else 4000
mesh, vel, flags, material = bench.make_problem(K, torch, nps)
form = K.HeatAdvectionDiffusion(mesh, velocity=vel)
methods = sys.argv[2].split(",") if len(sys.argv) > 2 else ("two_pass", "fused_generic", "fused", "fused_grid")
for method in methods:
    for at in ("steady", "transient"):
        fs = K.FEMSolver(analysis_type=at, verbose=False, assembly=method)
        bc = K.BoundaryCondition()
        for _ in range(2):
            out = K.Assemble(fs, form.function_spaces, form, mesh, material, bc)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10):
            out = K.Assemble(fs, form.function_spaces, form, mesh, material, bc)
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 10
        nn, ne, nnz = mesh.nnodes, mesh.nelem, out[0].nnz
        byt = 16 * ne + 32 * nn + 8 * nnz + 4 * (nn + 1) + 4 * nnz + (8 * nnz if at == "transient" else 0)
        print("nps=%d %-9s %-9s %.3f ms  %.1f GB/s algorithmic (%.1f%% of 6543)  plan=%s" % (
            nps, method, at, ms, byt / ms * 1e-6, byt / ms * 1e-6 / 65.434,
            getattr(fs.sparsity.tile_plan, "info", None) if method == "fused" else fs.assembly_used), flush=True)
