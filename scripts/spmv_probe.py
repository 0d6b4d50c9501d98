"""Development probe: SpMV kernel time (CUDA events) for the one-shot and the persistent TMA kernel."""
import sys
import numpy as np
import torch
sys.path.insert(0, ".")
import kiltro_b200 as K
from kiltro_b200 import _lib, device as D
import bench

nps = int(sys.argv[1]) if len(sys.argv) > 1 else 4000
lib = _lib.load()
mesh, vel, flags, material = bench.make_problem(K, torch, nps)
form = K.HeatAdvectionDiffusion(mesh, velocity=vel)
fs = K.FEMSolver(verbose=False)
A = K.Assemble(fs, form.function_spaces, form, mesh, material, K.BoundaryCondition())[0]
n = A.shape[0]
x = torch.randn(n, dtype=torch.float64, device="cuda")
y = torch.empty_like(x)
rb = A.row_blocks()
byt = 12 * A.nnz + 20 * n
ref = None
for simple in (1, 0):
    lib.kb_spmv_force_simple(simple)
    def run():
        _lib.check(lib.kb_spmv(n, A.nnz, D.ptr(A.indptr), D.ptr(A.indices), D.ptr(A.data), D.ptr(rb), D.ptr(x), D.ptr(y), D.stream_ptr()))
    for _ in range(3):
        run()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20):
        run()
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 20
    if ref is None:
        ref = y.clone()
    print("nps=%d simple=%d  %.4f ms  %.1f GB/s algorithmic (%.1f%% of 6543)  max|dy|=%.3e" % (
        nps, simple, ms, byt / ms * 1e-6, byt / ms * 1e-6 / 65.434, (y - ref).abs().max().item()), flush=True)
