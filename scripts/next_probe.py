"""Development probe: time of the SURVEY 8(f) operators and of the .vtu writer at S16."""
import sys, time, os, tempfile
import torch
sys.path.insert(0, ".")
import kiltro_b200 as K
import bench
nps = int(sys.argv[1]) if len(sys.argv) > 1 else 4000
mesh, vel, flags, material = bench.make_problem(K, torch, nps)
form = K.HeatAdvectionDiffusion(mesh, velocity=vel)
fs = K.FEMSolver(verbose=False); fs.ComputeSparsityFEM(mesh, form)
bc = K.BoundaryCondition(); bc.applied_robin_convection = 5.0; bc.ground_robin_convection = 300.0
for name, fn in (("AssemblyRobinForces", lambda: K.AssemblyRobinForces(fs, form.function_spaces[0], mesh, material, bc)),
                 ("AssembleCharacteristicGalerkin", lambda: K.AssembleCharacteristicGalerkin(form.function_spaces, form, mesh, material, bc, fem_solver=fs))):
    for _ in range(2): fn()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); e0.record()
    for _ in range(5): fn()
    e1.record(); torch.cuda.synchronize()
    print("%s: %.3f ms per call at %d DOF" % (name, e0.elapsed_time(e1) / 5, mesh.nnodes), flush=True)
pp = K.PostProcess(2, 1); pp.SetMesh(mesh); pp.SetSolution(torch.rand((mesh.nnodes, 1, 1), dtype=torch.float64, device="cuda"))
d = tempfile.mkdtemp()
t0 = time.perf_counter(); files = pp.WriteVTK(os.path.join(d, "out.vtu")); dt = time.perf_counter() - t0
sz = sum(os.path.getsize(f) for f in files)
print("WriteVTK: %.2f s for %.2f GB (%.2f GB/s), %d points" % (dt, sz / 1e9, sz / 1e9 / dt, mesh.nnodes))
