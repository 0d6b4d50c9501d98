#!/usr/bin/env python
"""Development probe (torchrun): per-iteration time of the partitioned BiCGSTAB with NCCL vs peer-memory exchange."""
import os, sys, time
import torch
import torch.distributed as dist
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import kiltro_b200 as K
from kiltro_b200 import distributed as KD

rank = int(os.environ.get("RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1")); local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
nps = int(sys.argv[1]) if len(sys.argv) > 1 else 4000
iters = int(sys.argv[2]) if len(sys.argv) > 2 else 300
nxn, nyn = nps, nps * world
rect = ((0.0, 0.0), (1.0, float(world)), nxn - 1, nyn - 1)
h = 1.0 / (nxn - 1); U = 0.45 * 2.0 / h
part = KD.StripPartition(nxn, nyn, world, rank)
def velocity_fn(pts):
    v = torch.zeros_like(pts); v[:, 0] = U; return v
def dirichlet_fn(pts):
    f = torch.full((pts.shape[0], 1), float("nan"), dtype=torch.float64, device=pts.device)
    idx = torch.arange(part.nrows_loc, device=pts.device) * nxn
    f[idx, 0] = 100.0; f[idx + nxn - 1, 0] = 500.0
    return f
mat = K.FourierConduction(2, k=1.0, area=1.0, density=1.0, heat_capacity_v=1.0, q=0.0)
comm = KD.DistComm()
sysm = KD.build_local_system(part, rect, mat, K.HeatAdvectionDiffusion, velocity_fn, dirichlet_fn)
for ex in ("nccl", "peer", "nccl", "peer"):
    torch.cuda.synchronize(); dist.barrier(); t0 = time.perf_counter()
    sols, info = KD.solve_steady([sysm], comm, rtol=1e-30, maxiter=iters, check_every=25, method="bicgstab", exchange=ex)
    torch.cuda.synchronize(); dist.barrier(); dt = time.perf_counter() - t0
    if rank == 0:
        print("%s: %d iterations, %.1f us/iteration (incl. setup), exchange=%s" % (ex, info["iterations"], dt / max(info["iterations"], 1) * 1e6, info.get("exchange")), flush=True)
dist.destroy_process_group()
