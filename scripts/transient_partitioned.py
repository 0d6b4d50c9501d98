#!/usr/bin/env python
"""BASELINE config 5: synthetic 2D mesh, 4000 x (4000*N) nodes (128 M DOF at N = 8), transient theta-method, row-block
partitioned with NCCL halo exchange.  Launch with torchrun (one rank per GPU):
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 scripts/transient_partitioned.py [--theta 0.5] [--steps 20]
Prints one JSON line (rank 0): DOF*steps/s, ms per step, Krylov iterations per step."""
import argparse
import json
import os
import sys
import time

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--nodes-per-side", type=int, default=4000)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--theta", type=float, default=0.0)
    ap.add_argument("--dt-factor", type=float, default=0.1)
    ap.add_argument("--rtol", type=float, default=1e-10)
    args = ap.parse_args()
    import torch
    import torch.distributed as dist
    import kiltro_b200 as K
    from kiltro_b200 import distributed as KD
    rank = int(os.environ.get("RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1"); os.environ.setdefault("MASTER_PORT", "29577")
    dist.init_process_group("nccl", device_id=torch.device("cuda", local), rank=rank, world_size=world)
    nps = args.nodes_per_side
    nxn, nyn = nps, nps * world
    rect = ((0.0, 0.0), (1.0, float(world)), nxn - 1, nyn - 1)
    h = 1.0 / (nxn - 1)
    part = KD.StripPartition(nxn, nyn, world, rank)
    U = 0.45 * 2.0 / h

    def velocity_fn(pts):
        v = torch.zeros_like(pts); v[:, 0] = U
        return v

    def dirichlet_fn(pts):
        f = torch.full((pts.shape[0], 1), float("nan"), dtype=torch.float64, device=pts.device)
        idx = torch.arange(part.nrows_loc, device=pts.device) * nxn
        f[idx + nxn - 1, 0] = 0.0
        return f

    def initial_fn(pts):
        return torch.full((pts.shape[0], 1), 200.0, dtype=torch.float64, device=pts.device)

    material = K.FourierConduction(2, k=1.0, area=1.0, density=1.0, heat_capacity_v=1.0, q=0.0)
    comm = KD.DistComm()
    t0 = time.perf_counter()
    sysm = KD.build_local_system(part, rect, material, K.HeatAdvectionDiffusion, velocity_fn, dirichlet_fn, None, True, initial_fn)
    torch.cuda.synchronize(); t_setup = time.perf_counter() - t0
    dt = args.dt_factor * 1.0 * 1.0 * h * h / 2.0            # FEMSolver.py:138-144 on the uniform mesh
    KD.run_transient([sysm], comm, dt, 2, theta=args.theta, rtol=args.rtol)     # warm-up
    dist.barrier(); torch.cuda.synchronize(); t0 = time.perf_counter()
    T, info = KD.run_transient([sysm], comm, dt, args.steps, theta=args.theta, rtol=args.rtol)
    torch.cuda.synchronize(); dist.barrier(); sec = time.perf_counter() - t0
    tmax = torch.tensor([sec], dtype=torch.float64, device="cuda")
    dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
    if rank == 0:
        ndof = nxn * nyn
        print(json.dumps({"metric": "transient DOF*steps/s (config 5)", "value": ndof * args.steps / float(tmax.item()),
                          "unit": "DOF*steps/s", "n_gpus": world, "ndof": ndof, "steps": args.steps, "theta": args.theta,
                          "dt": dt, "ms_per_step": float(tmax.item()) / args.steps * 1e3,
                          "krylov_iterations_per_step": info["iterations"] / args.steps, "method": info["method"],
                          "worst_relative_residual": info["worst_relative_residual"], "t_setup_s": t_setup,
                          "T_min": float(T[0].min()), "T_max": float(T[0].max())}))
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
