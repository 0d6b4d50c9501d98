#!/usr/bin/env python
"""Multi-GPU check of the native peer-memory solvers (csrc/krylov_peer.cuh) against the host-driven NCCL loops and the
closed form.  Launch with torchrun, one rank per GPU:
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 scripts/peer_check.py
Rank 0 prints one JSON line with "pass": true/false."""
import json
import os
import sys

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))


def main():
    import torch
    import torch.distributed as dist
    import kiltro_b200 as K
    import bench
    from kiltro_b200 import distributed as KD
    rank = int(os.environ.get("RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    comm = KD.DistComm()
    rec = {"n_gpus": world}
    # (1) closed form through FEMSolver(partition="rows"), peer exchange
    rec.update(bench.closed_form_check(K, torch, dist, world, rank, "peer"))
    # (2) noisy steady case, peer vs NCCL loops (BiCGSTAB), then pure diffusion (CG)
    nps = 1024
    ok = rec["closed_form_ok"]
    for sym in (False, True):
        part, mesh, vel, flags, material = bench.strip_problem(K, torch, nps, world, rank)
        form = K.HeatDiffusion(mesh) if sym else K.HeatAdvectionDiffusion(mesh, velocity=vel)
        bc = K.BoundaryCondition(); bc.SetDirichletCriteria(lambda: flags)
        sols = {}
        for ex in ("peer", "nccl"):
            fs = K.FEMSolver(analysis_type="steady", partition="rows", exchange=ex, verbose=False)
            solver = K.LinearSolver(tol=1e-11, check_every=25, maxiter=100000)
            out = fs.Solve(formulation=form, mesh=mesh, material=material, boundary_condition=bc, solver=solver)
            sols[ex] = (out.sol_device[:, 0, 0].clone(), dict(fs.solver_info))
            if ex == "peer":
                rel = KD.residual_check(fs._psys, comm, sols[ex][0])
                KD.close_peer_regions([fs._psys])
        d = torch.stack([((sols["peer"][0] - sols["nccl"][0]) ** 2).sum(), (sols["nccl"][0] ** 2).sum()])
        dist.all_reduce(d)
        diff = float((d[0] / d[1]).sqrt())
        tag = "cg" if sym else "bicgstab"
        rec[tag] = {"peer_vs_nccl_rel_l2": diff, "iterations_peer": sols["peer"][1]["iterations"],
                    "iterations_nccl": sols["nccl"][1]["iterations"], "status_peer": sols["peer"][1]["status"],
                    "residual_check_nccl_path": rel}
        ok = ok and diff < 1e-8 and sols["peer"][1]["status"] == 0 and rel < 1e-10
    # (3) transient loop, peer vs NCCL
    for theta in (0.0, 0.5):
        T = {}
        for ex in ("peer", "nccl"):
            part = K.StripPartition(nps, 256 * world, world, rank)
            mesh = K.Mesh(element_type="quad")
            mesh.Rectangle((0.0, 0.0), (1.0, (256 * world - 1) / (nps - 1.0)), nps - 1, 256 * world - 1, partition=part)
            h = 1.0 / (nps - 1)
            vel = torch.zeros((part.n_loc, 2), dtype=torch.float64, device="cuda"); vel[:, 0] = 0.45 * 2.0 / h
            flags = torch.full((part.n_loc, 1), float("nan"), dtype=torch.float64, device="cuda")
            flags[torch.arange(part.nrows_loc, device="cuda") * nps + nps - 1, 0] = 0.0
            gid = mesh.global_node_ids().to(torch.float64)
            T0 = (200.0 + 20.0 * torch.sin(gid * 1e-3)).reshape(-1, 1)
            bc = K.BoundaryCondition(); bc.SetDirichletCriteria(lambda: flags); bc.SetInitialConditions(lambda: T0)
            fs = K.FEMSolver(analysis_type="transient", partition="rows", exchange=ex, number_of_increments=11, theta=theta,
                             time_step=0.1 * h * h / 2.0, verbose=False)
            out = fs.Solve(formulation=K.HeatAdvectionDiffusion(mesh, velocity=vel), mesh=mesh,
                           material=K.FourierConduction(2, k=1.0, area=1.0, density=1.0, heat_capacity_v=1.0, q=0.0),
                           boundary_condition=bc, solver=K.LinearSolver(tol=1e-12))
            T[ex] = (out.sol_device[:, 0, 0].clone(), dict(fs.solver_info))
            if ex == "peer":
                KD.close_peer_regions([fs._psys])
        d = torch.stack([((T["peer"][0] - T["nccl"][0]) ** 2).sum(), (T["nccl"][0] ** 2).sum()])
        dist.all_reduce(d)
        diff = float((d[0] / d[1]).sqrt())
        rec["transient_theta_%g" % theta] = {"peer_vs_nccl_rel_l2": diff, "iterations_peer": T["peer"][1]["iterations"],
                                             "iterations_nccl": T["nccl"][1]["iterations"], "status_peer": T["peer"][1]["status"]}
        ok = ok and diff < 1e-8 and T["peer"][1]["status"] == 0
    rec["pass"] = bool(ok)
    if rank == 0:
        print(json.dumps(rec))
    dist.destroy_process_group()
    sys.exit(0 if ok else 1)


if __name__ == "__main__":
    main()
