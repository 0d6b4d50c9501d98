#!/usr/bin/env python
"""Multi-GPU check (torchrun, one rank per GPU): the peer-memory exchange path must reproduce the NCCL path and the
closed-form discrete solution.  Prints PASS/FAIL on rank 0 and exits non-zero on failure."""
import math
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import kiltro_b200 as K
from kiltro_b200 import distributed as KD


def main():
    rank = int(os.environ.get("RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    nps = int(sys.argv[1]) if len(sys.argv) > 1 else 600
    nxn, nyn = nps, 2 * world * 37 + 5                    # uneven strips on purpose
    rect = ((0.0, 0.0), (1.0, 1.0), nxn - 1, nyn - 1)
    h = 1.0 / (nxn - 1); Pe = 0.45; U = Pe * 2.0 / h
    part = KD.StripPartition(nxn, nyn, world, rank)

    def velocity_fn(pts):
        v = torch.zeros_like(pts); v[:, 0] = U
        return v

    def dirichlet_fn(pts):
        f = torch.full((pts.shape[0], 1), float("nan"), dtype=torch.float64, device=pts.device)
        idx = torch.arange(part.nrows_loc, device=pts.device) * nxn
        f[idx, 0] = 100.0; f[idx + nxn - 1, 0] = 500.0
        return f

    mat = K.FourierConduction(2, k=1.0, area=1.0, density=1.0, heat_capacity_v=1.0, q=0.0)
    comm = KD.DistComm()
    sysm = KD.build_local_system(part, rect, mat, K.HeatAdvectionDiffusion, velocity_fn, dirichlet_fn)
    res = {}
    for ex in ("nccl", "peer", "peer"):                   # the peer path twice: sequence numbers run on across solves
        sols, info = KD.solve_steady([sysm], comm, rtol=1e-12, maxiter=50000, check_every=25, method="bicgstab", exchange=ex)
        res.setdefault(ex, []).append((sols[0].clone(), info))
    n = nxn - 1
    lr = math.log((1 + Pe) / (1 - Pe))
    i = np.arange(nxn, dtype=np.float64)
    exact = 100.0 + 400.0 * np.exp((i - n) * lr) * (-np.expm1(-i * lr)) / (-np.expm1(-n * lr)); exact[0] = 100.0
    ex_t = torch.from_numpy(exact).cuda()
    ok = True
    msgs = []
    for name, lst in res.items():
        for k, (sol, info) in enumerate(lst):
            T = sol.reshape(-1, nxn)
            err = float((T - ex_t[None, :]).abs().max()) / 500.0
            ok = ok and info["status"] == 0 and err < 1e-8 and info.get("exchange") == name
            msgs.append("%s#%d: iterations %d, rel.res %.2e, max err vs closed form %.2e, exchange=%s" % (
                name, k, info["iterations"], info["relative_residual"], err, info.get("exchange")))
    d = float((res["nccl"][0][0] - res["peer"][1][0]).abs().max()) / 500.0
    ok = ok and d < 1e-8
    flag = torch.tensor([1 if ok else 0], device="cuda")
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    if rank == 0:
        print("\n".join(msgs)); print("max |nccl - peer| / 500 = %.2e" % d)
        print("PASS" if int(flag.item()) == 1 else "FAIL")
    dist.destroy_process_group()
    sys.exit(0 if int(flag.item()) == 1 else 1)


if __name__ == "__main__":
    main()
