"""Host-side logic of the row-block partition (K9) on CPU: partition arithmetic, and the exchange plumbing with a
world_size-2 gloo group on CPU tensors (the CUDA kernels themselves are covered by the single-device emulation in
tests/test_gpu_parity.py::test_partitioned_*)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from kiltro_b200.distributed import StripPartition, LocalComm, DistComm


def test_partition_covers_the_mesh():
    for nxn, nyn, world in [(10, 9, 2), (7, 32, 4), (4000, 32000, 8), (5, 17, 3), (3, 4, 1)]:
        parts = [StripPartition(nxn, nyn, world, r) for r in range(world)]
        assert parts[0].j0 == 0 and parts[-1].j1 == nyn
        for a, b in zip(parts[:-1], parts[1:]):
            assert a.j1 == b.j0 and a.ha == 1 and b.hb == 1
        assert parts[0].hb == 0 and parts[-1].ha == 0
        assert sum(p.n_own for p in parts) == nxn * nyn
        for p in parts:
            assert p.n_loc == p.n_own + (p.hb + p.ha) * nxn and p.own_off == p.hb * nxn
            assert p.global_off == p.j0 * nxn and p.j_first == p.j0 - p.hb
            assert max(q.j1 - q.j0 for q in parts) - min(q.j1 - q.j0 for q in parts) <= 1
    with pytest.raises(ValueError):
        StripPartition(10, 3, 2, 0)


def _global_to_local(part, g):
    lo = part.j_first * part.nxn
    return g[lo:lo + part.n_loc].clone()


def test_local_comm_halo_and_allreduce():
    nxn, nyn, world = 6, 11, 3
    parts = [StripPartition(nxn, nyn, world, r) for r in range(world)]
    g = torch.arange(nxn * nyn, dtype=torch.float64)
    vecs = []
    for p in parts:
        v = torch.full((p.n_loc,), -1.0, dtype=torch.float64)
        v[p.own()] = g[p.global_off:p.global_off + p.n_own]
        vecs.append(v)
    comm = LocalComm(world)
    comm.halo_exchange(parts, vecs)
    for p, v in zip(parts, vecs):
        assert torch.equal(v, _global_to_local(p, g))
    outs = [torch.tensor([1.0 * (r + 1), 2.0]) for r in range(world)]
    comm.allreduce(outs)
    for o in outs:
        assert o.tolist() == [6.0, 6.0]


def _worker(rank, world, port, nxn, nyn, ret):
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        part = StripPartition(nxn, nyn, world, rank)
        g = torch.arange(nxn * nyn, dtype=torch.float64) * 0.5
        v = torch.full((part.n_loc,), -1.0, dtype=torch.float64)
        v[part.own()] = g[part.global_off:part.global_off + part.n_own]
        comm = DistComm()
        comm.halo_exchange([part], [v])
        ok = torch.equal(v, _global_to_local(part, g))
        d = torch.tensor([float(rank + 1), 10.0], dtype=torch.float64)
        comm.allreduce([d])
        ok = ok and d.tolist() == [sum(range(1, world + 1)), 10.0 * world]
        ret[rank] = bool(ok)
    finally:
        dist.destroy_process_group()


def test_gloo_world2_halo_exchange_and_allreduce():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]
    world = 2
    ret = mp.Manager().dict()
    mp.spawn(_worker, args=(world, port, 8, 10, ret), nprocs=world, join=True)
    assert all(ret.get(r) for r in range(world)), dict(ret)


def test_best_iterate_checkpoint():
    """The partitioned loops' checkpoint of the best checked iterate (distributed._BestIterate): keeps the iterate with the
    smallest residual offered, hands it back on restore, and is inert when disabled (short inner solves)."""
    import torch
    from kiltro_b200.distributed import _BestIterate
    x = [torch.zeros(4, dtype=torch.float64), torch.ones(3, dtype=torch.float64)]
    keep = _BestIterate(x, 10.0)
    x[0] += 1.0; x[1] += 1.0
    keep.offer(x, 4.0)                                   # better: stored
    x[0] += 5.0; x[1] += 5.0
    keep.offer(x, 7.0)                                   # worse: ignored
    x[0].fill_(float("nan")); x[1].fill_(float("inf"))   # the iterate gets wrecked
    assert keep.restore(x) == 4.0
    assert torch.equal(x[0], torch.full((4,), 1.0, dtype=torch.float64)) and torch.equal(x[1], torch.full((3,), 2.0, dtype=torch.float64))
    off = _BestIterate(x, 1.0, enabled=False)
    off.offer(x, 0.5)
    assert off.x is None and off.restore(x) != off.restore(x)          # NaN: nothing to restore from
