"""CPU tier of bench.py's host logic: the synthetic workload generator (bit-identical in NumPy and torch, independent of
the partition), the reference arm (the reference itself from baseline/_ref when installed) and the JSON contract of
`--impl reference`."""
import json
import os
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_hash_field_numpy_equals_torch_and_is_partition_independent():
    torch = pytest.importorskip("torch")
    import bench
    gid = np.arange(0, 200000, dtype=np.int64) * 97 + 5
    for salt in (1, 2):
        a = bench.hash_uniform(gid, salt)
        b = bench.hash_uniform(torch.from_numpy(gid), salt).numpy()
        assert np.array_equal(a, b)
        assert 0.0 <= a.min() and a.max() < 1.0 and abs(a.mean() - 0.5) < 5e-3 and abs(a.std() - 12 ** -0.5) < 5e-3
    nps = 64
    v = bench.synthetic_velocity(np.arange(nps * nps, dtype=np.int64), nps)
    # a strip of the same grid (rows 20..40) sees the same values: the field depends on the global id only
    strip = np.arange(20 * nps, 40 * nps, dtype=np.int64)
    assert np.array_equal(bench.synthetic_velocity(strip, nps), v[20 * nps:40 * nps])
    vt = bench.synthetic_velocity(torch.from_numpy(strip), nps).numpy()
    assert np.array_equal(vt, v[20 * nps:40 * nps])
    U = bench.PECLET * 2.0 * (nps - 1)
    assert abs(v[:, 0].mean() / U - 1.0) < 0.01 and abs(v[:, 1].std() / U - bench.NOISE) < 0.01


def test_host_problem_matches_the_oracle_mesh():
    import bench
    from oracle import kiltro_oracle as O
    pts, vel, d = bench.host_problem(23)
    p2, el = O.mesh_rectangle((0.0, 0.0), (1.0, 1.0), 22, 22)
    assert np.array_equal(pts, p2)
    assert np.isnan(d).sum() == 23 * 23 - 2 * 23 and d[0, 0] == 100.0 and d[22, 0] == 500.0


@pytest.mark.skipif(not os.path.isfile(os.path.join(ROOT, "baseline", "_ref", "kiltro", "FiniteElements",
                                                    "ComputeSparsityPattern.so")),
                    reason="baseline/_ref not installed (python baseline/make_ref.py needs /root/reference)")
def test_reference_runner_agrees_with_the_oracle():
    """The timed reference arm really is the reference: its solution equals the oracle's restatement (reference mode)."""
    import bench
    from oracle import kiltro_oracle as O
    ref = bench.ReferenceRunner()
    nps = 17
    nn, dt, sol = ref.step(nps)
    n2, dt2, sol2 = ref.step(nps, robin_stub=True)
    assert nn == nps * nps and np.array_equal(sol, sol2)            # the stub does not change the result
    pts, vel, d = bench.host_problem(nps)
    el = O.mesh_rectangle((0.0, 0.0), (1.0, 1.0), nps - 1, nps - 1)[1]
    want = O.steady_solve(pts, el, vel, (1.0, 1.0, 1.0, 1.0, 0.0), d, None, "reference", False)
    assert np.linalg.norm(sol - want) / np.linalg.norm(want) < 1e-11


def test_reference_arm_json_contract():
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "2", "--warmup", "1",
                          "--ref-nodes-per-side", "13", "--ref-stubbed-nodes-per-side", "17"],
                         capture_output=True, text=True, timeout=300, cwd=ROOT)
    assert out.returncode == 0, out.stderr
    lines = [l for l in out.stdout.splitlines() if l.strip()]
    assert len(lines) == 1
    rec = json.loads(lines[0])
    assert rec["impl"] == "reference" and rec["unit"] == "DOF/s" and rec["higher_is_better"] is True
    assert rec["e2e"]["h2d_bytes_per_step"] == 0 and rec["e2e"]["value"] == rec["value"]
    assert rec["cpu_baseline"]["kind"] in ("reference", "port") and rec["cpu_baseline"]["value"] == rec["value"]
    assert "workload" in rec["config"] and "model" not in rec["config"]
    if rec["cpu_baseline"]["kind"] == "reference":
        assert rec["steps"] == 2 and rec["warmup"] == 1 and rec["cpu_baseline"]["ndof"] == 169
