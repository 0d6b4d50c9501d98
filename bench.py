#!/usr/bin/env python
"""bench.py -- headline benchmark of the B200 hot path (BASELINE.json: assembly+solve DOF/s, 2D conv-diff at 16M DOF).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]

One "step" = one pass of the hot path over the synthetic workload S16 (BASELINE configs[3]):
numeric assembly (K1+K3) + Dirichlet/Neumann (K4) + Jacobi-BiCGSTAB solve (K5+K6) + solution scatter, i.e. exactly
FEMSolver.Solve() with the sparsity pattern cached on the solver (the reference caches it the same way and times it
separately: FEMSolver.py:313-325).  `value` is measured with every input already resident in HBM; `e2e` runs the same
step from pinned HOST buffers (reference array formats: float64 points/velocity/flags, int64 connectivity) with the
host->device copies and the device->host read of the solution inside the timed region.

--impl reference times the reference's OWN CPU implementation of the path -- the unmodified jdlaubrie/Kiltro package
installed into the git-ignored baseline/_ref/ by baseline/make_ref.py (its Cython extension rebuilt for this interpreter),
driven through its own public API (FEMSolver.Solve) -- on the box's host cores, each step a bounded sample of the same
workload.  Only when baseline/_ref/ is absent does it fall back to the oracle port (oracle/kiltro_oracle.py).

The synthetic velocity field is an integer hash of the GLOBAL node id (bit-identical in NumPy and torch, on any
partition), so N = 1, N > 1, the parity tests and the CPU arms all see the same problem.
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "assembly+solve DOF/s, 2D conv-diff at 16M DOF"
UNIT = "DOF/s"

# stdout carries exactly ONE JSON line: everything else any library prints (NCCL's version banner, solver warnings, ...)
# is sent to stderr by pointing fd 1 at fd 2 for the duration of the run; emit() writes to the saved descriptor.
_REAL_STDOUT = None


def claim_stdout():
    global _REAL_STDOUT
    if _REAL_STDOUT is None:
        sys.stdout.flush()
        _REAL_STDOUT = os.dup(1)
        os.dup2(2, 1)


def emit(line):
    sys.stdout.flush()
    os.write(_REAL_STDOUT if _REAL_STDOUT is not None else 1, (json.dumps(line) + "\n").encode())


# ------------------------------------------------------------------------------------------------ clocks
class ClockSampler(object):
    """nvidia-smi sampled every 200 ms during the timed region (B200_PROFILING.md recipe)."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.proc, self.lines = index, None, []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "200"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._pump, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0])); mx.append(float(f[1]))
            except ValueError:
                continue
            for nm, v in zip(names, f[3:7]):
                if v.lower() == "active":
                    reasons.add(nm)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        return {"sm_mhz": statistics.median(sm), "sm_max_mhz": max(mx), "reasons": sorted(reasons), "samples": len(sm)}


# ------------------------------------------------------------------------------------------------ workload
_H1, _H2 = -7046029254386353131, -4658895280553007687          # 0x9E3779B97F4A7C15, 0xBF58476D1CE4E5B9 as int64
PECLET = 0.45                                                   # cell Peclet number U h / 2k of the mean flow
NOISE = 0.1                                                     # rms of the velocity noise, in units of U


def hash_uniform(gid, salt):
    """Uniform [0, 1) pseudo-random number per int64 node id: wrap-around int64 arithmetic only, so NumPy (CPU arms, tests)
    and torch (device) give bit-identical values.  gid: int64 ndarray or tensor."""
    off = ((salt * _H2 + (1 << 63)) % (1 << 64)) - (1 << 63)                   # salt * _H2 wrapped to int64
    if hasattr(gid, "numpy") or hasattr(gid, "is_cuda"):                       # torch tensor
        import torch
        x = gid * _H1 + off
        x = (x ^ (x >> 31)) * _H2
        x = x ^ (x >> 29)
        return ((x >> 10) & ((1 << 53) - 1)).to(torch.float64) / float(1 << 53)
    import numpy as np
    with np.errstate(over="ignore"):
        x = gid.astype(np.int64) * np.int64(_H1) + np.int64(off)
        x = (x ^ (x >> 31)) * np.int64(_H2)
        x = x ^ (x >> 29)
    return ((x >> 10) & np.int64((1 << 53) - 1)).astype(np.float64) / float(1 << 53)


def synthetic_velocity(gid, nodes_per_side):
    """Nodal velocity of the benchmark case at the nodes with GLOBAL ids `gid` (x-fastest numbering of an
    nodes_per_side-wide grid with spacing h = 1/(nodes_per_side-1)): U (1, 0) + NOISE U (n1, n2), n uniform with unit
    variance, U from the cell Peclet number (k = 1)."""
    h = 1.0 / (nodes_per_side - 1)
    U = PECLET * 2.0 / h
    r3 = 1.7320508075688772
    n1 = (2.0 * hash_uniform(gid, 1) - 1.0) * r3
    n2 = (2.0 * hash_uniform(gid, 2) - 1.0) * r3
    if hasattr(gid, "is_cuda"):
        import torch
        return torch.stack([U + NOISE * U * n1, NOISE * U * n2], dim=1).contiguous()
    import numpy as np
    return np.stack([U + NOISE * U * n1, NOISE * U * n2], axis=1)


def make_problem(K, torch, nps, seed=0, peclet=PECLET):
    """S16-style synthetic case, generated on device: unit square, nps x nps nodes, k=rho=c_v=area=1, q=0, Dirichlet
    x=0 -> 100, x=1 -> 500, nodal velocity synthetic_velocity() (cell Peclet 0.45 + 10 % noise)."""
    mesh = K.Mesh(element_type="quad")
    mesh.Rectangle(lower_left_point=(0.0, 0.0), upper_right_point=(1.0, 1.0), nx=nps - 1, ny=nps - 1)
    nn = mesh.nnodes
    vel = synthetic_velocity(torch.arange(nn, dtype=torch.int64, device="cuda"), nps)
    flags = torch.full((nn, 1), float("nan"), dtype=torch.float64, device="cuda")
    left = torch.arange(nps, device="cuda") * nps
    flags[left, 0] = 100.0
    flags[left + nps - 1, 0] = 500.0
    material = K.FourierConduction(2, k=1.0, area=1.0, density=1.0, heat_capacity_v=1.0, q=0.0)
    return mesh, vel, flags, material


def host_problem(nps):
    """The same case as NumPy arrays (CPU arms, parity tests): points, elements (int64), velocity, Dirichlet flags."""
    import numpy as np
    x = np.linspace(0.0, 1.0, nps)
    X, Y = np.meshgrid(x, x)
    pts = np.stack([X.ravel(), Y.ravel()], axis=1)
    vel = synthetic_velocity(np.arange(nps * nps, dtype=np.int64), nps)
    d = np.full((nps * nps, 1), np.nan); d[::nps] = 100.0; d[nps - 1::nps] = 500.0
    return pts, vel, d


def algorithmic_bytes_spmv(nrows, nnz):
    return 12 * nnz + 20 * nrows


# ------------------------------------------------------------------------------------------------ CPU arms
REF_DIR = os.path.join(ROOT, "baseline", "_ref")


def reference_available():
    return os.path.isfile(os.path.join(REF_DIR, "kiltro", "FiniteElements", "ComputeSparsityPattern.so"))


class ReferenceRunner(object):
    """The reference's own implementation (baseline/_ref/kiltro, installed by baseline/make_ref.py) through its own public
    API: Mesh.Rectangle -> FourierConduction -> BoundaryCondition -> Temperature(mesh, velocity) -> FEMSolver().Solve(...),
    i.e. ComputeSparsityPattern (Cython/C++) + the per-element Python loop of Assemble + ApplyDirichletGetReducedMatrices
    + scipy spsolve (kiltro/FEMSolver.py:73-228,313-364).  Declared harness shims: stdout -> /dev/null (the
    reference prints once per element); robin_stub=True replaces AssemblyRobinForces -- whose result the reference
    discards (Assembly.py:90) but which builds DENSE ndof x ndof matrices, the reason the stock code stops near 1e5 DOF --
    by a no-op, outputs unchanged.  `Temperature` is the library's stock formulation (its advection term is the scalar
    broadcast of SURVEY F3; same loops, same pattern, same solver as the consistent term the GPU arm assembles)."""

    def __init__(self):
        if REF_DIR not in sys.path:
            sys.path.insert(0, REF_DIR)
        import kiltro as R
        import kiltro.FiniteElements.Assembly as RA
        self.R, self.RA = R, RA
        self._robin = RA.AssemblyRobinForces

    def step(self, nps, robin_stub=False):
        """One pass of the hot path at nps x nps nodes, sparsity pattern included (a fresh FEMSolver like a user's script;
        the pattern is ~0.3 % of the time).  Returns (ndof, seconds, solution)."""
        import contextlib
        R = self.R
        pts, vel, d = host_problem(nps)
        self.RA.AssemblyRobinForces = (lambda *a, **k: (None, None)) if robin_stub else self._robin
        try:
            with open(os.devnull, "w") as dn, contextlib.redirect_stdout(dn):
                mesh = R.Mesh(element_type="quad")
                mesh.Rectangle(lower_left_point=(0.0, 0.0), upper_right_point=(1.0, 1.0), nx=nps - 1, ny=nps - 1)
                mat = R.FourierConduction(2, k=1.0, area=1.0, density=1.0, heat_capacity_v=1.0, q=0.0)
                bc = R.BoundaryCondition(); bc.SetDirichletCriteria(lambda: d)
                form = R.Temperature(mesh, velocity=vel)
                fs = R.FEMSolver(analysis_type="steady")
                t0 = time.perf_counter()
                out = fs.Solve(formulation=form, mesh=mesh, material=mat, boundary_condition=bc)
                dt = time.perf_counter() - t0
        finally:
            self.RA.AssemblyRobinForces = self._robin
        return mesh.nnodes, dt, out.sol[:, 0, 0]


def cpu_oracle_run(nps, steps, warmup):
    """Fallback when baseline/_ref is absent: the oracle port of the reference's algorithm (batched NumPy element
    matrices + scatter-add + scipy spsolve, pattern cached), same synthetic workload at nps x nps nodes."""
    import numpy as np
    from oracle import kiltro_oracle as O
    pts, vel, d = host_problem(nps)
    el = O.mesh_rectangle((0.0, 0.0), (1.0, 1.0), nps - 1, nps - 1)[1]
    mat = (1.0, 1.0, 1.0, 1.0, 0.0)
    nn = pts.shape[0]
    ind, ptr = O.sparsity_pattern_fast(el, nn)
    dl, dg = O.data_indices_fast(el, ind, ptr)                  # pattern cached, like the GPU arm

    def step():
        A = O.assemble(pts, el, vel, mat, "consistent", False, False, (ind, ptr, dl, dg))
        cin, cout, td = O.dirichlet_split(d)
        kb = O.apply_dirichlet_reduce(ind, ptr, A["K"], np.zeros(nn), cin, cout, td)
        sol = O.linear_solve(kb[0], kb[1], kb[2], kb[3])
        return O.total_component_sol(sol, cin, cout, td, nn)

    for _ in range(warmup):
        step()
    t0 = time.perf_counter()
    for _ in range(steps):
        step()
    dt = time.perf_counter() - t0
    return nn, dt / steps


def cpu_baseline_block(nps_stock, steps, nps_stubbed):
    """`cpu_baseline` of the JSON line: the reference itself when baseline/_ref is installed (kind "reference"), else
    the oracle port (kind "port").  Bounded: `steps` stock steps at nps_stock^2 DOF + one Robin-stubbed step at
    nps_stubbed^2 DOF (0 = skip)."""
    if reference_available():
        ref = ReferenceRunner()
        tot, nn = 0.0, 0
        for _ in range(steps):
            nn, dt, _ = ref.step(nps_stock)
            tot += dt
        blk = {"value": nn * steps / tot, "unit": UNIT, "cores": 1, "kind": "reference",
               "sample": "UNMODIFIED reference (baseline/_ref/kiltro: FEMSolver.Solve with the stock Temperature formulation, "
                         "Cython pattern + per-element Python assembly + scipy spsolve; single-threaded code) on %dx%d nodes "
                         "(%d DOF) of the same synthetic case, %d timed step(s), %.2f s per step" % (
                             nps_stock, nps_stock, nn, steps, tot / steps),
               "ndof": nn}
        if nps_stubbed:
            n2, dt2, _ = ref.step(nps_stubbed, robin_stub=True)
            blk["robin_stubbed"] = {"value": n2 / dt2, "unit": UNIT, "ndof": n2, "seconds": dt2,
                                    "note": "same code with the dead dense AssemblyRobinForces pass (result discarded upstream, "
                                            "Assembly.py:90) replaced by a no-op: the reference's best case, and the only way it "
                                            "runs past ~1e5 DOF"}
        return blk
    cn = 384
    cnn, csec = cpu_oracle_run(cn, 1, 0)
    return {"value": cnn / csec, "unit": UNIT, "cores": 1, "kind": "port", "ndof": cnn,
            "sample": "baseline/_ref not installed: %dx%d nodes (%d DOF) of the same synthetic case, 1 timed step of "
                      "oracle/kiltro_oracle.py (batched NumPy assembly + scipy spsolve), pattern cached" % (cn, cn, cnn)}


def gpu_config(nps, world, rtol, partitioned):
    """`config` of the JSON line -- shared by the GPU arm and the reference arm (which runs bounded samples of it)."""
    if partitioned:
        nxn, nyn = nps, nps * world
        return {"workload": "S16 per GPU: synthetic 2D structured quad mesh %dx%d nodes (%d DOF total), steady "
                            "convection-diffusion (consistent Galerkin, cell Pe 0.45 + 10%% hash noise), Jacobi-BiCGSTAB rtol "
                            "%.0e, pattern cached" % (nxn, nyn, nxn * nyn, rtol),
                "ndof_per_gpu": nxn * nps, "ndof": nxn * nyn, "rtol": rtol,
                "l2_policy": "inputs (CSR matrix 1.7 GB + vectors per GPU) far larger than the 126 MB L2"}
    return {"workload": "S16: synthetic 2D structured quad mesh %dx%d nodes (%d DOF per GPU), steady convection-diffusion "
                        "(consistent Galerkin, cell Pe 0.45 + 10%% hash noise), Jacobi-BiCGSTAB rtol %.0e, pattern cached"
                        % (nps, nps, nps * nps, rtol),
            "ndof_per_gpu": nps * nps, "ndof": nps * nps * world, "rtol": rtol,
            "l2_policy": "inputs (CSR matrix 1.7 GB + vectors) far larger than the 126 MB L2"}


def run_reference_arm(args, rank, world):
    """bench.py --impl reference: K timed steps (after W untimed ones) of the reference's own CPU path on a bounded sample
    of the GPU arm's workload; rank 0 only."""
    if rank != 0:
        return
    cfg = gpu_config(args.nodes_per_side, world, args.rtol, world > 1)
    if reference_available():
        ref = ReferenceRunner()
        nps = args.ref_nodes_per_side
        for _ in range(args.warmup):
            ref.step(nps)
        tot, nn = 0.0, 0
        for _ in range(args.steps):
            nn, dt, _ = ref.step(nps)
            tot += dt
        sec = tot / max(1, args.steps)
        val = nn / sec
        cpu = {"value": val, "unit": UNIT, "cores": 1, "kind": "reference", "ndof": nn,
               "sample": "each step = FEMSolver.Solve of the UNMODIFIED reference (baseline/_ref/kiltro, stock Temperature "
                         "formulation; Cython pattern + per-element Python assembly + scipy spsolve; single-threaded code, "
                         "%d host cores present) on %dx%d nodes (%d DOF) of the workload named in config; %d timed + %d "
                         "warm-up steps" % (os.cpu_count() or 1, nps, nps, nn, args.steps, args.warmup)}
        if args.ref_stubbed_nodes_per_side:
            n2, dt2, _ = ref.step(args.ref_stubbed_nodes_per_side, robin_stub=True)
            cpu["robin_stubbed"] = {"value": n2 / dt2, "unit": UNIT, "ndof": n2, "seconds": dt2,
                                    "note": "one extra step with the dead dense AssemblyRobinForces pass (Assembly.py:90) "
                                            "replaced by a no-op: the reference's best case"}
        steps, warmup = args.steps, args.warmup
    else:
        nps = 384
        steps = max(1, min(args.steps, 3)); warmup = min(args.warmup, 1)
        nn, sec = cpu_oracle_run(nps, steps, warmup)
        val = nn / sec
        cpu = {"value": val, "unit": UNIT, "cores": 1, "kind": "port", "ndof": nn,
               "sample": "baseline/_ref not installed -> oracle port (batched NumPy element matrices + np.add.at + scipy "
                         "spsolve, pattern cached) on %dx%d nodes (%d DOF); %d timed step(s)" % (nps, nps, nn, steps)}
    line = {"metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus, "steps": steps, "warmup": warmup,
            "ms_per_step": sec * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
            "data": "synthetic", "impl": "reference", "config": cfg, "cpu_baseline": cpu,
            "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    emit(line)


# ------------------------------------------------------------------------------------------------ N > 1: partitioned
def strip_problem(K, torch, nps, world, rank, uniform=False):
    """Rank `rank`'s strip of the weak-scaling case: global mesh nps x (nps * world) nodes on [0,1] x [0,world] (S16 per
    GPU), through the ordinary API with a StripPartition.  uniform=True: noise-free velocity (closed-form check)."""
    nxn, nyn = nps, nps * world
    part = K.StripPartition(nxn, nyn, world, rank)
    mesh = K.Mesh(element_type="quad")
    mesh.Rectangle(lower_left_point=(0.0, 0.0), upper_right_point=(1.0, float(world)), nx=nxn - 1, ny=nyn - 1, partition=part)
    if uniform:
        vel = torch.zeros((part.n_loc, 2), dtype=torch.float64, device="cuda")
        vel[:, 0] = PECLET * 2.0 * (nxn - 1)
    else:
        vel = synthetic_velocity(mesh.global_node_ids(), nxn)            # hash of the GLOBAL node id: same field for every N
    flags = torch.full((part.n_loc, 1), float("nan"), dtype=torch.float64, device="cuda")
    idx = torch.arange(part.nrows_loc, device="cuda") * nxn
    flags[idx, 0] = 100.0
    flags[idx + nxn - 1, 0] = 500.0
    material = K.FourierConduction(2, k=1.0, area=1.0, density=1.0, heat_capacity_v=1.0, q=0.0)
    return part, mesh, vel, flags, material


def closed_form_check(K, torch, dist, world, rank, exchange, nps=1024, rows_per_rank=256):
    """Correctness evidence for the exchange kernels on THIS box and THIS rank count: a short uniform-velocity solve
    (nps x rows_per_rank*world nodes, cell Peclet 0.45) through the same partitioned API, against the exact discrete
    solution T_i = T_0 + dT (r^i - 1)/(r^n - 1), r = (1+Pe)/(1-Pe), which does not depend on y."""
    import math
    nxn, nyn = nps, rows_per_rank * world
    part = K.StripPartition(nxn, nyn, world, rank)
    mesh = K.Mesh(element_type="quad")
    mesh.Rectangle((0.0, 0.0), (1.0, float(nyn - 1) / (nxn - 1)), nxn - 1, nyn - 1, partition=part)     # square cells
    vel = torch.zeros((part.n_loc, 2), dtype=torch.float64, device="cuda"); vel[:, 0] = PECLET * 2.0 * (nxn - 1)
    flags = torch.full((part.n_loc, 1), float("nan"), dtype=torch.float64, device="cuda")
    idx = torch.arange(part.nrows_loc, device="cuda") * nxn
    flags[idx, 0] = 100.0; flags[idx + nxn - 1, 0] = 500.0
    material = K.FourierConduction(2, k=1.0, area=1.0, density=1.0, heat_capacity_v=1.0, q=0.0)
    bc = K.BoundaryCondition(); bc.SetDirichletCriteria(lambda: flags)
    fs = K.FEMSolver(analysis_type="steady", partition="rows", exchange=exchange, verbose=False)
    solver = K.LinearSolver(linear_solver_type="bicgstab", tol=1e-12, check_every=50, maxiter=50000)
    out = fs.Solve(formulation=K.HeatAdvectionDiffusion(mesh, velocity=vel), mesh=mesh, material=material,
                   boundary_condition=bc, solver=solver)
    n = nxn - 1
    lr = math.log((1 + PECLET) / (1 - PECLET))
    i = torch.arange(nxn, dtype=torch.float64, device="cuda")
    exact = 100.0 + 400.0 * torch.exp((i - n) * lr) * (-torch.expm1(-i * lr)) / (-math.expm1(-n * lr))
    T = out.sol_device[:, 0, 0].reshape(-1, nxn)
    err = (T - exact[None, :]).abs().max().reshape(1) / 500.0
    if world > 1:
        dist.all_reduce(err, op=dist.ReduceOp.MAX)
    if getattr(fs, "_psys", None) is not None:
        from kiltro_b200 import distributed as KD
        KD.close_peer_regions([fs._psys])
    return {"closed_form_max_error": float(err.item()), "closed_form_ok": bool(float(err.item()) < 1e-8),
            "closed_form_case": "%dx%d nodes, uniform velocity, rtol 1e-12, %d iterations, status %d" % (
                nxn, nyn, fs.solver_info["iterations"], fs.solver_info["status"])}


def run_partitioned(args, K, torch, dist, lib, _lib, rank, world, local, barrier):
    """Weak scaling of the same metric: the global mesh is nps x (nps*N) nodes (S16 per GPU), split into N row blocks, one
    per rank, through the drop-in API -- FEMSolver(partition="rows").Solve(...) on the rank's strip.  Every SpMV needs one
    halo mesh row per neighbour and every dot product a sum over the ranks; with exchange="peer" (default) the kernels
    move both through NVLink peer memory themselves."""
    from kiltro_b200 import distributed as KD
    nps = args.nodes_per_side
    nxn, nyn = nps, nps * world
    part, mesh, vel, flags, material = strip_problem(K, torch, nps, world, rank)
    form = K.HeatAdvectionDiffusion(mesh, velocity=vel)
    bc = K.BoundaryCondition(); bc.SetDirichletCriteria(lambda: flags)
    fs = K.FEMSolver(analysis_type="steady", partition="rows", exchange=args.exchange, verbose=False)
    solver = K.LinearSolver(linear_solver="iterative", linear_solver_type="bicgstab", tol=args.rtol,
                            check_every=args.check_every, maxiter=args.maxiter or 100000)

    def step():
        return fs.Solve(formulation=form, mesh=mesh, material=material, boundary_condition=bc, solver=solver)

    torch.cuda.synchronize(); t0 = time.perf_counter()
    out = step()                                              # first call: symbolic phase + peer mappings (reported as setup)
    torch.cuda.synchronize(); t_setup = time.perf_counter() - t0
    for _ in range(max(0, args.warmup - 1)):
        out = step()
    sampler = ClockSampler(local); sampler.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    launches0 = lib.kb_launch_count()
    e0.record()
    phase = {"assemble_ms": 0.0, "bc_ms": 0.0, "solve_ms": 0.0}
    for _ in range(args.steps):
        out = step()
        for k in phase:
            phase[k] += fs.timings[k]
    e1.record()
    barrier()
    launches = lib.kb_launch_count() - launches0
    clocks = sampler.stop()
    info = dict(fs.solver_info)
    tmax = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device="cuda")
    dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
    ms = float(tmax.item())
    ndof = nxn * nyn
    value = ndof * args.steps / (ms * 1e-3)
    solve_ms = torch.tensor([phase["solve_ms"] / args.steps], dtype=torch.float64, device="cuda")
    dist.all_reduce(solve_ms, op=dist.ReduceOp.MAX)

    # ---- verify (outside the timed region): (a) the true residual of the timed solution recomputed through a DIFFERENT
    # exchange path (NCCL halo exchange + plain SpMV on the unscaled operator + all-reduce); (b) a short closed-form solve
    verify = None
    if args.verify:
        rel = KD.residual_check(fs._psys, KD.DistComm(), out.sol_device[:, 0, 0].contiguous())
        verify = {"true_relative_residual": rel, "rtol": args.rtol, "residual_ok": bool(rel <= 10.0 * args.rtol),
                  "note": "||b - A x|| / ||b|| of the timed solution on A' = [Kff 0; 0 I] (unscaled), recomputed with an NCCL "
                          "halo exchange + plain SpMV + all-reduce, i.e. independently of the peer-memory kernels that "
                          "produced x; closed_form_*: a short uniform-velocity solve through the same API and exchange "
                          "against the exact discrete solution"}
        verify.update(closed_form_check(K, torch, dist, world, rank, args.exchange))
    if info["status"] != 0:
        raise SystemExit("bench.py: the timed partitioned solve did not converge (status %d)" % info["status"])

    # ---- e2e: every step uploads the rank's inputs from pinned host memory in the reference's array formats (float64
    # points / velocity / flags, int64 connectivity), goes through the same API call and reads the owned solution back
    e2e = None
    if args.e2e_steps > 0:
        pin = lambda t: t.cpu().pin_memory()
        h_pts, h_el = pin(mesh.device_points()), pin(mesh.device_elements().to(torch.int64))
        h_vel, h_flags = pin(vel), pin(flags)
        h_sol = torch.empty((part.n_own, 1, 1), dtype=torch.float64).pin_memory()
        h2d = (h_pts.numel() + h_el.numel() + h_vel.numel() + h_flags.numel()) * 8

        def e2e_step():
            m = K.Mesh(element_type="quad")
            m.set_arrays(h_pts, h_el, partition=part)                # host -> device inside
            f = K.HeatAdvectionDiffusion(m, velocity=h_vel)
            b = K.BoundaryCondition(); b.SetDirichletCriteria(lambda: h_flags)
            o = fs.Solve(formulation=f, mesh=m, material=material, boundary_condition=b, solver=solver)
            h_sol.copy_(o.sol_device, non_blocking=True)
            torch.cuda.current_stream().synchronize()
            return o

        e2e_step()                                               # warm-up
        barrier(); t0 = time.perf_counter()
        for _ in range(args.e2e_steps):
            e2e_step()
        barrier(); wall = time.perf_counter() - t0
        nb = torch.tensor([float(h2d)], dtype=torch.float64, device="cuda")
        dist.all_reduce(nb, op=dist.ReduceOp.SUM)
        t2 = torch.tensor([wall], dtype=torch.float64, device="cuda")
        dist.all_reduce(t2, op=dist.ReduceOp.MAX)
        e2e = {"value": ndof * args.e2e_steps / float(t2.item()), "unit": UNIT, "h2d_bytes_per_step": int(nb.item()),
               "d2h_bytes_per_step": int(ndof * 8), "steps": args.e2e_steps,
               "ms_per_step": float(t2.item()) / args.e2e_steps * 1e3, "iterations": fs.solver_info["iterations"],
               "note": "every rank uploads its strip (points, int64 connectivity, velocity, flags) from pinned host memory, "
                       "calls FEMSolver(partition='rows').Solve and reads its owned solution back, inside the timed "
                       "region; bytes are summed over the ranks"}
    # ---- BASELINE config 5 beside the headline, measured in the same run (bounded: `--config5-time-steps` time steps of
    # the theta = 0 loop on the same 4000 x 4000 N mesh; the stand-alone form is `--workload transient`)
    config5 = None
    if args.config5_time_steps > 0 and args.maxiter is None:
        KD.close_peer_regions([fs._psys])
        fs._psys = None
        tl = measure_transient(args, K, torch, dist, lib, rank, world, local, barrier, 1, 1, args.config5_time_steps, 0.0)
        config5 = {k: tl[k] for k in ("metric", "value", "unit", "ms_per_step", "config", "time_loop", "verify", "t_setup_ms")}
    if rank == 0:
        cfg = gpu_config(nps, world, args.rtol, True)
        peer = info.get("exchange") == "peer"
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
                "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
                "dtype": "f64", "data": "synthetic", "config": cfg,
                "parallelism": "row-block partition x%d through FEMSolver(partition='rows'); per SpMV one halo mesh row per "
                               "neighbour, per dot product a P-way sum: %s" % (
                                   world, "written / summed by the kernels through NVLink peer memory (no collective call and no "
                                   "host round trip inside an iteration)" if peer else "NCCL send/recv + all-reduce"),
                "solver": {"method": info["method"], "iterations": info["iterations"],
                           "relative_residual": info["relative_residual"], "status": info["status"],
                           "attainable_accuracy": info.get("attainable_accuracy"), "restarts": info.get("restarts"),
                           "exchange": info.get("exchange"),
                           "ms_per_iteration": float(solve_ms.item()) / max(1, info["iterations"])},
                "phases_ms_per_step": {k: v / args.steps for k, v in phase.items()},
                "verify": verify, "t_setup_ms": t_setup * 1e3, "roofline": None, "cpu_baseline": None, "e2e": e2e,
                "gpu_launches": int(launches), "clocks": clocks, "config5_transient": config5}
        emit(line)
    if fs._psys is not None:
        KD.close_peer_regions([fs._psys])


# ------------------------------------------------------------------------------------------------ config 5: transient
def run_transient_workload(args, K, torch, dist, lib, _lib, rank, world, local, barrier):
    line = measure_transient(args, K, torch, dist, lib, rank, world, local, barrier, args.steps, args.warmup, args.time_steps,
                             args.theta)
    if rank == 0:
        emit(line)


def measure_transient(args, K, torch, dist, lib, rank, world, local, barrier, steps, warmup, nt, theta):
    """BASELINE configs[4] / SURVEY 8(d) config 5: 4000 x (4000 N) nodes (128 M DOF at N = 8), theta-method time loop
    (theta = 0: the reference's intended forward Euler with a consistent-mass solve per step, FEMSolver.py:253-297),
    row-block partitioned, through FEMSolver(analysis_type='transient', partition='rows').  A bench "step" = `nt` time
    steps of the loop; value = DOF * time steps / s.  Returns the JSON line (every rank; rank 0 emits or embeds it)."""
    from kiltro_b200 import distributed as KD
    nps = args.nodes_per_side
    nxn, nyn = nps, nps * world
    part = K.StripPartition(nxn, nyn, world, rank)
    mesh = K.Mesh(element_type="quad")
    mesh.Rectangle((0.0, 0.0), (1.0, float(world)), nxn - 1, nyn - 1, partition=part)
    h = 1.0 / (nxn - 1)
    vel = torch.zeros((part.n_loc, 2), dtype=torch.float64, device="cuda"); vel[:, 0] = PECLET * 2.0 / h
    flags = torch.full((part.n_loc, 1), float("nan"), dtype=torch.float64, device="cuda")
    idx = torch.arange(part.nrows_loc, device="cuda") * nxn
    flags[idx + nxn - 1, 0] = 0.0
    T0 = torch.full((part.n_loc, 1), 200.0, dtype=torch.float64, device="cuda")
    material = K.FourierConduction(2, k=1.0, area=1.0, density=1.0, heat_capacity_v=1.0, q=0.0)
    form = K.HeatAdvectionDiffusion(mesh, velocity=vel)
    bc = K.BoundaryCondition(); bc.SetDirichletCriteria(lambda: flags); bc.SetInitialConditions(lambda: T0)
    dt = 0.1 * h * h / 2.0                                    # FEMSolver.py:138-144 on the uniform mesh (rho c_v = k = 1)
    fs = K.FEMSolver(analysis_type="transient", partition="rows", exchange=args.exchange, number_of_increments=nt + 1,
                     theta=theta, time_step=dt, verbose=False)
    solver = K.LinearSolver(tol=args.rtol, maxiter=2000)

    def step():
        return fs.Solve(formulation=form, mesh=mesh, material=material, boundary_condition=bc, solver=solver)

    torch.cuda.synchronize(); t0 = time.perf_counter()
    out = step()
    torch.cuda.synchronize(); t_setup = time.perf_counter() - t0
    for _ in range(max(0, warmup - 1)):
        out = step()
    sampler = ClockSampler(local); sampler.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    launches0 = lib.kb_launch_count()
    e0.record()
    loop_ms = 0.0
    for _ in range(steps):
        out = step()
        loop_ms += fs.timings["solve_ms"]
    e1.record()
    barrier()
    launches = lib.kb_launch_count() - launches0
    clocks = sampler.stop()
    info = dict(fs.solver_info)
    t = torch.tensor([e0.elapsed_time(e1), loop_ms], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms, loop_ms = float(t[0].item()), float(t[1].item())
    if info["status"] != 0:
        raise SystemExit("bench.py: the transient loop failed (status %d at step %d)" % (info["status"], info["failed_step"]))
    T = out.sol_device[:, 0, 0]
    rng = torch.stack([T.min(), -T.max()])
    if world > 1:
        dist.all_reduce(rng, op=dist.ReduceOp.MIN)
    ndof = nxn * nyn
    line = None
    if True:
        line = {"metric": "transient DOF*steps/s, 2D conv-diff theta-method (config 5)", "value": ndof * nt * steps / (ms * 1e-3),
                "unit": "DOF*steps/s", "n_gpus": world, "steps": steps, "warmup": warmup,
                "ms_per_step": ms / steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
                "dtype": "f64", "data": "synthetic",
                "config": {"workload": "config 5: synthetic 2D structured quad mesh %dx%d nodes (%d DOF), transient "
                                       "convection-diffusion, theta = %g, %d time steps per bench step (assembly of K and M + "
                                       "boundary conditions + the whole time loop), inner Jacobi-%s rtol %.0e" % (
                                           nxn, nyn, ndof, theta, nt, info["method"].upper(), args.rtol),
                           "ndof": ndof, "ndof_per_gpu": nxn * nps, "theta": theta, "time_steps": nt, "dt": dt,
                           "l2_policy": "operators (2 x 1.15 GB + vectors per GPU) far larger than the 126 MB L2"},
                "time_loop": {"ms_per_time_step": loop_ms / steps / nt, "krylov_iterations_per_time_step":
                              info["iterations"] / nt, "method": info["method"], "exchange": info.get("exchange"),
                              "worst_relative_residual": info["worst_relative_residual"], "status": info["status"],
                              "dof_steps_per_s_loop_only": ndof * nt / (loop_ms / steps * 1e-3)},
                "verify": {"T_min": float(rng[0].item()), "T_max": float(-rng[1].item()),
                           "note": "bounded by the initial / boundary data up to the explicit scheme's small overshoot; parity "
                                   "of the partitioned loop against the single-GPU loop: tests/test_gpu_parity.py::"
                                   "test_peer_partitioned_transient_matches_single_gpu"},
                "t_setup_ms": t_setup * 1e3, "roofline": None, "cpu_baseline": None, "e2e": None,
                "gpu_launches": int(launches), "clocks": clocks}
    if getattr(fs, "_psys", None) is not None:
        KD.close_peer_regions([fs._psys])
    return line


# ------------------------------------------------------------------------------------------------ GPU arm
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="steady", choices=["steady", "transient"],
                    help="steady: the headline metric (BASELINE configs[3]); transient: config 5 (theta-method time loop)")
    ap.add_argument("--theta", type=float, default=0.0, help="--workload transient: theta of the time scheme")
    ap.add_argument("--time-steps", type=int, default=50, help="--workload transient: time steps per bench step")
    ap.add_argument("--config5-time-steps", type=int, default=50,
                    help="steady workload: time steps of the bounded config-5 (transient) record embedded in the line "
                         "(0 = skip)")
    ap.add_argument("--nodes-per-side", type=int, default=4000)
    ap.add_argument("--rtol", type=float, default=1.0e-10)
    ap.add_argument("--check-every", type=int, default=100)
    ap.add_argument("--maxiter", type=int, default=None, help="cap Krylov iterations (profiling runs only)")
    ap.add_argument("--e2e-steps", type=int, default=2)
    ap.add_argument("--ref-nodes-per-side", type=int, default=101,
                    help="reference arm / cpu_baseline: sample size of the stock (unmodified) reference, nodes per side")
    ap.add_argument("--ref-stubbed-nodes-per-side", type=int, default=401,
                    help="one extra reference step with the dead dense Robin pass stubbed, nodes per side (0 = skip)")
    ap.add_argument("--cpu-baseline-steps", type=int, default=3)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-verify", dest="verify", action="store_false", help="skip the verify block (profiling runs)")
    ap.add_argument("--exchange", default="peer", choices=["peer", "nccl"],
                    help="N > 1: halo rows / dot products through NVLink peer memory written by the kernels (default) or NCCL")
    ap.add_argument("--profiler-range", action="store_true", help="cudaProfilerStart/Stop around the timed steps (N = 1), so "
                    "that `ncu --profile-from-start off` lists exactly the launches of the timed region")
    ap.add_argument("--no-multigrid", dest="multigrid", action="store_false",
                    help="skip the value_multigrid leg (the opt-in multigrid-preconditioned solver, reported beside the headline)")
    ap.add_argument("--idx32", action="store_true", help="keep 32-bit column indices in the solver's SpMV (default: 16-bit "
                    "offsets when the matrix band fits)")
    args = ap.parse_args()
    claim_stdout()

    rank = int(os.environ.get("RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference_arm(args, rank, world)
        return

    import numpy as np
    import torch
    import torch.distributed as dist
    import kiltro_b200 as K
    from kiltro_b200 import _lib

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: kiltro_b200 has no CPU fallback")
    torch.cuda.set_device(local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")   # keep stdout to the one JSON line (NCCL logs its version)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    lib = _lib.load()
    if args.idx32:
        lib.kb_spmv_force_idx32(1)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    if args.workload == "transient":
        run_transient_workload(args, K, torch, dist, lib, _lib, rank, world, local, barrier)
        if world > 1:
            dist.destroy_process_group()
        return
    if world > 1:
        run_partitioned(args, K, torch, dist, lib, _lib, rank, world, local, barrier)
        dist.destroy_process_group()
        return

    nps = args.nodes_per_side
    mesh, vel, flags, material = make_problem(K, torch, nps, seed=rank)
    nn = mesh.nnodes
    form = K.HeatAdvectionDiffusion(mesh, velocity=vel)
    bc = K.BoundaryCondition(); bc.SetDirichletCriteria(lambda: flags)
    fs = K.FEMSolver(analysis_type="steady", verbose=False)
    solver = K.LinearSolver(linear_solver="iterative", linear_solver_type="bicgstab", tol=args.rtol,
                            check_every=args.check_every, maxiter=args.maxiter)
    torch.cuda.synchronize(); t0 = time.perf_counter()
    sp = fs.ComputeSparsityFEM(mesh, form)                       # symbolic phase: once per mesh, reported separately
    torch.cuda.synchronize(); t_symbolic = time.perf_counter() - t0

    def step():
        return fs.Solve(formulation=form, mesh=mesh, material=material, boundary_condition=bc, solver=solver)

    for _ in range(args.warmup):
        out = step()
    iters = solver.info["iterations"]
    # ---- timed region: inputs resident in HBM
    sampler = ClockSampler(local); sampler.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    _lib.check(lib.kb_profile_spmv_begin(8))
    barrier()
    launches0 = lib.kb_launch_count()
    if args.profiler_range:
        torch.cuda.cudart().cudaProfilerStart()
    e0.record()
    phase = {"assemble_ms": 0.0, "bc_ms": 0.0, "solve_ms": 0.0}
    for _ in range(args.steps):
        out = step()
        for k in phase:
            phase[k] += fs.timings[k]
    e1.record()
    barrier()
    if args.profiler_range:
        torch.cuda.cudart().cudaProfilerStop()
    launches = lib.kb_launch_count() - launches0
    ns, tot = _lib.I64(0), _lib.F64(0.0)
    lib.kb_profile_spmv_end(ns, tot)
    clocks = sampler.stop()
    ms = e0.elapsed_time(e1)
    tmax = torch.tensor([ms], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
    ms = float(tmax.item())
    value = world * nn * args.steps / (ms * 1e-3)
    info = dict(solver.info)

    # ---- roofline of the dominant kernel (k_spmv inside the BiCGSTAB loop), timed live with CUDA events
    peaks, peak_src = None, "fallback"
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            peaks = json.load(f)
        peak = float(peaks["hbm_gbs"]); peak_src = "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        peak = 6650.0; peak_src = "fallback 6.65 TB/s (B200_PROFILING.md)"
    nfree = bc.n_free
    nnz_b = None
    roof = None
    if ns.value > 0:
        # the solver iterates on the reduced system K_b; its nnz is known from the last BC pass
        nnz_b = int(fs.last_reduced_nnz) if hasattr(fs, "last_reduced_nnz") else sp.nnz
        spmv_ms = tot.value / ns.value
        achieved = algorithmic_bytes_spmv(nfree, nnz_b) / (spmv_ms * 1e-3) * 1e-9
        traffic, traffic_src = None, None
        try:                                              # dram bytes per launch from the committed ncu capture
            with open(os.path.join(ROOT, "profiles", "spmv_traffic.json")) as f:
                tj = json.load(f)
            if nps == 4000:
                traffic, traffic_src = tj["traffic_bytes_per_launch"], tj["source"]
        except Exception:
            pass
        form_name = "32-bit indices forced" if args.idx32 else "row patterns: 8 B/nnz + 5 B/row"
        roof = {"bound": "hbm", "kernel": "k_spmv_pat (persistent TMA-fed CSR SpMV, row-pattern form, + fused dots, inside BiCGSTAB)"
                if not args.idx32 else "k_spmv_tma (persistent TMA-fed CSR SpMV + fused dots, inside BiCGSTAB)",
                "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": traffic,
                "traffic_source": traffic_src, "peak_source": peak_src,
                "avg_launch_ms": spmv_ms, "launches_timed": int(ns.value),
                "algorithmic_bytes_per_launch": algorithmic_bytes_spmv(nfree, nnz_b),
                "note": "achieved = SURVEY 8(d) algorithmic bytes (12 B/nnz + 20 B/row) / CUDA-event time of the SpMV launches "
                        "inside the timed solver loop; the kernel does not stream column indices at all when the rows of the "
                        "matrix repeat <= 256 offset patterns (%s here), so fewer bytes than the algorithmic count cross "
                        "HBM (see traffic) and frac can exceed 1" % form_name}

    # ---- second kernel of the path: numeric assembly (k_assemble_tmpl / k_assemble_fused), timed alone with CUDA events
    roof_asm = None
    if world == 1:
        for _ in range(3):
            K.Assemble(fs, form.function_spaces, form, mesh, material, bc)
        a0, a1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        n_asm = 20
        torch.cuda.synchronize()
        a0.record()
        for _ in range(n_asm):
            K.Assemble(fs, form.function_spaces, form, mesh, material, bc)
        a1.record(); torch.cuda.synchronize()
        asm_ms = a0.elapsed_time(a1) / n_asm
        asm_bytes = 16 * mesh.nelem + 32 * nn + 8 * sp.nnz + 4 * (nn + 1) + 4 * sp.nnz
        asm_traffic, asm_src = None, None
        try:
            with open(os.path.join(ROOT, "profiles", "assembly_traffic.json")) as f:
                tj = json.load(f)
            if nps == 4000 and tj.get("assembly_used") == fs.assembly_used:
                asm_traffic, asm_src = tj["traffic_bytes_per_launch"], tj["source"]
        except Exception:
            pass
        roof_asm = {"bound": "hbm", "kernel": "numeric assembly (%s): element matrices on chip -> CSR values" % fs.assembly_used,
                    "achieved": asm_bytes / (asm_ms * 1e-3) * 1e-9, "peak": peak, "unit": "GB/s",
                    "frac": asm_bytes / (asm_ms * 1e-3) * 1e-9 / peak, "traffic": asm_traffic, "traffic_source": asm_src,
                    "peak_source": peak_src, "avg_launch_ms": asm_ms, "launches_timed": n_asm,
                    "algorithmic_bytes_per_launch": asm_bytes,
                    "note": "Assemble() calls back to back (output allocation included), CUDA events; algorithmic bytes = "
                            "16 nelem + 32 nnode + 12 nnz + 4 (nnode+1), SURVEY 8(d); " + (
                                "grid kernel: element matrices in registers, exchanged by warp shuffles; connectivity and "
                                "column indices are not read (DESIGN 4.2)" if fs.assembly_used == "fused_grid" else
                                "tile kernels: bound by the shared-memory pipe and exposed latency at 16 warps/SM (DESIGN 4.2)")}

    # ---- e2e: same step from pinned host buffers in the reference's array formats
    e2e = None
    if args.e2e_steps > 0:
        h_pts = mesh._points_d.cpu().pin_memory()
        h_el = mesh._elements_d.to(torch.int64).cpu().pin_memory()
        h_vel = vel.cpu().pin_memory()
        h_flags = flags.cpu().pin_memory()
        h_sol = torch.empty((nn, 1, 1), dtype=torch.float64).pin_memory()
        h2d = h_pts.numel() * 8 + h_el.numel() * 8 + h_vel.numel() * 8 + h_flags.numel() * 8
        d2h = h_sol.numel() * 8

        def e2e_step():
            m = K.Mesh(element_type="quad")
            m.set_arrays(h_pts, h_el)                            # host -> device inside
            f = K.HeatAdvectionDiffusion(m, velocity=h_vel)
            b = K.BoundaryCondition(); b.SetDirichletCriteria(lambda: h_flags)
            o = fs.Solve(formulation=f, mesh=m, material=material, boundary_condition=b, solver=solver)
            h_sol.copy_(o.sol_device, non_blocking=True)
            torch.cuda.current_stream().synchronize()
            return o

        e2e_step()                                               # warm-up
        barrier()
        t0 = time.perf_counter()
        ee0, ee1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ee0.record()
        for _ in range(args.e2e_steps):
            e2e_step()
        ee1.record()
        barrier()
        wall = time.perf_counter() - t0
        ems = max(ee0.elapsed_time(ee1), wall * 1e3)
        t2 = torch.tensor([ems], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(t2, op=dist.ReduceOp.MAX)
        e2e = {"value": world * nn * args.e2e_steps / (float(t2.item()) * 1e-3), "unit": UNIT,
               "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h), "steps": args.e2e_steps,
               "ms_per_step": float(t2.item()) / args.e2e_steps}

    # ---- verify: correctness evidence of exactly the solution the timed steps produced (outside the timed region).
    # The residual is recomputed from scratch -- fresh assembly, fresh Dirichlet reduction, one plain SpMV on the
    # UNSCALED reduced system -- so it does not reuse anything of the solver's own bookkeeping.
    verify = None
    if args.verify:
        Kv, _, _ = K.Assemble(fs, form.function_spaces, form, mesh, material, bc)
        Fv = torch.zeros((nn, 1), dtype=torch.float64, device="cuda")
        Kb, Fb = bc.ApplyDirichletGetReducedMatrices(Kv, Fv, bc.applied_dirichlet_device)[:2]
        T = out.sol_device[:, 0, 0].contiguous()
        xf = T[bc.columns_in_device]
        res = Fb - Kb.dot(xf)
        rel = float(torch.linalg.vector_norm(res) / torch.linalg.vector_norm(Fb))
        td = T[bc.columns_out_device]
        verify = {"true_relative_residual": rel, "rtol": args.rtol,
                  "residual_ok": bool(rel <= 10.0 * args.rtol),
                  "dirichlet_max_error": float((td - bc.applied_dirichlet_device).abs().max()),
                  "solution_range": [float(T.min()), float(T.max())],
                  "note": "||F_b - K_b x|| / ||F_b|| recomputed from a fresh assembly + Dirichlet reduction with one plain SpMV "
                          "(unscaled system); parity of this very case against the reference's direct solve at 511^2 / 999^2 "
                          "nodes and rtol 1e-10 vs 1e-13 at full size: tests/test_gpu_benchcase.py"}
        del Kv, Kb, Fb, Fv, res

    # ---- beside the headline, never instead of it: the same step with the OPT-IN multigrid-preconditioned BiCGSTAB
    # (csrc/multigrid.cu, csrc/krylov_mg.cuh).  Same problem, same tolerance, same FEMSolver.Solve call; the numeric set-up
    # of the hierarchy (level-0 stencils from K_b, rediscretised coarse operators, coarsest inverse) is redone in every
    # step, inside the timed region, like the assembly.
    mgleg = None
    if args.multigrid and args.maxiter is None:
        from kiltro_b200.multigrid import GridMultigrid
        mg = GridMultigrid(mesh, form, material, bc)
        mg_solver = K.LinearSolver(tol=args.rtol, preconditioner=mg, maxiter=200)

        def mg_step():
            return fs.Solve(formulation=form, mesh=mesh, material=material, boundary_condition=bc, solver=mg_solver)

        for _ in range(max(2, args.warmup)):
            mg_out = mg_step()
        m0, m1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        barrier()
        mg_launches0 = lib.kb_launch_count()
        m0.record()
        mg_phase = {"assemble_ms": 0.0, "bc_ms": 0.0, "solve_ms": 0.0}
        for _ in range(args.steps):
            mg_out = mg_step()
            for k in mg_phase:
                mg_phase[k] += fs.timings[k]
        m1.record()
        barrier()
        mg_ms = m0.elapsed_time(m1)
        mg_launches = lib.kb_launch_count() - mg_launches0
        minfo = dict(mg_solver.info)
        Tm = mg_out.sol_device[:, 0, 0].contiguous()
        Tj = out.sol_device[:, 0, 0]
        mg_verify = {"rel_l2_vs_jacobi_bicgstab_solution": float(torch.linalg.vector_norm(Tm - Tj) / torch.linalg.vector_norm(Tj))}
        Kv, _, _ = K.Assemble(fs, form.function_spaces, form, mesh, material, bc)
        Fv = torch.zeros((nn, 1), dtype=torch.float64, device="cuda")
        Kb, Fb = bc.ApplyDirichletGetReducedMatrices(Kv, Fv, bc.applied_dirichlet_device)[:2]
        mg_verify["true_relative_residual"] = float(torch.linalg.vector_norm(Fb - Kb.dot(Tm[bc.columns_in_device])) /
                                                    torch.linalg.vector_norm(Fb))
        mg_verify["residual_ok"] = bool(mg_verify["true_relative_residual"] <= 10.0 * args.rtol)
        # set-up and one V-cycle alone (CUDA events), for the cost model in DESIGN 6c
        mg.setup(Kb)
        s0, s1, s2, s3 = (torch.cuda.Event(enable_timing=True) for _ in range(4))
        s0.record(); mg.setup(Kb); s1.record()
        rr = Fb.reshape(-1).clone()
        mg.apply(rr)
        s2.record()
        for _ in range(10):
            mg.apply(rr)
        s3.record(); torch.cuda.synchronize()
        del Kv, Kb, Fb, Fv
        mg_e2e = None
        if args.e2e_steps > 0:
            def mg_e2e_step():
                m = K.Mesh(element_type="quad")
                m.set_arrays(h_pts, h_el)
                f = K.HeatAdvectionDiffusion(m, velocity=h_vel)
                b = K.BoundaryCondition(); b.SetDirichletCriteria(lambda: h_flags)
                sv = K.LinearSolver(tol=args.rtol, preconditioner=GridMultigrid(m, f, material, b), maxiter=200)
                o = fs.Solve(formulation=f, mesh=m, material=material, boundary_condition=b, solver=sv)
                h_sol.copy_(o.sol_device, non_blocking=True)
                torch.cuda.current_stream().synchronize()
                return sv

            mg_e2e_step()
            barrier()
            t0 = time.perf_counter()
            for _ in range(args.e2e_steps):
                sv = mg_e2e_step()
            barrier()
            wall = time.perf_counter() - t0
            mg_e2e = {"value": nn * args.e2e_steps / wall, "unit": UNIT, "h2d_bytes_per_step": int(h2d),
                      "d2h_bytes_per_step": int(d2h), "steps": args.e2e_steps, "ms_per_step": wall * 1e3 / args.e2e_steps,
                      "status": sv.info["status"], "iterations": sv.info["iterations"]}
        # algorithmic bytes of one V-cycle (fp32 levels): per node of a level the down-leg moves the 8 stencil
        # coefficients, d, the rhs and writes x and the residual (+ on level 0 the fp64 vector in, the node map, 1/d and the
        # staged rhs out), the restriction reads the residual, the up-leg moves the coefficients, rhs and x in and x out (on
        # level 0 the node map and the fp64 result instead)
        vc_bytes, visits = 0, 1
        for li, (lx, ly) in enumerate(mg.levels[:-1]):
            npad = ly * ((lx + 1 + 31) // 32 * 32)
            down = (32 + 4 + 4 + 4 + 4) * npad + ((8 + 4 + 4 + 4) * npad if li == 0 else 0)
            up = (32 + 4 + 4) * npad + ((4 + 8) * npad if li == 0 else 4 * npad)
            reps = mg.cycle[li] if li < len(mg.cycle) else 1
            # a repeated coarse-grid correction adds a prolongation (x in / out + coarse x), a residual (coefficients, d, rhs,
            # x in, r out) and a restriction per repetition
            extra = (reps - 1) * ((4 + 4 + 1) * npad + (32 + 4 + 4 + 4 + 4) * npad + 4 * npad)
            vc_bytes += visits * (down + 4 * npad + up + extra)
            visits *= reps
        vc_ms = s2.elapsed_time(s3) / 10
        mg_roof = {"bound": "hbm", "kernel": "one multigrid cycle (k_mg_leg down / up legs + k_mg_restrict per level, nu = %d, coarse-grid corrections "
                                       "per level %s)" % (mg.nu, list(mg.cycle)),
                   "achieved": vc_bytes / (vc_ms * 1e-3) * 1e-9, "peak": peak, "unit": "GB/s",
                   "frac": vc_bytes / (vc_ms * 1e-3) * 1e-9 / peak, "algorithmic_bytes_per_cycle": int(vc_bytes),
                   "ms_per_cycle": vc_ms, "peak_source": peak_src,
                   "note": "whole cycle (all levels), CUDA events over 10 back-to-back cycles; the levels below 251^2 nodes are "
                           "latency-bound chains of small kernels"}
        mgleg = {"value": nn * args.steps / (mg_ms * 1e-3), "unit": UNIT, "ms_per_step": mg_ms / args.steps,
                 "roofline": mg_roof, "nu": mg.nu, "omega": mg.omega, "cycle": list(mg.cycle),
                 "steps": args.steps, "method": minfo["method"], "iterations": minfo["iterations"], "levels": minfo["levels"],
                 "level_shapes": mg.levels, "status": minfo["status"], "relative_residual": minfo["relative_residual"],
                 "rtol": args.rtol, "phases_ms_per_step": {k: v / args.steps for k, v in mg_phase.items()},
                 "setup_ms": s0.elapsed_time(s1), "vcycle_ms": s2.elapsed_time(s3) / 10, "gpu_launches": int(mg_launches),
                 "verify": mg_verify, "e2e": mg_e2e,
                 "note": "opt-in LinearSolver(preconditioner=GridMultigrid(...)): BiCGSTAB right-preconditioned by one V-cycle "
                         "(fp32 nine-point stencil levels, rediscretised Petrov-Galerkin coarse operators), outer recurrence and "
                         "verdict in fp64 on the caller's matrix; solve_ms includes the numeric set-up of the hierarchy; parity "
                         "vs the oracle's direct solve: tests/test_gpu_multigrid.py.  The headline `value` stays the prescribed "
                         "Jacobi-BiCGSTAB"}
        if minfo["status"] != 0:
            mgleg["value"] = None

    profiling_run = args.maxiter is not None          # --maxiter caps the iterations on purpose (ncu launch lists / captures)
    if info["status"] != 0 and profiling_run:
        value = None                                   # no throughput is claimed for a capped solve
    if info["status"] != 0 and not profiling_run:
        raise SystemExit("bench.py: the timed solve did not converge (status %d, relative residual %.3e) -- no throughput "
                         "record is emitted for a failed solve" % (info["status"], info["relative_residual"]))

    # ---- BASELINE config 5 beside the headline (bounded: `--config5-time-steps` time steps of the theta = 0 loop on the same
    # mesh through FEMSolver(analysis_type='transient', partition='rows') with one strip; stand-alone: `--workload transient`)
    config5 = None
    if args.config5_time_steps > 0 and not profiling_run:
        tl = measure_transient(args, K, torch, dist, lib, rank, world, local, barrier, 1, 1, args.config5_time_steps, 0.0)
        config5 = {k: tl[k] for k in ("metric", "value", "unit", "ms_per_step", "config", "time_loop", "verify", "t_setup_ms")}

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        cpu = cpu_baseline_block(args.ref_nodes_per_side, args.cpu_baseline_steps, args.ref_stubbed_nodes_per_side)

    if rank == 0:
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
                "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak",
                "vs_baseline": None, "dtype": "f64", "data": "synthetic",
                "config": gpu_config(nps, world, args.rtol, False),
                "nnz": sp.nnz, "parallelism": "1 GPU",
                "solver": {"method": info["method"], "iterations": info["iterations"],
                           "relative_residual": info["relative_residual"], "status": info["status"],
                           "attainable_accuracy": info.get("attainable_accuracy"), "restarts": info.get("restarts"),
                           "ms_per_iteration": (phase["solve_ms"] / args.steps) / max(1, info["iterations"])},
                "verify": verify,
                "phases_ms_per_step": {k: v / args.steps for k, v in phase.items()},
                "t_symbolic_ms": t_symbolic * 1e3,
                "roofline": roof, "roofline_assembly": roof_asm, "cpu_baseline": cpu, "e2e": e2e, "gpu_launches": int(launches), "clocks": clocks,
                "value_multigrid": mgleg, "config5_transient": config5}
        if profiling_run:
            line["profiling_run"] = "--maxiter %d: iterations capped for a profiler, `value` withheld unless the solve converged" % args.maxiter
        emit(line)


if __name__ == "__main__":
    main()
