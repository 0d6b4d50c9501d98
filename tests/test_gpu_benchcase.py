"""Parity pinned on the BENCHMARKED case: bench.py's own generator --
cell Peclet 0.45 mean flow + 10 % hash noise, HeatAdvectionDiffusion, Dirichlet x=0 -> 100, x=1 -> 500 -- at the
oracle-sized n = 511 / 999 against oracle.steady_solve (the restatement of kiltro/FEMSolver.py:156-228,350-364 with the
reference's direct solve), at the bench's rtol 1e-10 AND at 1e-13; and at the full S16 size through self-consistency of the
two tolerances plus an independently recomputed true residual.  Tolerances: CSR pattern bit-exact, values <= 1e-12
relative, solution <= 1e-8 relative L2 (north_star)."""
import numpy as np
import pytest

from conftest import rel_l2, rel_max

pytestmark = pytest.mark.gpu
torch = pytest.importorskip("torch")

VAL_TOL = 1e-12
SOL_TOL = 1e-8


def _solve_device(K, bench, nps, rtol, fs=None, cache=None):
    """One bench step at nps x nps nodes through the reference-shaped API, on the bench's own device generator."""
    if cache is None:
        cache = {}
    if "mesh" not in cache:
        cache["mesh"], cache["vel"], cache["flags"], cache["mat"] = bench.make_problem(K, torch, nps)
    mesh, vel, flags, mat = cache["mesh"], cache["vel"], cache["flags"], cache["mat"]
    form = K.HeatAdvectionDiffusion(mesh, velocity=vel)
    bc = K.BoundaryCondition(); bc.SetDirichletCriteria(lambda: flags)
    fs = fs or K.FEMSolver(analysis_type="steady", verbose=False)
    solver = K.LinearSolver(linear_solver="iterative", linear_solver_type="bicgstab", tol=rtol, check_every=100,
                            maxiter=200000)
    out = fs.Solve(formulation=form, mesh=mesh, material=mat, boundary_condition=bc, solver=solver)
    return out, solver, fs, form, bc, cache


def _true_relative_residual(K, fs, form, mesh, mat, bc, T):
    """||F_b - K_b x|| / ||F_b|| from a fresh assembly + Dirichlet reduction and one plain SpMV (unscaled system)."""
    Kc = K.Assemble(fs, form.function_spaces, form, mesh, mat, bc)[0]
    F = torch.zeros((mesh.nnodes, 1), dtype=torch.float64, device="cuda")
    Kb, Fb = bc.ApplyDirichletGetReducedMatrices(Kc, F, bc.applied_dirichlet_device)[:2]
    x = T[bc.columns_in_device]
    return float(torch.linalg.vector_norm(Fb - Kb.dot(x)) / torch.linalg.vector_norm(Fb))


@pytest.mark.parametrize("n", [511, 999])
def test_bench_case_vs_oracle(n):
    """n x n ELEMENTS (nx = ny in {127, 511, 999}; 127 is covered by test_steady_vs_oracle_spsolve)."""
    import bench
    import kiltro_b200 as K
    from oracle import kiltro_oracle as O
    nps = n + 1
    # host side of the same generator: bit-identical inputs (integer hash), checked below
    pts, vel, d = bench.host_problem(nps)
    el = O.mesh_rectangle((0.0, 0.0), (1.0, 1.0), n, n)[1]
    out10, solver10, fs, form, bc, cache = _solve_device(K, bench, nps, 1e-10)
    mesh, mat = cache["mesh"], cache["mat"]
    assert np.array_equal(K.to_numpy(mesh.device_points()), pts)
    assert np.array_equal(K.to_numpy(cache["vel"]), vel)
    assert np.array_equal(K.to_numpy(mesh.device_elements()).astype(np.int64), np.asarray(el, dtype=np.int64))
    # oracle: pattern, values, direct solve
    nn = pts.shape[0]
    ind, ptr = O.sparsity_pattern_fast(el, nn)
    dl, dg = O.data_indices_fast(np.asarray(el), ind, ptr)
    A = O.assemble(pts, el, vel, (1.0, 1.0, 1.0, 1.0, 0.0), "consistent", False, False, (ind, ptr, dl, dg))
    Kc = K.Assemble(fs, form.function_spaces, form, mesh, mat, bc)[0]
    assert np.array_equal(K.to_numpy(Kc.indptr), ptr) and np.array_equal(K.to_numpy(Kc.indices), ind)    # bit-exact
    assert rel_max(K.to_numpy(Kc.data), A["K"]) < VAL_TOL
    cin, cout, td = O.dirichlet_split(d)
    kb = O.apply_dirichlet_reduce(ind, ptr, A["K"], np.zeros(nn), cin, cout, td)
    ref = O.total_component_sol(O.linear_solve(kb[0], kb[1], kb[2], kb[3]), cin, cout, td, nn)
    # the bench's tolerance
    assert solver10.info["status"] == 0, solver10.info
    assert solver10.info["relative_residual"] <= 1e-10, solver10.info
    e10 = rel_l2(out10.sol[:, 0, 0], ref)
    assert e10 < SOL_TOL, (e10, solver10.info)
    # the tight tolerance
    out13, solver13 = _solve_device(K, bench, nps, 1e-13, cache=cache)[:2]
    assert solver13.info["status"] == 0, solver13.info
    e13 = rel_l2(out13.sol[:, 0, 0], ref)
    assert e13 < SOL_TOL, (e13, solver13.info)
    print("bench case n=%d: rel L2 vs spsolve %.2e (rtol 1e-10, %d it), %.2e (rtol 1e-13, %d it)" % (
        n, e10, solver10.info["iterations"], e13, solver13.info["iterations"]))


def test_bench_case_fullsize_self_consistency():
    """S16 (4000 x 4000 nodes, the timed configuration): the rtol 1e-10 solution the bench times agrees with the rtol 1e-13
    one to <= 1e-8 relative L2, and its true residual -- recomputed from scratch -- meets the tolerance."""
    import bench
    import kiltro_b200 as K
    nps = 4000
    out10, solver10, fs, form, bc, cache = _solve_device(K, bench, nps, 1e-10)
    assert solver10.info["status"] == 0, solver10.info
    T10 = out10.sol_device[:, 0, 0].clone()
    rel = _true_relative_residual(K, fs, form, cache["mesh"], cache["mat"], bc, T10)
    assert rel <= 1e-10 * 1.01, (rel, solver10.info)           # 1 %: norm recomputed by a different reduction order
    out13, solver13 = _solve_device(K, bench, nps, 1e-13, fs=fs, cache=cache)[:2]
    T13 = out13.sol_device[:, 0, 0]
    # 1e-13 sits at the fp64 floor of a 16 M-DOF system: the verdict may be "attainable accuracy" (status 0 within 10x)
    assert solver13.info["relative_residual"] <= 1e-12, solver13.info
    diff = float(torch.linalg.vector_norm(T10 - T13) / torch.linalg.vector_norm(T13))
    assert diff <= SOL_TOL, (diff, solver10.info, solver13.info)
    # both respect the Dirichlet data and the discrete maximum principle loosely
    assert float(T10.min()) > 100.0 - 1e-6 and float(T10.max()) < 500.0 + 1e-6
    print("S16: ||x(1e-10) - x(1e-13)|| / ||x|| = %.2e; true residual %.2e; iterations %d / %d" % (
        diff, rel, solver10.info["iterations"], solver13.info["iterations"]))
