"""Opt-in multigrid-preconditioned BiCGSTAB (csrc/multigrid.cu, csrc/krylov_mg.cuh, kiltro_b200/multigrid.py),
reported beside the prescribed Jacobi method.  Parity bar = every other solve test's: solution <= 1e-8 relative L2 against
the ORACLE's direct solve (oracle.linear_solve restates kiltro/FEMSolver.py:350-364), on the bench's own generator at
n = 255 / 511 / 999 and on smaller meshes with other boundary sets; at the full S16 size through the
independently recomputed true residual and agreement with the Jacobi-BiCGSTAB solution."""
import numpy as np
import pytest

from conftest import rel_l2

pytestmark = pytest.mark.gpu
torch = pytest.importorskip("torch")

SOL_TOL = 1e-8


def _oracle_solution(pts, el, vel, flags_host, nn):
    from oracle import kiltro_oracle as O
    ind, ptr = O.sparsity_pattern_fast(el, nn)
    dl, dg = O.data_indices_fast(np.asarray(el), ind, ptr)
    mode = "consistent"
    A = O.assemble(pts, el, vel, (1.0, 1.0, 1.0, 1.0, 0.0), mode, False, False, (ind, ptr, dl, dg))
    cin, cout, td = O.dirichlet_split(flags_host)
    kb = O.apply_dirichlet_reduce(ind, ptr, A["K"], np.zeros(nn), cin, cout, td)
    return O.total_component_sol(O.linear_solve(kb[0], kb[1], kb[2], kb[3]), cin, cout, td, nn)


def _mg_solve(K, mesh, form, mat, flags, tol, **mg_kw):
    from kiltro_b200.multigrid import GridMultigrid
    bc = K.BoundaryCondition(); bc.SetDirichletCriteria(lambda: flags)
    mg = GridMultigrid(mesh, form, mat, bc, **mg_kw)
    solver = K.LinearSolver(tol=tol, preconditioner=mg, maxiter=200)
    fs = K.FEMSolver(analysis_type="steady", verbose=False)
    out = fs.Solve(formulation=form, mesh=mesh, material=mat, boundary_condition=bc, solver=solver)
    return out, solver, mg, fs, bc


@pytest.mark.parametrize("n", [255, 511, 999])
def test_multigrid_bench_case_vs_oracle(n):
    """The benchmarked problem (cell Peclet 0.45 + 10 % hash noise) at n x n elements against the oracle's direct solve."""
    import bench
    import kiltro_b200 as K
    from oracle import kiltro_oracle as O
    nps = n + 1
    pts, vel, d = bench.host_problem(nps)
    el = O.mesh_rectangle((0.0, 0.0), (1.0, 1.0), n, n)[1]
    ref = _oracle_solution(pts, el, vel, d, pts.shape[0])
    mesh, dvel, flags, mat = bench.make_problem(K, torch, nps)
    form = K.HeatAdvectionDiffusion(mesh, velocity=dvel)
    for tol in (1e-10, 1e-13):
        out, solver, mg, _, _ = _mg_solve(K, mesh, form, mat, flags, tol)
        info = solver.info
        assert info["status"] == 0 and info["method"] == "bicgstab+multigrid", info
        assert info["relative_residual"] <= tol * (10.0 if info["attainable_accuracy"] else 1.0), info
        err = rel_l2(out.sol[:, 0, 0], ref)
        assert err < SOL_TOL, (err, info)
        assert info["iterations"] <= 40, info
        print("multigrid bench case n=%d tol %.0e: rel L2 vs spsolve %.2e, %d iterations, %d levels" % (
            n, tol, err, info["iterations"], info["levels"]))


@pytest.mark.parametrize("nx,ny,peclet,sides", [(128, 96, 0.45, "lr"), (99, 81, 0.45, "lr"), (128, 128, 0.0, "lr"),
                                                (257, 190, 0.3, "all"), (64, 200, 0.45, "bt"), (37, 23, 0.2, "lr")])
def test_multigrid_small_cases_vs_oracle(nx, ny, peclet, sides):
    """Non-nested hierarchies, pure diffusion (no velocity array), Dirichlet data on other sides, anisotropic node counts."""
    import kiltro_b200 as K
    from oracle import kiltro_oracle as O
    m = K.Mesh(element_type="quad"); m.Rectangle((0.0, 0.0), (1.0, 0.8), nx, ny)
    P = m.device_points()
    nn = m.nnodes
    g = torch.Generator(device="cuda"); g.manual_seed(1)
    flags = torch.full((nn, 1), float("nan"), dtype=torch.float64, device="cuda")
    if sides in ("lr", "all"):
        flags[P[:, 0] == 0.0, 0] = 100.0
        flags[P[:, 0] == 1.0, 0] = 500.0
    if sides in ("bt", "all"):
        flags[P[:, 1] == 0.0, 0] = 200.0
        flags[P[:, 1] == 0.8, 0] = 300.0
    mat = K.FourierConduction(2, k=1.0, area=1.0, density=1.0, heat_capacity_v=1.0, q=0.0)
    if peclet > 0.0:
        U = peclet * 2.0 * nx
        vel = torch.zeros((nn, 2), dtype=torch.float64, device="cuda"); vel[:, 0] = U
        vel += 0.1 * U * torch.randn((nn, 2), dtype=torch.float64, device="cuda", generator=g)
        form = K.HeatAdvectionDiffusion(m, velocity=vel)
        vel_h = K.to_numpy(vel)
    else:
        form = K.HeatDiffusion(m)
        vel_h = np.zeros((nn, 2))
    out, solver, mg, _, _ = _mg_solve(K, m, form, mat, flags, 1e-12)
    info = solver.info
    assert info["status"] == 0, info
    assert len(mg.levels) >= 2
    ref = _oracle_solution(K.to_numpy(P), K.to_numpy(m.device_elements()).astype(np.int64), vel_h,
                           K.to_numpy(flags).reshape(-1, 1), nn)
    err = rel_l2(out.sol.ravel(), ref)
    assert err < SOL_TOL, (err, info)
    # far fewer iterations than the prescribed method on the same problem
    bc = K.BoundaryCondition(); bc.SetDirichletCriteria(lambda: flags)
    jac = K.LinearSolver(tol=1e-12)
    K.FEMSolver(verbose=False).Solve(formulation=form, mesh=m, material=mat, boundary_condition=bc, solver=jac)
    assert info["iterations"] <= 40 and info["iterations"] * 3 < jac.info["iterations"], (info, jac.info)


@pytest.mark.parametrize("nx,ny,graded", [(6, 5, False), (9, 9, False), (120, 90, True), (300, 40, True)])
def test_multigrid_tiny_and_graded_meshes_vs_oracle(nx, ny, graded):
    """Hierarchies of one or two levels (the whole problem fits the dense coarsest solve) and grid-NUMBERED meshes with
    graded, non-uniform coordinates handed over as arrays: the coarse levels sample the finer level's nodes, so the
    preconditioner follows the geometry; the solution is held to the same bar."""
    import kiltro_b200 as K
    from oracle import kiltro_oracle as O
    pts, el = O.mesh_rectangle((0.0, 0.0), (1.0, 0.6), nx, ny)
    pts = np.array(pts, dtype=np.float64)
    if graded:                                       # cluster the nodes towards x = 1 and y = 0 (boundary-layer style)
        pts[:, 0] = 1.0 - (1.0 - pts[:, 0]) ** 1.6
        pts[:, 1] = 0.6 * (pts[:, 1] / 0.6) ** 1.4
    nn = pts.shape[0]
    m = K.Mesh(element_type="quad"); m.set_arrays(pts, np.asarray(el, dtype=np.int64))
    rng = np.random.default_rng(nx)
    vel = np.array([0.4 * nx, 0.1 * nx]) + 0.05 * nx * rng.standard_normal((nn, 2))
    d = np.full((nn, 1), np.nan); d[pts[:, 0] == 0.0] = 100.0; d[pts[:, 0] == 1.0] = 500.0
    flags = torch.as_tensor(d, device="cuda")
    mat = K.FourierConduction(2, k=1.0, area=1.0, density=1.0, heat_capacity_v=1.0, q=0.0)
    form = K.HeatAdvectionDiffusion(m, velocity=vel)
    out, solver, mg, _, _ = _mg_solve(K, m, form, mat, flags, 1e-12)
    info = solver.info
    assert info["status"] == 0 and info["iterations"] <= 60, info
    ref = _oracle_solution(pts, np.asarray(el), vel, d, nn)
    err = rel_l2(out.sol.ravel(), ref)
    assert err < SOL_TOL, (err, info, mg.levels)
    print("multigrid %dx%d elements (graded=%s): %d levels, %d iterations, rel L2 vs spsolve %.2e" % (
        nx, ny, graded, len(mg.levels), info["iterations"], err))


def test_vcycle_is_a_contraction_and_linear():
    """One V-cycle z = M^-1 r: ||r - A z|| / ||r|| well below 1 on a random residual, and M^-1 (a r1 + r2) = a M^-1 r1 + M^-1 r2
    to fp32 accuracy (BiCGSTAB needs a fixed linear preconditioner)."""
    import bench
    import kiltro_b200 as K
    from kiltro_b200.multigrid import GridMultigrid
    mesh, vel, flags, mat = bench.make_problem(K, torch, 301)
    form = K.HeatAdvectionDiffusion(mesh, velocity=vel)
    bc = K.BoundaryCondition(); bc.SetDirichletCriteria(lambda: flags)
    fs = K.FEMSolver(verbose=False)
    Kc = K.Assemble(fs, form.function_spaces, form, mesh, mat, bc)[0]
    bc.GetDirichletBoundaryConditions(form, mesh, mat, fs)
    F = torch.zeros((mesh.nnodes, 1), dtype=torch.float64, device="cuda")
    Kb, Fb = bc.ApplyDirichletGetReducedMatrices(Kc, F, bc.applied_dirichlet_device)[:2]
    mg = GridMultigrid(mesh, form, mat, bc)
    mg.setup(Kb)
    g = torch.Generator(device="cuda"); g.manual_seed(3)
    n = Kb.shape[0]
    r1 = torch.randn(n, dtype=torch.float64, device="cuda", generator=g)
    r2 = Fb.reshape(-1).clone()
    z1, z2 = mg.apply(r1), mg.apply(r2)
    for r, z in ((r1, z1), (r2, z2)):
        red = float(torch.linalg.vector_norm(r - Kb.dot(z)) / torch.linalg.vector_norm(r))
        assert red < 0.6, red
    z12 = mg.apply(0.37 * r1 + r2)
    lin = float(torch.linalg.vector_norm(z12 - (0.37 * z1 + z2)) / torch.linalg.vector_norm(z12))
    assert lin < 1e-4, lin
    # deterministic
    assert torch.equal(mg.apply(r1), z1)


@pytest.mark.parametrize("nps,nu,omega,cycle", [(301, 2, 0.67, ()), (301, 1, 0.67, ()), (301, 3, 0.67, ()), (130, 2, 0.6, ()),
                                                (47, 3, 0.67, ()), (130, 3, (0.6, 0.67, 0.55), ()), (301, 3, 0.67, (1, 2, 3)),
                                                (130, 2, 0.67, (2, 1, 2, 2))])
def test_fused_legs_bitwise(nps, nu, omega, cycle):
    """The fused legs (one kernel per leg and level: temporal blocking in registers) reproduce the one-kernel-per-sweep
    cycle bit for bit, for every task height, with and without the s = r - alpha v update riding in the staging kernel,
    for V-cycles and for schedules that repeat the coarse-grid correction of some levels."""
    import bench
    import kiltro_b200 as K
    from kiltro_b200 import _lib
    from kiltro_b200.multigrid import GridMultigrid
    lib = _lib.load()
    mesh, vel, flags, mat = bench.make_problem(K, torch, nps)
    form = K.HeatAdvectionDiffusion(mesh, velocity=vel)
    bc = K.BoundaryCondition(); bc.SetDirichletCriteria(lambda: flags)
    fs = K.FEMSolver(verbose=False)
    Kc = K.Assemble(fs, form.function_spaces, form, mesh, mat, bc)[0]
    bc.GetDirichletBoundaryConditions(form, mesh, mat, fs)
    F = torch.zeros((mesh.nnodes, 1), dtype=torch.float64, device="cuda")
    Kb, Fb = bc.ApplyDirichletGetReducedMatrices(Kc, F, bc.applied_dirichlet_device)[:2]
    mg = GridMultigrid(mesh, form, mat, bc, nu=nu, omega=omega, cycle=cycle)
    mg.setup(Kb)
    g = torch.Generator(device="cuda"); g.manual_seed(5)
    r = torch.randn(Kb.shape[0], dtype=torch.float64, device="cuda", generator=g)
    try:
        lib.kb_mg_set_fused(0)
        ref = mg.apply(r)
        x_ref, info_ref = mg.solve(Kb, Fb.reshape(-1), rtol=1e-11, reuse_setup=True)
        lib.kb_mg_set_fused(1)
        for rows in (64, 8, 23, 1000):
            lib.kb_mg_set_rows_per_task(rows)
            assert torch.equal(mg.apply(r), ref), rows
        lib.kb_mg_set_tail(0)                                  # the coarsest levels as separate kernels instead of k_mg_tail
        assert torch.equal(mg.apply(r), ref)
        lib.kb_mg_set_tail(1)
        lib.kb_mg_set_rows_per_task(0)
        x, info = mg.solve(Kb, Fb.reshape(-1), rtol=1e-11, reuse_setup=True)
        assert info["iterations"] == info_ref["iterations"] and info["status"] == 0, (info, info_ref)
        assert torch.equal(x, x_ref)
    finally:
        lib.kb_mg_set_fused(1); lib.kb_mg_set_rows_per_task(0); lib.kb_mg_set_tail(1)


@pytest.mark.parametrize("nps,peclet_scale", [(301, 1.0), (130, 0.0), (517, 2.0)])
def test_coarse_assembly_tiled_matches_per_node(nps, peclet_scale):
    """The tiled rediscretisation of the coarse operators (every element evaluated once per CTA tile, gathered from shared
    memory) against the per-node form (every element evaluated by each of its four nodes): same sums in the same order;
    the cycles built on them agree to fp32 rounding (the element routine is inlined into two kernels, whose FMA
    contraction may differ in the last bit of a coefficient)."""
    import bench
    import kiltro_b200 as K
    from kiltro_b200 import _lib
    from kiltro_b200.multigrid import GridMultigrid
    lib = _lib.load()
    mesh, vel, flags, mat = bench.make_problem(K, torch, nps)
    form = K.HeatAdvectionDiffusion(mesh, velocity=vel * peclet_scale) if peclet_scale > 0 else K.HeatDiffusion(mesh)
    bc = K.BoundaryCondition(); bc.SetDirichletCriteria(lambda: flags)
    fs = K.FEMSolver(verbose=False)
    Kc = K.Assemble(fs, form.function_spaces, form, mesh, mat, bc)[0]
    bc.GetDirichletBoundaryConditions(form, mesh, mat, fs)
    F = torch.zeros((mesh.nnodes, 1), dtype=torch.float64, device="cuda")
    Kb, Fb = bc.ApplyDirichletGetReducedMatrices(Kc, F, bc.applied_dirichlet_device)[:2]
    g = torch.Generator(device="cuda"); g.manual_seed(11)
    r = torch.randn(Kb.shape[0], dtype=torch.float64, device="cuda", generator=g)
    mg = GridMultigrid(mesh, form, mat, bc)
    try:
        lib.kb_mg_set_assemble_tiled(0)
        mg.setup(Kb)
        ref = mg.apply(r)
        lib.kb_mg_set_assemble_tiled(1)
        mg.setup(Kb)
        got = mg.apply(r)
    finally:
        lib.kb_mg_set_assemble_tiled(1)
    err = float(torch.linalg.vector_norm(got - ref) / torch.linalg.vector_norm(ref))
    print("tiled vs per-node coarse assembly, %d^2 nodes: relative difference of one cycle %.2e (%s)" % (
        nps, err, "bit-identical" if torch.equal(got, ref) else "rounding"))
    assert err < 1e-5, err


def test_multigrid_rejects_what_it_cannot_precondition():
    import kiltro_b200 as K
    from kiltro_b200.multigrid import GridMultigrid
    from kiltro_b200 import _lib
    m = K.Mesh(element_type="quad"); m.Rectangle((0.0, 0.0), (1.0, 1.0), 12, 9)
    nn = m.nnodes
    mat = K.FourierConduction(2, k=1.0, area=1.0, density=1.0, heat_capacity_v=1.0, q=0.0)
    flags = torch.full((nn, 1), float("nan"), dtype=torch.float64, device="cuda"); flags[0, 0] = 1.0
    bc = K.BoundaryCondition(); bc.SetDirichletCriteria(lambda: flags)
    vel = torch.ones((nn, 2), dtype=torch.float64, device="cuda")
    with pytest.raises(ValueError):
        GridMultigrid(m, K.Temperature(m, velocity=vel), mat, bc)          # scalar-broadcast advection (KB_MODE_REFERENCE)
    # a scrambled numbering is not a grid
    pts = K.to_numpy(m.device_points()); el = K.to_numpy(m.device_elements()).astype(np.int64)
    perm = np.random.RandomState(0).permutation(nn)
    inv = np.empty(nn, np.int64); inv[perm] = np.arange(nn)
    m2 = K.Mesh(element_type="quad"); m2.set_arrays(pts[perm], inv[el])
    with pytest.raises(_lib.KiltroB200Error):
        GridMultigrid(m2, K.HeatDiffusion(m2), mat, bc)


def test_multigrid_fullsize_S16():
    """S16 (the timed configuration): status 0 at rtol 1e-10 in a few dozen iterations, true residual recomputed from a fresh
    assembly, and agreement with the prescribed Jacobi-BiCGSTAB solution of the same problem to <= 1e-8."""
    import bench
    import kiltro_b200 as K
    from test_gpu_benchcase import _solve_device, _true_relative_residual
    nps = 4000
    mesh, vel, flags, mat = bench.make_problem(K, torch, nps)
    form = K.HeatAdvectionDiffusion(mesh, velocity=vel)
    out, solver, mg, fs, bc = _mg_solve(K, mesh, form, mat, flags, 1e-10)
    info = solver.info
    assert info["status"] == 0 and info["iterations"] <= 40, info
    T = out.sol_device[:, 0, 0].clone()
    rel = _true_relative_residual(K, fs, form, mesh, mat, bc, T)
    assert rel <= 1e-10 * 1.01, (rel, info)
    cache = {"mesh": mesh, "vel": vel, "flags": flags, "mat": mat}
    out_j, solver_j = _solve_device(K, bench, nps, 1e-12, cache=cache)[:2]
    diff = float(torch.linalg.vector_norm(T - out_j.sol_device[:, 0, 0]) / torch.linalg.vector_norm(T))
    assert diff <= SOL_TOL, (diff, info, solver_j.info)
    print("S16 multigrid: %d iterations (Jacobi-BiCGSTAB %d), true residual %.2e, ||x_mg - x_jacobi|| / ||x|| = %.2e" % (
        info["iterations"], solver_j.info["iterations"], rel, diff))
