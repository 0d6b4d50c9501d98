"""GPU parity tests: the CUDA path (through the C ABI, via the reference-shaped Python API) against the CPU oracle and
the reference-generated golden fixtures.  Integer/index outputs must be bit-exact; matrix values <= 1e-12 relative;
solutions <= 1e-8 relative L2 (BASELINE.json north_star)."""
import numpy as np
import pytest

from conftest import Golden, rel_l2, rel_max

pytestmark = pytest.mark.gpu

torch = pytest.importorskip("torch")

PAT = Golden("pattern")
ASM = Golden("asm")
STE = Golden("steady")

VAL_TOL = 1e-12
SOL_TOL = 1e-8


def _K():
    import kiltro_b200 as K
    return K


def make_mesh(points, elements):
    K = _K()
    m = K.Mesh(element_type="quad" if elements.shape[1] == 4 else "line")
    m.points = points
    m.elements = elements
    return m


def formulation_for(name, mesh, velocity):
    K = _K()
    if name.endswith("_pg"):
        return K.TemperaturePG(mesh, velocity=velocity)
    tail = name.split("_")[-1]
    if tail in ("cons", "consistent"):
        return K.HeatAdvectionDiffusion(mesh, velocity=velocity)
    return K.Temperature(mesh, velocity=velocity)


def material_for(vec, ndim):
    K = _K()
    k, rho, cv, area, q = [float(v) for v in vec]
    return K.FourierConduction(ndim, k=k, area=area, density=rho, heat_capacity_v=cv, q=q)


# ------------------------------------------------------------------------------------------------- K8 mesh
@pytest.mark.parametrize("args", [((0.0, 0.0), (1.0, 1.0), 2, 2), ((0, 0), (0.5, 0.01), 5, 3), ((-1.5, 0.25), (3.7, 9.1), 37, 23),
                                  ((0, 0), (1, 1), 1, 1), ((0.0, 0.0), (1.0, 1.0), 300, 7)])
def test_mesh_rectangle_bit_exact(args):
    from oracle import kiltro_oracle as O
    ll, ur, nx, ny = args
    m = _K().Mesh(element_type="quad"); m.Rectangle(lower_left_point=ll, upper_right_point=ur, nx=nx, ny=ny)
    pts, el = O.mesh_rectangle(ll, ur, nx, ny)
    assert m.nnodes == pts.shape[0] and m.nelem == el.shape[0] and m.ndim == 2
    assert m.elements.dtype == np.int64 and np.array_equal(m.elements, el)
    assert np.array_equal(m.points, pts)


@pytest.mark.parametrize("args", [(0.0, 0.5, 5), (-1.0, 2.0, 17), (0.0, 1.0, 1), (3.0, 3.5, 1000)])
def test_mesh_line_bit_exact(args):
    from oracle import kiltro_oracle as O
    m = _K().Mesh(element_type="line"); m.Line(*args)
    pts, el = O.mesh_line(*args)
    assert np.array_equal(m.points, pts) and np.array_equal(m.elements, el) and m.ndim == 1


# ------------------------------------------------------------------------------------------------- K2 pattern
@pytest.mark.parametrize("name", [c for c in PAT.cases() if "nvar2" not in c])
def test_pattern_golden(name):
    c = PAT.case(name)
    el = c["elements"]
    m = make_mesh(np.zeros((int(c["nnodes"]), 2 if el.shape[1] == 4 else 1)), el)
    ind, ptr, dl, dg = _K().ComputeSparsityPattern(m, 1)
    assert ind.dtype == np.int32 and ptr.dtype == np.int32 and dl.dtype == np.int32 and dg.dtype == np.int32
    assert np.array_equal(ptr, c["indptr"]) and np.array_equal(ind, c["indices"])
    assert np.array_equal(dl, c["data_local"]) and np.array_equal(dg, c["data_global"])


def test_pattern_nvar2_golden():
    c = PAT.case("rect2x2_nvar2")
    m = make_mesh(np.zeros((9, 2)), PAT.case("rect2x2")["elements"])
    ind, ptr, dl, dg = _K().ComputeSparsityPattern(m, 2)
    assert np.array_equal(ptr, c["indptr"]) and np.array_equal(ind, c["indices"])
    assert np.array_equal(dl, c["data_local"]) and np.array_equal(dg, c["data_global"])


@pytest.mark.parametrize("nx,ny", [(1, 1), (1, 5), (7, 1), (2, 2), (40, 31), (257, 129)])
def test_pattern_grid_shortcut_bit_exact(nx, ny):
    """kb_pattern_grid (pattern of a Mesh.Rectangle-numbered mesh written down from the shape) equals the sort-based
    phase bit for bit -- indptr, indices, diagonal positions -- and the lazily built scatter maps are the sort-based ones."""
    import importlib
    CSP = importlib.import_module("kiltro_b200.FiniteElements.ComputeSparsityPattern")    # (the package re-exports a function of that name)
    K = _K()
    m = K.Mesh(element_type="quad"); m.Rectangle((0.0, 0.0), (1.0, 2.0), nx, ny)
    fast = CSP.ComputeSparsityPatternDevice(m, 1)
    assert fast._maps_builder is not None and fast.grid == (nx + 1, ny + 1)          # the shortcut was taken, maps not built yet
    CSP.GRID_SHORTCUT = False
    try:
        slow = CSP.ComputeSparsityPatternDevice(m, 1)
    finally:
        CSP.GRID_SHORTCUT = True
    assert fast.nnz == slow.nnz
    for name in ("indptr", "indices", "diag_pos", "data_local_indices", "data_global_indices", "seg_ptr", "gather_src"):
        assert torch.equal(getattr(fast, name), getattr(slow, name)), name
    assert fast._maps_builder is None


@pytest.mark.parametrize("bad_id", [-1, 9, 1 << 20])
def test_pattern_rejects_out_of_range_connectivity(bad_id):
    """A node id outside [0, nnode) in a user-provided mesh returns KB_EINVAL instead of overflowing the sort key and
    writing indptr out of bounds."""
    from kiltro_b200 import _lib
    el = PAT.case("rect2x2")["elements"].copy()
    el[1, 2] = bad_id
    m = make_mesh(np.zeros((9, 2)), el)
    with pytest.raises(_lib.KiltroB200Error):
        _K().ComputeSparsityPattern(m, 1)
    # the library is still usable afterwards, and the scratch of the failed call was released (no error on the next one)
    m2 = make_mesh(np.zeros((9, 2)), PAT.case("rect2x2")["elements"])
    ind, ptr, _, _ = _K().ComputeSparsityPattern(m2, 1)
    assert np.array_equal(ptr, PAT.case("rect2x2")["indptr"])


@pytest.mark.parametrize("nx,ny,scramble", [(40, 31, False), (64, 64, True), (257, 129, False)])
def test_pattern_vs_oracle_c(nx, ny, scramble):
    from oracle import kiltro_oracle as O, native as ON
    pts, el = O.mesh_rectangle((0, 0), (1, 1), nx, ny)
    if scramble:
        rng = np.random.default_rng(3)
        el = el[rng.permutation(el.shape[0])]
        for e in range(0, el.shape[0], 3):
            el[e] = np.roll(el[e], e % 4)
    m = make_mesh(pts, el)
    got = _K().ComputeSparsityPattern(m, 1)
    want = ON.c_sparsity_pattern(el, pts.shape[0], 1)
    for g, w in zip(got, want):
        assert np.array_equal(g, w)
    assert got[0].shape[0] == (3 * (nx + 1) - 2) * (3 * (ny + 1) - 2)


@pytest.mark.parametrize("n,petrov,transient", [(2, False, False), (3, True, True), (257, True, False), (10001, False, True),
                                                (100000, True, True)])
def test_line_grid_assembly_matches_tile_kernels(n, petrov, transient):
    """csrc/assemble_line.cu (the grid form of Mesh.Line-numbered meshes: what assembly='auto' picks) equals the tile kernel
    and the two-pass seam to rounding (K, M and the source vector), and a scrambled numbering falls back to the tile kernel."""
    K = _K()
    m = K.Mesh(element_type="line"); m.Line(0.0, 2.0, n - 1)
    g = torch.Generator(device="cuda"); g.manual_seed(n)
    vel = 3.0 + torch.rand((m.nnodes, 1), dtype=torch.float64, device="cuda", generator=g)
    form = (K.TemperaturePG if petrov else K.Temperature)(m, velocity=vel)
    mat = K.FourierConduction(1, k=0.7, area=1.3, density=1.1, heat_capacity_v=0.9, q=2.0)
    bc = K.BoundaryCondition()
    out = {}
    for method in ("auto", "fused_generic", "two_pass"):
        fs = K.FEMSolver(analysis_type="transient" if transient else "steady", assembly=method, verbose=False)
        Kc, F, M = K.Assemble(fs, form.function_spaces, form, m, mat, bc)
        out[method] = (Kc.data.clone(), F.clone(), M.data.clone() if transient else None, fs.assembly_used)
    assert out["auto"][3] == "fused_line" and out["fused_generic"][3] == "fused"
    worst = [0.0]

    def close(a, b):                                 # rounding level: same sums in the same order, but the element routine is
        scale = float(b.abs().max())                 # inlined into different kernels and nvcc contracts its products into FMAs
        err = float((a - b).abs().max()) / (scale if scale > 0 else 1.0)             # differently (K itself is <= 1e-12 of the
        worst[0] = max(worst[0], err)                                                # reference: test_element_matrices_and_assembly_golden)
        return err <= 1e-13
    for other in ("fused_generic", "two_pass"):
        assert close(out["auto"][0], out[other][0]) and close(out["auto"][1], out[other][1])
        if transient:
            assert close(out["auto"][2], out[other][2])
    print("line grid form vs tile kernel, n=%d petrov=%s: max relative difference %.2e" % (n, petrov, worst[0]))
    if n >= 3:
        pts = K.to_numpy(m.device_points()); el = K.to_numpy(m.device_elements()).astype(np.int64)
        m2 = K.Mesh(element_type="line"); m2.set_arrays(pts, el[::-1].copy())            # same nodes, elements reversed
        fs = K.FEMSolver(verbose=False)
        K.Assemble(fs, form.function_spaces, (K.TemperaturePG if petrov else K.Temperature)(m2, velocity=vel), m2, mat, bc)
        assert fs.assembly_used != "fused_line"


def test_pattern_line_vs_oracle():
    from oracle import kiltro_oracle as O, native as ON
    pts, el = O.mesh_line(0, 1, 5000)
    got = _K().ComputeSparsityPattern(make_mesh(pts, el), 1)
    want = ON.c_sparsity_pattern(el, pts.shape[0], 1)
    for g, w in zip(got, want):
        assert np.array_equal(g, w)
    assert got[0].shape[0] == 3 * 5001 - 2


# ------------------------------------------------------------------------------------------------- K1 + K3
@pytest.mark.parametrize("name", ASM.cases())
def test_element_matrices_and_assembly_golden(name):
    K = _K()
    c = ASM.case(name)
    m = make_mesh(c["points"], c["elements"])
    form = formulation_for(name, m, c["velocity"])
    mat = material_for(c["material"], m.ndim)
    Ke, Fe, Me = form.GetAllElementalMatrices(m, mat, want_mass=True)
    assert rel_max(K.to_numpy(Ke), c["Ke"]) < VAL_TOL
    assert rel_max(K.to_numpy(Fe), c["Fe"]) < VAL_TOL
    assert rel_max(K.to_numpy(Me), c["Me"]) < VAL_TOL
    transient = "M_data" in c
    fs = K.FEMSolver(analysis_type="transient" if transient else "steady", verbose=False)
    Kc, Flux, Mm = K.Assemble(fs, form.function_spaces, form, m, mat, K.BoundaryCondition())
    assert np.array_equal(K.to_numpy(Kc.indices), c["K_indices"]) and np.array_equal(K.to_numpy(Kc.indptr), c["K_indptr"])
    assert rel_max(K.to_numpy(Kc.data), c["K_data"]) < VAL_TOL
    assert rel_max(K.to_numpy(Flux).ravel(), c["Flux"]) < VAL_TOL
    if transient:
        assert rel_max(K.to_numpy(Mm.data), c["M_data"]) < VAL_TOL
    # the reference's per-element seam
    I, J, V, f, Im, Jm, Vm = form.GetElementalMatrices(3, form.function_spaces[0], m, mat, fs)
    assert rel_max(V, c["Ke"][3]) < VAL_TOL and rel_max(f.ravel(), c["Fe"][3]) < VAL_TOL


@pytest.mark.parametrize("mode,pg", [("reference", False), ("reference", True), ("consistent", False), ("consistent", True)])
def test_assembly_vs_oracle_medium(mode, pg):
    from oracle import kiltro_oracle as O
    K = _K()
    rng = np.random.default_rng(11)
    nx, ny = 61, 47
    pts, el = O.mesh_rectangle((0, 0), (2.0, 1.0), nx, ny)
    hx, hy = 2.0 / nx, 1.0 / ny
    inner = (pts[:, 0] > 0) & (pts[:, 0] < 2.0) & (pts[:, 1] > 0) & (pts[:, 1] < 1.0)
    pts[inner] += 0.2 * np.array([hx, hy]) * rng.uniform(-1, 1, (inner.sum(), 2))
    vel = np.array([3.0, -1.0]) + rng.standard_normal(pts.shape)
    matv = (1.3, 0.9, 1.7, 1.1, 2.5)
    A = O.assemble(pts, el, vel, matv, mode, pg, True)
    m = make_mesh(pts, el)
    cls = {("reference", False): K.Temperature, ("reference", True): K.TemperaturePG,
           ("consistent", False): K.HeatAdvectionDiffusion, ("consistent", True): K.HeatAdvectionDiffusionPG}[(mode, pg)]
    form = cls(m, velocity=vel)
    fs = K.FEMSolver(analysis_type="transient", verbose=False)
    Kc, Flux, Mm = K.Assemble(fs, form.function_spaces, form, m, material_for(matv, 2), K.BoundaryCondition())
    assert np.array_equal(K.to_numpy(Kc.indices), A["indices"]) and np.array_equal(K.to_numpy(Kc.indptr), A["indptr"])
    assert rel_max(K.to_numpy(Kc.data), A["K"]) < VAL_TOL
    assert rel_max(K.to_numpy(Mm.data), A["M"]) < VAL_TOL
    assert rel_max(K.to_numpy(Flux).ravel(), A["Flux"]) < VAL_TOL
    # determinism: a second assembly is bitwise identical
    K2 = K.Assemble(fs, form.function_spaces, form, m, material_for(matv, 2), K.BoundaryCondition())[0]
    assert torch.equal(K2.data, Kc.data)


# ------------------------------------------------------------------------------------------------- K4 + K6 (steady)
@pytest.mark.parametrize("name", STE.cases())
def test_steady_golden(name):
    K = _K()
    c = STE.case(name)
    m = make_mesh(c["points"], c["elements"])
    form = formulation_for(name, m, c["velocity"])
    mat = material_for(c["material"], m.ndim)
    bc = K.BoundaryCondition()
    bc.SetDirichletCriteria(lambda: c["dirichlet_flags"].copy())
    if "neumann_flags" in c:
        bc.SetNeumannCriteria(lambda: c["neumann_flags"].copy())
    fs = K.FEMSolver(analysis_type="steady", verbose=False)
    solver = K.LinearSolver(tol=1e-13, check_every=10)
    sol = fs.Solve(formulation=form, mesh=m, material=mat, boundary_condition=bc, solver=solver)
    assert np.array_equal(bc.columns_in, c["columns_in"]) and np.array_equal(bc.columns_out, c["columns_out"])
    assert bc.columns_out.dtype == np.int64
    # seam: reduced system
    Kc = K.Assemble(fs, form.function_spaces, form, m, mat, bc)[0]
    Res = -bc.ComputeNeumannFluxes(m, mat)
    K_b, F_b, F = bc.ApplyDirichletGetReducedMatrices(Kc, Res, bc.applied_dirichlet)
    assert np.array_equal(K.to_numpy(K_b.indices), c["Kb_indices"]) and np.array_equal(K.to_numpy(K_b.indptr), c["Kb_indptr"])
    assert rel_max(K.to_numpy(K_b.data), c["Kb_data"]) < VAL_TOL
    assert rel_max(K.to_numpy(F_b), c["F_b"]) < 1e-11
    assert sol.sol.shape == c["sol"].shape
    assert solver.info["status"] == 0, solver.info
    assert rel_l2(sol.sol, c["sol"]) < SOL_TOL, solver.info


def test_spmv_vs_scipy():
    from oracle import kiltro_oracle as O
    K = _K()
    rng = np.random.default_rng(5)
    pts, el = O.mesh_rectangle((0, 0), (1, 1), 150, 97)
    vel = rng.standard_normal(pts.shape)
    m = make_mesh(pts, el)
    form = K.HeatAdvectionDiffusion(m, velocity=vel)
    fs = K.FEMSolver(verbose=False)
    Kc = K.Assemble(fs, form.function_spaces, form, m, material_for((1, 1, 1, 1, 0), 2), K.BoundaryCondition())[0]
    x = rng.standard_normal(pts.shape[0])
    y = K.to_numpy(Kc.dot(x))
    assert rel_max(y, Kc.tocsr() @ x) < 1e-13


@pytest.mark.parametrize("n,mode", [(127, "consistent"), (255, "diffusion"), (200, "reference_pg")])
def test_steady_vs_oracle_spsolve(n, mode):
    """Synthetic structured case of BASELINE config 4 at oracle-sized n: Dirichlet x=0 -> 100, x=1 -> 500."""
    from oracle import kiltro_oracle as O
    K = _K()
    pts, el = O.mesh_rectangle((0, 0), (1, 1), n, n)
    h = 1.0 / n
    rng = np.random.default_rng(0)
    U = 0.45 * 2.0 / h                                         # cell Peclet 0.45
    vel = np.zeros(pts.shape); vel[:, 0] = U; vel += 0.1 * U * rng.standard_normal(pts.shape)
    d = np.full((pts.shape[0], 1), np.nan); d[pts[:, 0] == 0.0] = 100.0; d[pts[:, 0] == 1.0] = 500.0
    matv = (1.0, 1.0, 1.0, 1.0, 0.0)
    m = make_mesh(pts, el)
    if mode == "consistent":
        form = K.HeatAdvectionDiffusion(m, velocity=vel); ref = O.steady_solve(pts, el, vel, matv, d, None, "consistent", False)
    elif mode == "diffusion":
        form = K.HeatDiffusion(m); ref = O.steady_solve(pts, el, None, matv, d, None, "reference", False)
    else:
        form = K.TemperaturePG(m, velocity=vel * 1e-3); ref = O.steady_solve(pts, el, vel * 1e-3, matv, d, None, "reference", True)
    bc = K.BoundaryCondition(); bc.SetDirichletCriteria(lambda: d.copy())
    solver = K.LinearSolver(tol=1e-13, maxiter=20000)
    sol = K.FEMSolver(verbose=False).Solve(formulation=form, mesh=m, material=material_for(matv, 2), boundary_condition=bc,
                                           solver=solver)
    assert solver.info["status"] == 0, solver.info
    assert rel_l2(sol.sol.ravel(), ref) < SOL_TOL, solver.info


def test_closed_form_galerkin_advection():
    """1D-like 2D problem: consistent Galerkin advection has the discrete solution T_i = T_0 + dT (r^i-1)/(r^n-1),
    r = (1+Pe)/(1-Pe)."""
    K = _K()
    nx, ny = 64, 8
    m = K.Mesh(element_type="quad"); m.Rectangle((0, 0), (1.0, 0.125), nx, ny)
    h = 1.0 / nx; Pe = 0.45
    vel = np.zeros((m.nnodes, 2)); vel[:, 0] = Pe * 2.0 / h
    pts = m.points
    d = np.full((m.nnodes, 1), np.nan); d[pts[:, 0] == 0.0] = 100.0; d[pts[:, 0] == 1.0] = 500.0
    bc = K.BoundaryCondition(); bc.SetDirichletCriteria(lambda: d)
    sol = K.FEMSolver(verbose=False).Solve(formulation=K.HeatAdvectionDiffusion(m, velocity=vel), mesh=m,
                                           material=K.FourierConduction(2, k=1.0, area=1.0, density=1.0, c_v=1.0),
                                           boundary_condition=bc, solver=K.LinearSolver(tol=1e-13))
    r = (1 + Pe) / (1 - Pe)
    exact = 100.0 + 400.0 * (r ** np.arange(nx + 1) - 1) / (r ** nx - 1)
    got = sol.sol.reshape(ny + 1, nx + 1)
    assert np.abs(got - exact[None, :]).max() / 500.0 < 1e-8


# ------------------------------------------------------------------------------------------------- fused assembly (K1+K3)
@pytest.mark.parametrize("nx,ny,transient,cls", [(70, 53, True, "HeatAdvectionDiffusion"), (45, 31, False, "TemperaturePG"),
                                                 (15, 15, True, "Temperature"), (16, 14, False, "HeatAdvectionDiffusionPG"),
                                                 (3, 2, True, "HeatAdvectionDiffusion"), (1, 1, True, "HeatAdvectionDiffusion"),
                                                 (40, 150, True, "HeatAdvectionDiffusion"), (95, 130, False, "Temperature")])
def test_fused_assembly_structured_tiles(nx, ny, transient, cls):
    """Device-generated Rectangle mesh (2-D node-patch tiles), nodes perturbed in place on device: fused == two-pass
    bitwise, and both match the oracle."""
    from oracle import kiltro_oracle as O
    from kiltro_b200.FiniteElements.Assembly import GRID_IN_AUTO
    K = _K()
    m = K.Mesh(element_type="quad"); m.Rectangle((0.0, 0.0), (2.0, 1.0), nx, ny)
    g = torch.Generator(device="cuda"); g.manual_seed(7)
    P = m.device_points()
    hx, hy = 2.0 / nx, 1.0 / ny
    inner = (P[:, 0] > 0) & (P[:, 0] < 2.0) & (P[:, 1] > 0) & (P[:, 1] < 1.0)
    noise = (torch.rand(P.shape, dtype=torch.float64, device="cuda", generator=g) * 2 - 1) * torch.tensor([hx, hy], dtype=torch.float64, device="cuda") * 0.2
    P += noise * inner[:, None]
    vel = torch.randn(P.shape, dtype=torch.float64, device="cuda", generator=g) * 3.0 + 2.0
    matv = (1.3, 0.9, 1.7, 1.1, 2.5)
    form = getattr(K, cls)(m, velocity=vel)
    at = "transient" if transient else "steady"
    res = {}
    used = {"fused": "fused_template", "fused_generic": "fused", "two_pass": "two_pass", "fused_grid": "fused_grid",
            "auto": "fused_grid" if GRID_IN_AUTO else "fused_template"}
    for method in ("fused", "fused_generic", "two_pass", "fused_grid", "auto"):
        fs = K.FEMSolver(analysis_type=at, verbose=False, assembly=method)
        Kc, Flux, Mm = K.Assemble(fs, form.function_spaces, form, m, material_for(matv, 2), K.BoundaryCondition())
        assert fs.assembly_used == used[method]
        res[method] = (Kc, Flux, Mm)
        if method == "fused":
            info = fs.sparsity.tile_plan.info
            assert info["structured"] and fs.sparsity.tile_plan.templated and 1 <= info["templates"] <= 9
    # template kernel == generic fused kernel == grid kernel, bit for bit (same element routine, same summation order)
    for other in ("fused_generic", "fused_grid", "auto"):
        assert torch.equal(res["fused"][0].data, res[other][0].data), other
        assert torch.equal(res["fused"][1], res[other][1]), other
        if transient:
            assert torch.equal(res["fused"][2].data, res[other][2].data), other
    # same device function, same accumulation order; the two kernels may contract FMAs differently in the tanh/PG branch
    assert rel_max(K.to_numpy(res["fused"][0].data), K.to_numpy(res["two_pass"][0].data)) < 1e-14
    assert torch.equal(res["fused"][1], res["two_pass"][1])
    if transient:
        assert torch.equal(res["fused"][2].data, res["two_pass"][2].data)
    mode = "consistent" if cls.startswith("Heat") else "reference"
    A = O.assemble(K.to_numpy(P), m.elements, K.to_numpy(vel), matv, mode, cls.endswith("PG"), transient)
    assert np.array_equal(K.to_numpy(res["fused"][0].indices), A["indices"])
    assert rel_max(K.to_numpy(res["fused"][0].data), A["K"]) < VAL_TOL
    assert rel_max(K.to_numpy(res["fused"][1]).ravel(), A["Flux"]) < VAL_TOL
    if transient:
        assert rel_max(K.to_numpy(res["fused"][2].data), A["M"]) < VAL_TOL


@pytest.mark.parametrize("cls,transient", [("HeatAdvectionDiffusion", True), ("TemperaturePG", False)])
def test_fused_assembly_nonconvex_fixup(cls, transient):
    """Non-convex quads (Jacobian determinant changes sign inside the element) are declined by the sum-factorised
    element routine; their tiles are redone by the fix-up launch with the general routine.  All three paths must agree
    with the oracle, which follows the reference (dV = |det J|, VariationalPrinciple.py:145)."""
    from oracle import kiltro_oracle as O
    K = _K()
    nx, ny = 47, 33
    m = K.Mesh(element_type="quad"); m.Rectangle((0.0, 0.0), (1.0, 1.0), nx, ny)
    P = m.device_points()
    hx, hy = 1.0 / nx, 1.0 / ny
    # pull a few interior nodes most of the way to their lower-left neighbour: the element above-right turns non-convex
    for (i, j) in [(5, 7), (16, 16), (17, 16), (30, 2), (44, 31), (15, 15)]:
        n = j * (nx + 1) + i
        P[n, 0] -= 0.8 * hx; P[n, 1] -= 0.8 * hy
    g = torch.Generator(device="cuda"); g.manual_seed(3)
    vel = torch.randn(P.shape, dtype=torch.float64, device="cuda", generator=g) * 2.0 + 1.0
    matv = (0.7, 1.9, 0.6, 1.3, 0.0)
    at = "transient" if transient else "steady"
    form = getattr(K, cls)(m, velocity=vel)
    mode = "consistent" if cls.startswith("Heat") else "reference"
    # (connectivity taken from the device copy: reading m.elements would make the host arrays authoritative)
    A = O.assemble(K.to_numpy(P), K.to_numpy(m.device_elements()).astype(np.int64), K.to_numpy(vel), matv, mode,
                   cls.endswith("PG"), transient)
    res = {}
    for method in ("fused", "fused_generic", "two_pass", "fused_grid"):
        fs = K.FEMSolver(analysis_type=at, verbose=False, assembly=method)
        Kc, Flux, Mm = K.Assemble(fs, form.function_spaces, form, m, material_for(matv, 2), K.BoundaryCondition())
        res[method] = Kc
        assert rel_max(K.to_numpy(Kc.data), A["K"]) < VAL_TOL, method
        if transient:
            assert rel_max(K.to_numpy(Mm.data), A["M"]) < VAL_TOL, method
        if method == "fused_grid":           # the grid kernel redoes the whole mesh with the general routine: flag raised
            assert fs.assembly_used == "fused_grid" and int(fs.sparsity.grid_scratch.view(torch.int32)[4]) == 1
    assert torch.equal(res["fused"].data, res["fused_generic"].data)
    # the tile flags say the fix-up launch really ran
    fs = K.FEMSolver(analysis_type=at, verbose=False, assembly="fused")
    K.Assemble(fs, form.function_spaces, form, m, material_for(matv, 2), K.BoundaryCondition())
    flags = K.to_numpy(fs.sparsity.tile_plan.scratch)
    assert flags[:fs.sparsity.tile_plan.ntiles].sum() >= 4


@pytest.mark.parametrize("name", ["quad_cons", "quad_scrambled_pg", "line_pg", "line_cons"])
def test_fused_assembly_row_chunk_tiles_golden(name):
    """Host-provided connectivity (row-chunk tiles; scrambled meshes may fall back to the two-pass seam)."""
    K = _K()
    c = ASM.case(name)
    m = make_mesh(c["points"], c["elements"])
    form = formulation_for(name, m, c["velocity"])
    mat = material_for(c["material"], m.ndim)
    fs = K.FEMSolver(analysis_type="transient", verbose=False, assembly="auto")
    Kc, Flux, Mm = K.Assemble(fs, form.function_spaces, form, m, mat, K.BoundaryCondition())
    assert rel_max(K.to_numpy(Kc.data), c["K_data"]) < VAL_TOL
    assert rel_max(K.to_numpy(Mm.data), c["M_data"]) < VAL_TOL
    assert rel_max(K.to_numpy(Flux).ravel(), c["Flux"]) < VAL_TOL
    fs2 = K.FEMSolver(analysis_type="transient", verbose=False, assembly="two_pass")
    K2 = K.Assemble(fs2, form.function_spaces, form, m, mat, K.BoundaryCondition())[0]
    assert rel_max(K.to_numpy(K2.data), K.to_numpy(Kc.data)) < 1e-14


def test_grid_assembly_on_host_provided_arrays():
    """User-provided (host) arrays: the Mesh.Rectangle numbering is recognised from the connectivity itself and checked
    on device; a mesh with any other numbering is refused by assembly='fused_grid' (and served by the tile kernels)."""
    from oracle import kiltro_oracle as O
    K = _K()
    pts, el = O.mesh_rectangle((0.0, 0.0), (1.5, 1.0), 37, 29)
    rng = np.random.default_rng(5)
    vel = rng.standard_normal(pts.shape) * 2.0 + 1.0
    matv = (1.1, 0.8, 1.3, 0.9, 1.7)
    m = make_mesh(pts, el)
    form = K.HeatAdvectionDiffusion(m, velocity=vel)
    fs = K.FEMSolver(analysis_type="transient", verbose=False, assembly="fused_grid")
    Kc, Flux, Mm = K.Assemble(fs, form.function_spaces, form, m, material_for(matv, 2), K.BoundaryCondition())
    assert fs.assembly_used == "fused_grid" and fs.sparsity.grid == (38, 30)
    A = O.assemble(pts, el, vel, matv, "consistent", False, True)
    assert np.array_equal(K.to_numpy(Kc.indices), A["indices"])
    assert rel_max(K.to_numpy(Kc.data), A["K"]) < VAL_TOL and rel_max(K.to_numpy(Mm.data), A["M"]) < VAL_TOL
    assert rel_max(K.to_numpy(Flux).ravel(), A["Flux"]) < VAL_TOL
    el2 = el.copy(); el2[[3, 11]] = el2[[11, 3]]                       # two elements swapped: not the grid numbering
    m2 = make_mesh(pts, el2)
    form2 = K.HeatAdvectionDiffusion(m2, velocity=vel)
    from kiltro_b200._lib import KiltroB200Error
    with pytest.raises(KiltroB200Error):
        K.Assemble(K.FEMSolver(verbose=False, assembly="fused_grid"), form2.function_spaces, form2, m2, material_for(matv, 2),
                   K.BoundaryCondition())
    fs3 = K.FEMSolver(verbose=False, assembly="auto")
    K3 = K.Assemble(fs3, form2.function_spaces, form2, m2, material_for(matv, 2), K.BoundaryCondition())[0]
    assert fs3.assembly_used != "fused_grid" and rel_max(K.to_numpy(K3.data), A["K"]) < VAL_TOL


def test_fused_assembly_line_mesh_large():
    from oracle import kiltro_oracle as O
    K = _K()
    m = K.Mesh(element_type="line"); m.Line(0.0, 3.0, 5000)
    vel = np.full((m.nnodes, 1), 7.0)
    matv = (0.8, 1.1, 1.4, 2.0, -2.0)
    form = K.TemperaturePG(m, velocity=vel)
    fs = K.FEMSolver(analysis_type="transient", verbose=False, assembly="fused")
    Kc, Flux, Mm = K.Assemble(fs, form.function_spaces, form, m, material_for(matv, 1), K.BoundaryCondition())
    A = O.assemble(m.points, m.elements, vel, matv, "reference", True, True)
    assert rel_max(K.to_numpy(Kc.data), A["K"]) < VAL_TOL and rel_max(K.to_numpy(Mm.data), A["M"]) < VAL_TOL
    assert rel_max(K.to_numpy(Flux).ravel(), A["Flux"]) < VAL_TOL


# ------------------------------------------------------------------------------------------------- K7 transient
TRA = Golden("transient")


def _transient_setup(c, name, theta=0.0, snaps=None, nincr=None):
    K = _K()
    m = make_mesh(c["points"], c["elements"])
    vel = c["velocity"] if np.any(c["velocity"]) else None
    form = K.HeatAdvectionDiffusion(m, velocity=vel) if name.endswith("consistent") else K.Temperature(m, velocity=vel)
    mat = material_for(c["material"], m.ndim)
    bc = K.BoundaryCondition()
    bc.SetInitialConditions(lambda: c["initial"].copy())
    bc.SetDirichletCriteria(lambda: c["dirichlet_flags"].copy())
    if "neumann_flags" in c:
        bc.SetNeumannCriteria(lambda: c["neumann_flags"].copy())
    fs = K.FEMSolver(analysis_type="transient", number_of_increments=int(c["nincr"]) if nincr is None else nincr,
                     theta=theta, time_step=float(c["dt"]), verbose=False)
    if snaps is not None:
        fs.snapshot_increments = snaps
    return K, m, form, mat, bc, fs


@pytest.mark.parametrize("name", TRA.cases())
def test_transient_golden(name):
    """theta = 0: forward Euler with consistent mass and (M^-1)[in,in] -- the reference's intent (FEMSolver.py:253-297),
    pinned by fixtures built on the reference's own Assemble and by its shipped figures."""
    c = TRA.case(name)
    K, m, form, mat, bc, fs = _transient_setup(c, name, snaps="all")
    T = fs.Solve(formulation=form, mesh=m, material=mat, boundary_condition=bc, solver=K.LinearSolver(tol=1e-13))
    assert T.shape == c["T_all"].shape
    assert fs.solver_info["status"] == 0, fs.solver_info
    assert rel_l2(T, c["T_all"]) < SOL_TOL, fs.solver_info
    K, m, form, mat, bc, fs = _transient_setup(c, name)
    S = fs.Solve(formulation=form, mesh=m, material=mat, boundary_condition=bc, solver=K.LinearSolver(tol=1e-13))
    assert S.shape == c["snapshots"].shape and rel_l2(S, c["snapshots"]) < SOL_TOL


def test_transient_time_step_rule():
    c = TRA.case("ex2d_dt010")
    K, m, form, mat, bc, fs = _transient_setup(c, "ex2d_dt010")
    fs.time_step = None                                   # FEMSolver.py:138-144 with the upstream factor 0.1
    fs.Solve(formulation=form, mesh=m, material=mat, boundary_condition=bc, solver=K.LinearSolver(tol=1e-13))
    assert abs(fs.time_increment - float(c["dt"])) < 1e-12 * float(c["dt"])


@pytest.mark.parametrize("theta", [0.5, 1.0])
def test_transient_theta_method(theta):
    """theta > 0 (new capability): checked against a dense NumPy statement of the same scheme built on the oracle's
    assembled matrices."""
    from oracle import kiltro_oracle as O
    from scipy.sparse import csr_matrix
    c = TRA.case("distorted_consistent")
    nincr, dt = 25, float(c["dt"]) * 6.0                  # beyond the explicit stability limit on purpose
    K, m, form, mat, bc, fs = _transient_setup(c, "distorted_consistent", theta=theta, snaps="all", nincr=nincr)
    fs.time_step = dt
    T = fs.Solve(formulation=form, mesh=m, material=mat, boundary_condition=bc, solver=K.LinearSolver(tol=1e-13))
    A = O.assemble(c["points"], c["elements"], c["velocity"], tuple(c["material"]), "consistent", False, True)
    nn = c["points"].shape[0]
    Kd = csr_matrix((A["K"], A["indices"], A["indptr"]), shape=(nn, nn)).toarray()
    Md = csr_matrix((A["M"], A["indices"], A["indptr"]), shape=(nn, nn)).toarray()
    cin, cout, td = O.dirichlet_split(c["dirichlet_flags"])
    Fn = O.neumann_fluxes(c["neumann_flags"], nn)
    free = np.zeros(nn, bool); free[cin] = True
    Kff = Kd * free[:, None] * free[None, :]
    F = np.zeros(nn); F[cin] = Kd[np.ix_(cin, cout)] @ td; F -= Fn; F[~free] = 0.0
    Tn = c["initial"].ravel().copy(); Tn[cout] += td
    ref = [Tn.copy()]
    Amat = Md + theta * dt * Kff
    for _ in range(nincr - 1):
        y = np.linalg.solve(Amat, Kff @ Tn + F)
        Tn = np.where(free, Tn - dt * y, 0.0); Tn[cout] = td
        ref.append(Tn.copy())
    ref = np.array(ref).T
    assert np.isfinite(T).all() and rel_l2(T, ref) < SOL_TOL, fs.solver_info


# ------------------------------------------------------------------------------------------------- K9 partitioned solves
def _dist_problem(nx, ny, world, transient, sym=False):
    """Global problem + its row-block partition emulated on one device (LocalComm)."""
    K = _K()
    from kiltro_b200 import distributed as KD
    rect = ((0.0, 0.0), (1.0, 0.75), nx, ny)
    U = 0.45 * 2.0 * nx

    def velocity_fn(pts):
        v = torch.empty_like(pts)
        v[:, 0] = U * (1.0 + 0.1 * torch.sin(7.0 * pts[:, 1]))
        v[:, 1] = 0.1 * U * torch.cos(5.0 * pts[:, 0])
        return v

    def dirichlet_fn(pts):
        f = torch.full((pts.shape[0], 1), float("nan"), dtype=torch.float64, device=pts.device)
        f[pts[:, 0] == 0.0, 0] = 100.0
        f[pts[:, 0] == 1.0, 0] = 500.0
        return f

    def neumann_fn(pts):
        f = torch.full((pts.shape[0], 1), float("nan"), dtype=torch.float64, device=pts.device)
        f[(pts[:, 1] == 0.75) & (pts[:, 0] > 0.2) & (pts[:, 0] < 0.8), 0] = 3.0
        return f

    def initial_fn(pts):
        return (200.0 + 50.0 * torch.sin(3.0 * pts[:, 0] + pts[:, 1])).reshape(-1, 1)

    mat = K.FourierConduction(2, k=1.0, area=1.0, density=2.0, heat_capacity_v=1.5, q=0.0)
    cls = K.HeatDiffusion if sym else K.HeatAdvectionDiffusion
    vfn = None if sym else velocity_fn
    parts = [KD.StripPartition(nx + 1, ny + 1, world, r) for r in range(world)]
    systems = [KD.build_local_system(p, rect, mat, cls, vfn, dirichlet_fn, neumann_fn, transient, initial_fn) for p in parts]
    # the same problem on one GPU through the ordinary API
    m = K.Mesh(element_type="quad"); m.Rectangle(rect[0], rect[1], nx, ny)
    P = m.device_points()
    form = cls(m) if sym else cls(m, velocity=velocity_fn(P))
    bc = K.BoundaryCondition()
    bc.SetDirichletCriteria(lambda: dirichlet_fn(P)); bc.SetNeumannCriteria(lambda: neumann_fn(P))
    bc.SetInitialConditions(lambda: initial_fn(P))
    return K, KD, systems, KD.LocalComm(world), (m, form, mat, bc)


@pytest.mark.parametrize("world,sym,nx", [(2, False, 44), (3, False, 44), (4, True, 44), (2, False, 63), (3, True, 95)])
def test_partitioned_steady_matches_single_gpu(world, sym, nx):
    """nx = 63 / 95: 64 / 96 nodes per mesh row -> 16-byte aligned row views -> the row-pattern form of the SpMV is used."""
    K, KD, systems, comm, (m, form, mat, bc) = _dist_problem(nx, 37, world, False, sym)
    assert all((s.row_pat is not None) == ((nx + 1) % 16 == 0) for s in systems)
    for s, p in zip(systems, [s.part for s in systems]):      # strip coordinates are bit-identical to the global mesh
        assert torch.equal(s.mesh.device_points(), m.device_points()[p.j_first * p.nxn:p.j_first * p.nxn + p.n_loc])
    sols, info = KD.solve_steady(systems, comm, rtol=1e-13, check_every=10)
    assert info["status"] == 0 and info["method"] == ("cg" if sym else "bicgstab"), info
    solver = K.LinearSolver(tol=1e-13)
    ref = K.FEMSolver(verbose=False).Solve(formulation=form, mesh=m, material=mat, boundary_condition=bc, solver=solver)
    got = torch.cat(sols).cpu().numpy()
    assert rel_l2(got, ref.sol.ravel()) < SOL_TOL, (info, solver.info)


@pytest.mark.parametrize("sym", [False, True])
def test_partitioned_steady_restart_cycles(sym):
    """Restart path of the partitioned driver: cycles capped at 40 iterations (restarted BiCGSTAB / CG from the true
    residual) must reach the same solution as the single-GPU solve."""
    K, KD, systems, comm, (m, form, mat, bc) = _dist_problem(44, 37, 3, False, sym)
    sols, info = KD.solve_steady(systems, comm, rtol=1e-13, check_every=10, cycle_iterations=40, max_restarts=400)
    assert info["status"] == 0 and info["restarts"] >= 2, info
    assert info["true_relative_residual"] <= 1e-12, info
    ref = K.FEMSolver(verbose=False).Solve(formulation=form, mesh=m, material=mat, boundary_condition=bc,
                                           solver=K.LinearSolver(tol=1e-13))
    assert rel_l2(torch.cat(sols).cpu().numpy(), ref.sol.ravel()) < SOL_TOL, info


@pytest.mark.parametrize("sym", [False, True])
def test_partitioned_steady_recovers_from_breakdown(sym):
    """A non-finite dot product in the middle of a cycle (injected once into the reduction) wrecks the iterate; the
    loop must hand back the best checked iterate and the driver must restart from it and still converge."""
    K, KD, systems, comm, (m, form, mat, bc) = _dist_problem(44, 37, 2, False, sym)

    class FaultyComm(KD.LocalComm):
        calls = 0

        def allreduce(self, outs):
            KD.LocalComm.allreduce(self, outs)
            FaultyComm.calls += 1
            if FaultyComm.calls == 36:                      # inside the second batch of 10 iterations
                for o in outs:
                    o.fill_(float("nan"))

    sols, info = KD.solve_steady(systems, FaultyComm(2), rtol=1e-13, check_every=10)
    assert FaultyComm.calls > 36
    assert info["status"] == 0 and info["restarts"] >= 1, info
    ref = K.FEMSolver(verbose=False).Solve(formulation=form, mesh=m, material=mat, boundary_condition=bc,
                                           solver=K.LinearSolver(tol=1e-13))
    assert rel_l2(torch.cat(sols).cpu().numpy(), ref.sol.ravel()) < SOL_TOL, info


@pytest.mark.parametrize("theta", [0.0, 0.5])
def test_partitioned_transient_matches_single_gpu(theta):
    K, KD, systems, comm, (m, form, mat, bc) = _dist_problem(30, 23, 3, True)
    fs = K.FEMSolver(analysis_type="transient", number_of_increments=13, theta=theta, verbose=False)
    fs.snapshot_increments = [12]
    ref = fs.Solve(formulation=form, mesh=m, material=mat, boundary_condition=bc, solver=K.LinearSolver(tol=1e-13))
    T, info = KD.run_transient(systems, comm, fs.time_increment, 12, theta=theta, rtol=1e-13)
    got = torch.cat(T).cpu().numpy()
    assert rel_l2(got, ref[:, 0]) < SOL_TOL, (info, fs.solver_info)


# ---- the native peer-memory solvers (csrc/krylov_peer.cuh): all ranks of the partition in one process on one device,
# one host thread + stream per rank, kernels exchanging halo rows / dot products through (here: same-device) peer memory
@pytest.mark.parametrize("world,sym,nx", [(2, False, 43), (3, False, 63), (4, True, 43), (3, True, 95), (4, False, 63), (8, False, 63)])
def test_peer_partitioned_steady_matches_single_gpu(world, sym, nx):
    """kb_peer_cg_solve / kb_peer_bicgstab_solve vs the single-GPU solve of the same global problem (<= 1e-8), and vs the
    host-driven partitioned loop (iteration counts within a few: same algorithm, different summation order)."""
    K, KD, systems, comm, (m, form, mat, bc) = _dist_problem(nx, 53, world, False, sym)
    sols, info = KD.solve_steady(systems, comm, rtol=1e-13, check_every=10, exchange="peer", maxiter=20000)
    assert info["exchange"] == "peer" and info["method"] == ("cg" if sym else "bicgstab"), info
    assert info["status"] == 0, info
    assert info["relative_residual"] <= 1e-12, info
    solver = K.LinearSolver(tol=1e-13)
    ref = K.FEMSolver(verbose=False).Solve(formulation=form, mesh=m, material=mat, boundary_condition=bc, solver=solver)
    got = torch.cat(sols).cpu().numpy()
    assert rel_l2(got, ref.sol.ravel()) < SOL_TOL, (info, solver.info)
    sols2, info2 = KD.solve_steady(systems, comm, rtol=1e-13, check_every=10, exchange="nccl")
    assert rel_l2(got, torch.cat(sols2).cpu().numpy()) < SOL_TOL
    assert abs(info["iterations"] - info2["iterations"]) <= max(20, info2["iterations"] // 5), (info, info2)
    # a second solve on the same regions (event / halo counters run on)
    sols3, info3 = KD.solve_steady(systems, comm, rtol=1e-13, check_every=10, exchange="peer")
    assert info3["status"] == 0 and rel_l2(torch.cat(sols3).cpu().numpy(), got) < 1e-10, info3
    KD.close_peer_regions(systems)


def test_peer_partitioned_solution_is_bitwise_reproducible():
    """Every rank adds the partial sums in rank order: two runs give the same bits and the same iteration count."""
    K, KD, systems, comm, _ = _dist_problem(63, 41, 3, False, False)
    a, ia = KD.solve_steady(systems, comm, rtol=1e-12, check_every=7, exchange="peer")
    b, ib = KD.solve_steady(systems, comm, rtol=1e-12, check_every=7, exchange="peer")
    assert ia["iterations"] == ib["iterations"] and ia["status"] == 0
    assert all(torch.equal(x, y) for x, y in zip(a, b))
    KD.close_peer_regions(systems)


@pytest.mark.parametrize("theta,sym", [(0.0, False), (0.5, False), (0.5, True)])
def test_peer_partitioned_transient_matches_single_gpu(theta, sym):
    """kb_peer_transient_run (time loop + inner CG / BiCGSTAB through peer memory) vs the single-GPU kb_transient_run."""
    K, KD, systems, comm, (m, form, mat, bc) = _dist_problem(31, 29, 3, True, sym)
    fs = K.FEMSolver(analysis_type="transient", number_of_increments=13, theta=theta, verbose=False)
    fs.snapshot_increments = [12]
    ref = fs.Solve(formulation=form, mesh=m, material=mat, boundary_condition=bc, solver=K.LinearSolver(tol=1e-13))
    T, info = KD.run_transient(systems, comm, fs.time_increment, 12, theta=theta, rtol=1e-13, exchange="peer")
    assert info["status"] == 0 and info["steps_done"] == 12 and info["exchange"] == "peer", info
    assert info["method"] == ("cg" if (theta == 0.0 or sym) else "bicgstab"), info
    got = torch.cat(T).cpu().numpy()
    assert rel_l2(got, ref[:, 0]) < SOL_TOL, (info, fs.solver_info)
    KD.close_peer_regions(systems)


def test_partitioned_api_single_rank_equals_ordinary_solve():
    """FEMSolver(partition="rows") on a one-rank partition (the whole mesh, no halo) goes through the partitioned drivers
    -- identity rows for prescribed DOFs, native peer loop with P = 1 -- and must reproduce the ordinary Solve(); the
    transient branch likewise (last increment)."""
    K = _K()
    nx, ny = 63, 40
    part = K.StripPartition(nx + 1, ny + 1, 1, 0)
    m = K.Mesh(element_type="quad"); m.Rectangle((0.0, 0.0), (1.0, 0.6), nx, ny, partition=part)
    m0 = K.Mesh(element_type="quad"); m0.Rectangle((0.0, 0.0), (1.0, 0.6), nx, ny)
    assert torch.equal(m.device_points(), m0.device_points()) and torch.equal(m.device_elements(), m0.device_elements())
    assert torch.equal(m.global_node_ids(), torch.arange(m0.nnodes, device="cuda"))
    P = m.device_points()
    vel = torch.stack([40.0 * (1.0 + 0.2 * torch.sin(9.0 * P[:, 1])), 5.0 * torch.cos(4.0 * P[:, 0])], dim=1).contiguous()
    flags = torch.full((m.nnodes, 1), float("nan"), dtype=torch.float64, device="cuda")
    flags[P[:, 0] == 0.0, 0] = 100.0; flags[P[:, 0] == 1.0, 0] = 500.0
    mat = K.FourierConduction(2, k=1.0, area=1.0, density=2.0, heat_capacity_v=1.5, q=0.0)

    def solve(mesh, **kw):
        bc = K.BoundaryCondition(); bc.SetDirichletCriteria(lambda: flags)
        fs = K.FEMSolver(verbose=False, **kw)
        out = fs.Solve(formulation=K.HeatAdvectionDiffusion(mesh, velocity=vel), mesh=mesh, material=mat, boundary_condition=bc,
                       solver=K.LinearSolver(tol=1e-13))
        return out, fs

    ref, _ = solve(m0)
    for ex in ("peer", "nccl"):
        got, fs = solve(m, partition="rows", exchange=ex)
        assert fs.solver_info["status"] == 0 and got.partition is part and got.converged, fs.solver_info
        assert got.sol_device.shape == (m.nnodes, 1, 1)
        assert rel_l2(got.sol.ravel(), ref.sol.ravel()) < SOL_TOL, (ex, fs.solver_info)
    with pytest.raises(ValueError):
        solve(m0, partition="rows")                            # a mesh without a partition
    # transient
    T0 = (200.0 + 30.0 * torch.sin(3.0 * P[:, 0] + P[:, 1])).reshape(-1, 1)

    def tsolve(mesh, **kw):
        bc = K.BoundaryCondition(); bc.SetDirichletCriteria(lambda: flags); bc.SetInitialConditions(lambda: T0)
        fs = K.FEMSolver(analysis_type="transient", number_of_increments=9, theta=0.5, time_step=2.0e-5, verbose=False, **kw)
        fs.snapshot_increments = [8]
        return fs.Solve(formulation=K.HeatAdvectionDiffusion(mesh, velocity=vel), mesh=mesh, material=mat, boundary_condition=bc,
                        solver=K.LinearSolver(tol=1e-13)), fs

    tref, _ = tsolve(m0)
    tgot, fs = tsolve(m, partition="rows")
    assert fs.solver_info["status"] == 0 and fs.solver_info["steps_done"] == 8, fs.solver_info
    assert rel_l2(tgot.sol.ravel(), np.asarray(tref)[:, 0]) < SOL_TOL, fs.solver_info


# ------------------------------------------------------------------------------------------------- "next" row n1
def test_load_increments_and_heat_source():
    """cumulative load increments (the reference's loop crashes for L > 1, FEMSolver.py:167-228)
    and the opt-in heat-source term (the reference assembles Flux and drops it)."""
    from oracle import kiltro_oracle as O
    K = _K()
    c = STE.case("mixed_consistent")
    m = make_mesh(c["points"], c["elements"])
    matv = list(c["material"]); matv[4] = 7.5                      # q != 0
    mat = material_for(matv, 2)

    def solve(**kw):
        form = K.HeatAdvectionDiffusion(m, velocity=c["velocity"])
        bc = K.BoundaryCondition()
        bc.SetDirichletCriteria(lambda: c["dirichlet_flags"].copy()); bc.SetNeumannCriteria(lambda: c["neumann_flags"].copy())
        fs = K.FEMSolver(verbose=False, **kw)
        return fs.Solve(formulation=form, mesh=m, material=mat, boundary_condition=bc, solver=K.LinearSolver(tol=1e-13)).sol

    one = solve()
    assert rel_l2(one, c["sol"]) < SOL_TOL                         # q does not reach the RHS by default (F6)
    three = solve(number_of_load_increments=3)
    assert three.shape == (m.nnodes, 1, 3)
    for i in range(3):
        assert rel_l2(three[:, :, i], one[:, :, 0] * (i + 1) / 3.0) < SOL_TOL
    # with the source: RHS = -F_neumann + Flux
    nn = m.nnodes
    A = O.assemble(c["points"], c["elements"], c["velocity"], tuple(matv), "consistent", False, False)
    cin, cout, td = O.dirichlet_split(c["dirichlet_flags"])
    res = -O.neumann_fluxes(c["neumann_flags"], nn) + A["Flux"]
    kb = O.apply_dirichlet_reduce(A["indices"], A["indptr"], A["K"], res, cin, cout, td)
    ref = O.total_component_sol(O.linear_solve(kb[0], kb[1], kb[2], kb[3]), cin, cout, td, nn)
    src = solve(include_heat_source=True)
    assert rel_l2(src.ravel(), ref) < SOL_TOL and rel_l2(src.ravel(), one.ravel()) > 1e-4


# ------------------------------------------------------------------------------------------------- example scripts (configs 1-3)
@pytest.mark.parametrize("script", ["steady_1d.py", "steady_2d.py", "transient_1d.py", "transient_2d.py"])
def test_example_scripts_run(script):
    import os, subprocess, sys
    from conftest import ROOT
    out = subprocess.run([sys.executable, os.path.join(ROOT, "examples", script)], capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stderr[-2000:]
    if script == "steady_1d.py":
        assert "23.7" in out.stdout and "112.8" in out.stdout            # steady_1d_addi.pdf
    if script == "steady_2d.py":
        assert "105.18" in out.stdout and "249.7" in out.stdout          # steady_2d_addi.pdf
    if script == "transient_2d.py":
        assert "time increment 0.8" in out.stdout and "199.8" in out.stdout
    if script == "transient_1d.py":                                      # tests/golden/transient.npz ex1d (reference-built)
        assert "time increment 0.8" in out.stdout and "189.527" in out.stdout and "118.625" in out.stdout


# ------------------------------------------------------------------------------------------------- edge cases
def test_all_dofs_prescribed_gives_empty_system():
    """Empty reduced system: LinearSolver returns a copy of b with a warning (FEMSolver.py:358-360)."""
    import warnings
    K = _K()
    m = K.Mesh(element_type="quad"); m.Rectangle((0, 0), (1, 1), 3, 2)
    d = np.arange(m.nnodes, dtype=np.float64).reshape(-1, 1) + 1.0
    bc = K.BoundaryCondition(); bc.SetDirichletCriteria(lambda: d)
    with warnings.catch_warnings(record=True) as w:
        warnings.simplefilter("always")
        sol = K.FEMSolver(verbose=False).Solve(formulation=K.HeatDiffusion(m), mesh=m,
                                               material=K.FourierConduction(2, k=1.0, area=1.0, density=1.0, c_v=1.0),
                                               boundary_condition=bc)
    assert np.array_equal(sol.sol[:, 0, 0], d[:, 0])
    assert any("Empty linear system" in str(x.message) for x in w)
    assert bc.n_free == 0 and bc.columns_in.shape == (0,)


def test_single_element_and_single_free_dof():
    from oracle import kiltro_oracle as O
    K = _K()
    pts, el = O.mesh_rectangle((0, 0), (2.0, 1.0), 1, 1)
    d = np.full((4, 1), np.nan); d[0] = 10.0; d[1] = 20.0; d[3] = -5.0          # one free DOF
    m = make_mesh(pts, el)
    bc = K.BoundaryCondition(); bc.SetDirichletCriteria(lambda: d)
    sol = K.FEMSolver(verbose=False).Solve(formulation=K.HeatDiffusion(m), mesh=m,
                                           material=K.FourierConduction(2, k=3.0, area=1.0, density=1.0, c_v=1.0),
                                           boundary_condition=bc, solver=K.LinearSolver(tol=1e-14))
    ref = O.steady_solve(pts, el, None, (3.0, 1.0, 1.0, 1.0, 0.0), d, None, "reference", False)
    assert rel_l2(sol.sol.ravel(), ref) < 1e-12


def test_pattern_with_unused_node_and_repeated_solve_reuse():
    """A node that belongs to no element yields an empty CSR row (indptr stays flat); the C oracle agrees."""
    from oracle import kiltro_oracle as O, native as ON
    K = _K()
    pts, el = O.mesh_rectangle((0, 0), (1, 1), 4, 3)
    pts = np.vstack([pts, [[5.0, 5.0]]])                                          # node 20 is unused
    m = make_mesh(pts, el)
    got = K.ComputeSparsityPattern(m, 1)
    want = ON.c_sparsity_pattern(el, pts.shape[0], 1)
    for g, w in zip(got, want):
        assert np.array_equal(g, w)
    assert got[1][-1] == got[1][-2]                                              # empty last row


def test_line_mesh_steady_large_vs_oracle():
    from oracle import kiltro_oracle as O
    K = _K()
    n = 20000
    pts, el = O.mesh_line(0.0, 1.0, n)
    vel = np.full((n + 1, 1), 0.6 * 2.0 * n)                                      # cell Peclet 0.6
    d = np.full((n + 1, 1), np.nan); d[0] = 1.0; d[-1] = 0.0
    m = make_mesh(pts, el)
    bc = K.BoundaryCondition(); bc.SetDirichletCriteria(lambda: d)
    solver = K.LinearSolver(tol=1e-12, maxiter=200000, check_every=200)
    sol = K.FEMSolver(verbose=False).Solve(formulation=K.HeatAdvectionDiffusion(m, velocity=vel), mesh=m,
                                           material=K.FourierConduction(1, k=1.0, area=1.0, density=1.0, heat_capacity_v=1.0),
                                           boundary_condition=bc, solver=solver)
    ref = O.steady_solve(pts, el, vel, (1.0, 1.0, 1.0, 1.0, 0.0), d, None, "consistent", False)
    assert solver.info["status"] == 0, solver.info
    assert rel_l2(sol.sol.ravel(), ref) < SOL_TOL, solver.info


# ------------------------------------------------------------------------------------------------- SURVEY 8(f) rows n2, n3
NEXT = Golden("next")


@pytest.mark.parametrize("name", NEXT.cases())
def test_robin_and_characteristic_galerkin_operators_golden(name):
    """kb_element_operator + kb_assemble_from_elements against the dense matrices the reference's own
    AssemblyRobinForces (Assembly.py:121-175) and AssembleCharacteristicGalerkin (Assembly.py:95-117) produce."""
    K = _K()
    c = NEXT.case(name)
    m = make_mesh(c["points"], c["elements"])
    mat = material_for(c["material"], m.ndim)
    form = K.TemperaturePG(m, velocity=c["velocity"])
    fs = K.FEMSolver(verbose=False)
    bc = K.BoundaryCondition()
    bc.SetRobinCriteria(lambda: {"type": "Convection", "flags": None, "data": float(c["robin_h"])})
    bc.ground_robin_convection = float(c["robin_tinf"])
    Kr, Fr = K.AssemblyRobinForces(fs, form.function_spaces[0], m, mat, bc)
    assert rel_max(Kr.toarray(), c["K_robin"]) < VAL_TOL and rel_max(K.to_numpy(Fr).ravel(), c["F_robin"]) < VAL_TOL
    Ku, Fu = K.AssembleCharacteristicGalerkin(form.function_spaces, form, m, mat, bc, fem_solver=fs)
    assert rel_max(Ku.toarray(), c["K_cg"]) < VAL_TOL and rel_max(K.to_numpy(Fu).ravel(), c["F_cg"]) < VAL_TOL
    # same pattern object as the stiffness matrix (what makes folding them into K a vector operation)
    assert Kr.indices.data_ptr() == fs.sparsity.indices.data_ptr() and Ku.indptr.data_ptr() == fs.sparsity.indptr.data_ptr()


def test_steady_with_robin_term():
    """include_robin=True: (K + K_e) T = b + Flux_e on the free DOFs, against a dense NumPy solve on the oracle's matrices."""
    from oracle import kiltro_oracle as O
    from scipy.sparse import csr_matrix
    K = _K()
    c = NEXT.case("quad")
    pts, el, vel = c["points"], c["elements"], c["velocity"]
    nn = pts.shape[0]
    matv = tuple(c["material"])
    h, tinf = 4.0, 300.0
    flags = np.full((nn, 1), np.nan); flags[pts[:, 0] == 0.0, 0] = 350.0; flags[pts[:, 0] == pts[:, 0].max(), 0] = 280.0
    m = make_mesh(pts, el)
    form = K.HeatAdvectionDiffusion(m, velocity=vel)
    bc = K.BoundaryCondition(); bc.SetDirichletCriteria(lambda: flags.copy())
    bc.applied_robin_convection = h; bc.ground_robin_convection = tinf
    fs = K.FEMSolver(verbose=False, include_robin=True)
    sol = fs.Solve(formulation=form, mesh=m, material=material_for(matv, 2), boundary_condition=bc,
                   solver=K.LinearSolver(tol=1e-13))
    assert fs.solver_info["status"] == 0
    A = O.assemble(pts, el, vel, matv, "consistent", False, False)
    Ke, Fe = O.robin_matrices(pts, el, h, tinf)
    R = O.assemble_values(el, nn, Ke, Fe)
    Kd = csr_matrix((A["K"] + R["data"], A["indices"], A["indptr"]), shape=(nn, nn)).toarray()
    cin, cout, td = O.dirichlet_split(flags)
    T = np.zeros(nn); T[cout] = td
    T[cin] = np.linalg.solve(Kd[np.ix_(cin, cin)], R["F"][cin] - Kd[np.ix_(cin, cout)] @ td)
    assert rel_l2(sol.sol[:, 0, 0], T) < SOL_TOL
    # the term matters: without it the solution is different
    fs0 = K.FEMSolver(verbose=False)
    bc0 = K.BoundaryCondition(); bc0.SetDirichletCriteria(lambda: flags.copy())
    sol0 = fs0.Solve(formulation=form, mesh=m, material=material_for(matv, 2), boundary_condition=bc0,
                     solver=K.LinearSolver(tol=1e-13))
    assert rel_l2(sol0.sol[:, 0, 0], T) > 1e-4


@pytest.mark.parametrize("name", ["distorted_consistent", "ex2d_dt010"])
def test_transient_characteristic_galerkin(name):
    """characteristic_galerkin=True: the explicit step with the dt^2/2 M^-1 (K_u T + F_u) correction the reference
    carries commented out (FEMSolver.py:264-266,287-292), against the dense restatement in the oracle."""
    from oracle import kiltro_oracle as O
    c = TRA.case(name)
    K, m, form, mat, bc, fs = _transient_setup(c, name, snaps="all", nincr=21)
    fs.characteristic_galerkin = True
    vel = c["velocity"] if np.any(c["velocity"]) else np.zeros_like(c["points"])
    mode = "consistent" if name.endswith("consistent") else "reference"
    ref = O.transient_characteristic_galerkin(c["points"], c["elements"], vel, tuple(c["material"]), c["dirichlet_flags"],
                                              c["neumann_flags"] if "neumann_flags" in c else None, c["initial"], 21,
                                              float(c["dt"]), mode=mode)
    T = fs.Solve(formulation=form, mesh=m, material=mat, boundary_condition=bc, solver=K.LinearSolver(tol=1e-13))
    assert T.shape == ref.shape and rel_l2(T, ref) < SOL_TOL, fs.solver_info
    if np.any(c["velocity"]):
        plain = O.transient_forward_euler(c["points"], c["elements"], vel, tuple(c["material"]), c["dirichlet_flags"],
                                          c["neumann_flags"] if "neumann_flags" in c else None, c["initial"], 21,
                                          float(c["dt"]), mode=mode)
        assert rel_l2(plain, ref) > 1e-9                   # the correction is not a no-op when there is advection
    with pytest.raises(ValueError):
        K2, m2, form2, mat2, bc2, fs2 = _transient_setup(c, name, theta=0.5, nincr=5)
        fs2.characteristic_galerkin = True
        fs2.Solve(formulation=form2, mesh=m2, material=mat2, boundary_condition=bc2, solver=K.LinearSolver())


# ------------------------------------------------------------------------------------------------- K5 index forms
@pytest.mark.parametrize("case", ["rect_bicgstab", "rect_cg", "line", "scrambled"])
def test_spmv_index_forms_bitwise(case):
    """The three index forms of the solver's SpMV -- row patterns (8 B/nnz), 16-bit column offsets (10 B/nnz), 32-bit
    indices (12 B/nnz) -- stream different bytes but multiply and add in the same order: iterates, iteration counts and
    solutions are bit-identical.  A scrambled numbering has too many row patterns and silently uses the offset form."""
    from oracle import kiltro_oracle as O
    from kiltro_b200 import _lib
    K = _K()
    lib = _lib.load()
    rng = np.random.default_rng(3)
    if case == "line":
        m = K.Mesh(element_type="line"); m.Line(0.0, 1.0, 3000)
        vel = np.full((m.nnodes, 1), 40.0)
        d = np.full((m.nnodes, 1), np.nan); d[0] = 1.0; d[-1] = 2.0
        form = K.HeatAdvectionDiffusion(m, velocity=vel); ndim = 1
    else:
        pts, el = O.mesh_rectangle((0, 0), (1, 1), 93, 71)
        if case == "scrambled":
            perm = rng.permutation(pts.shape[0]); inv = np.empty_like(perm); inv[perm] = np.arange(perm.size)
            pts = pts[perm]; el = inv[el]
        m = make_mesh(pts, el)
        vel = np.zeros(pts.shape); vel[:, 0] = 60.0; vel += 6.0 * rng.standard_normal(pts.shape)
        d = np.full((pts.shape[0], 1), np.nan); d[pts[:, 0] == 0.0] = 100.0; d[pts[:, 0] == 1.0] = 500.0
        form = K.HeatDiffusion(m) if case == "rect_cg" else K.HeatAdvectionDiffusion(m, velocity=vel); ndim = 2
    mat = material_for((1.0, 1.0, 1.0, 1.0, 0.0), ndim)
    out = {}
    try:
        for mode in ("pattern", "offset16", "index32"):
            lib.kb_spmv_force_no_patterns(0 if mode == "pattern" else 1)
            lib.kb_spmv_force_idx32(1 if mode == "index32" else 0)
            bc = K.BoundaryCondition(); bc.SetDirichletCriteria(lambda: d.copy())
            solver = K.LinearSolver(tol=1e-12, maxiter=30000, check_every=7)
            sol = K.FEMSolver(verbose=False).Solve(formulation=form, mesh=m, material=mat, boundary_condition=bc, solver=solver)
            assert solver.info["status"] == 0, (mode, solver.info)
            out[mode] = (sol.sol_device.clone(), solver.info["iterations"])
    finally:
        lib.kb_spmv_force_no_patterns(0); lib.kb_spmv_force_idx32(0)
    for mode in ("offset16", "index32"):
        assert out[mode][1] == out["pattern"][1], (mode, out[mode][1], out["pattern"][1])
        assert torch.equal(out[mode][0], out["pattern"][0]), mode


def test_write_vtk_from_device_solution(tmp_path):
    """a solution that lives on the device is streamed into the .vtu files (one per load increment)."""
    import sys
    sys.path.insert(0, str(__import__("pathlib").Path(__file__).parent))
    from test_vtk_cpu import read_vtu
    K = _K()
    m = K.Mesh(element_type="quad"); m.Rectangle((0.0, 0.0), (1.0, 0.5), 12, 7)
    d = np.full((m.nnodes, 1), np.nan); d[m.points[:, 0] == 0.0] = 1.0; d[m.points[:, 0] == 1.0] = 3.0
    bc = K.BoundaryCondition(); bc.SetDirichletCriteria(lambda: d)
    sol = K.FEMSolver(verbose=False, number_of_load_increments=2).Solve(
        formulation=K.HeatDiffusion(m), mesh=m, material=K.FourierConduction(2, k=1.0, area=1.0, density=1.0, c_v=1.0),
        boundary_condition=bc, solver=K.LinearSolver(tol=1e-13))
    assert sol._sol_h is None and sol.sol_device is not None          # still on the device when the writer starts
    files = sol.WriteVTK(str(tmp_path / "field.vtu"))
    assert len(files) == 2
    for i, f in enumerate(files):
        v = read_vtu(f)
        assert np.array_equal(v["T"], sol.sol[:, 0, i]) and np.array_equal(v["connectivity"], m.elements.ravel())
        assert np.array_equal(v["points"][:, :2], m.points)
