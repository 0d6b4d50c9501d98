"""Parity at BASELINE.json's full single-GPU size (S16: 4000 x 4000 nodes, 16 M DOF) through size-independent properties:
the oracle cannot run here in reasonable time, so the checks are closed forms, conservation identities, structural
invariants and bitwise determinism."""
import math

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
torch = pytest.importorskip("torch")

NPS = 4000


@pytest.fixture(scope="module")
def big():
    import kiltro_b200 as K
    mesh = K.Mesh(element_type="quad")
    mesh.Rectangle(lower_left_point=(0.0, 0.0), upper_right_point=(1.0, 1.0), nx=NPS - 1, ny=NPS - 1)
    fs = K.FEMSolver(analysis_type="transient", verbose=False)
    mat = K.FourierConduction(2, k=1.0, area=1.0, density=2.0, heat_capacity_v=1.5, q=3.0)
    return K, mesh, fs, mat


def test_pattern_structure_fullsize(big):
    K, mesh, fs, mat = big
    form = K.HeatDiffusion(mesh)
    sp = fs.ComputeSparsityFEM(mesh, form)
    n = NPS
    assert sp.nnz == (3 * n - 2) ** 2
    ip = sp.indptr.to(torch.int64)
    lens = ip[1:] - ip[:-1]
    assert int(ip[0]) == 0 and int(ip[-1]) == sp.nnz and int(lens.min()) == 4 and int(lens.max()) == 9
    assert int((lens == 4).sum()) == 4 and int((lens == 6).sum()) == 4 * (n - 2)
    # strictly ascending columns inside every row
    d = sp.indices[1:].to(torch.int64) - sp.indices[:-1].to(torch.int64)
    row_start = torch.zeros(sp.nnz, dtype=torch.bool, device="cuda"); row_start[ip[1:-1]] = True
    assert bool(((d > 0) | row_start[1:]).all())
    # the reference's scatter maps: data_global is a surjection onto [0, nnz) with 1..4 contributors per entry
    cnt = torch.bincount(sp.data_global_indices.to(torch.int64), minlength=sp.nnz)
    assert int(cnt.min()) == 1 and int(cnt.max()) == 4 and int(cnt.sum()) == 16 * mesh.nelem


def test_assembly_identities_fullsize(big):
    K, mesh, fs, mat = big
    form = K.HeatDiffusion(mesh)
    Kc, Flux, M = K.Assemble(fs, form.function_spaces, form, mesh, mat, K.BoundaryCondition())
    assert fs.assembly_used == "fused_grid" and fs.sparsity.grid == (NPS, NPS)
    nn = mesh.nnodes
    ones = torch.ones(nn, dtype=torch.float64, device="cuda")
    kmax = float(Kc.data.abs().max())
    # pure diffusion: constants are in the null space (row sums vanish)
    assert float(Kc.dot(ones).abs().max()) < 1e-12 * kmax
    # mass matrix and source vector integrate constants exactly: sum M = rho c_v |Omega|, sum Flux = area q |Omega|
    assert abs(float(M.dot(ones).sum()) - 2.0 * 1.5 * 1.0) < 1e-9
    assert abs(float(Flux.sum()) - 1.0 * 3.0 * 1.0) < 1e-9
    # symmetry through the bilinear form
    g = torch.Generator(device="cuda"); g.manual_seed(1)
    x = torch.randn(nn, dtype=torch.float64, device="cuda", generator=g)
    y = torch.randn(nn, dtype=torch.float64, device="cuda", generator=g)
    a, b = float(torch.dot(x, Kc.dot(y))), float(torch.dot(y, Kc.dot(x)))
    assert abs(a - b) <= 1e-10 * max(abs(a), abs(b), 1.0)
    # deterministic: a second assembly and the two-pass seam give the same bits
    K2 = K.Assemble(fs, form.function_spaces, form, mesh, mat, K.BoundaryCondition())[0]
    assert torch.equal(K2.data, Kc.data)
    fs2 = K.FEMSolver(analysis_type="transient", verbose=False, assembly="two_pass")
    K3, F3, M3 = K.Assemble(fs2, form.function_spaces, form, mesh, mat, K.BoundaryCondition())
    assert torch.equal(K3.data, Kc.data) and torch.equal(M3.data, M.data) and torch.equal(F3, Flux)
    del K3, M3, F3, fs2
    # ... and so does the template form of the tile kernels (the default before the grid kernel)
    fs3 = K.FEMSolver(analysis_type="transient", verbose=False, assembly="fused")
    K4, F4, M4 = K.Assemble(fs3, form.function_spaces, form, mesh, mat, K.BoundaryCondition())
    assert fs3.assembly_used == "fused_template" and fs3.sparsity.tile_plan.info["structured"]
    assert torch.equal(K4.data, Kc.data) and torch.equal(M4.data, M.data) and torch.equal(F4, Flux)


def test_steady_closed_form_fullsize(big):
    """Uniform velocity (cell Peclet 0.45), Dirichlet x=0 -> 100, x=1 -> 500: the consistent Galerkin scheme has the exact
    discrete solution T_i = T_0 + dT (r^i - 1)/(r^n - 1), r = (1+Pe)/(1-Pe), independent of y."""
    K, mesh, _, _ = big
    n = NPS - 1
    h = 1.0 / n
    Pe = 0.45
    nn = mesh.nnodes
    vel = torch.zeros((nn, 2), dtype=torch.float64, device="cuda"); vel[:, 0] = Pe * 2.0 / h
    flags = torch.full((nn, 1), float("nan"), dtype=torch.float64, device="cuda")
    left = torch.arange(NPS, device="cuda") * NPS
    flags[left, 0] = 100.0; flags[left + NPS - 1, 0] = 500.0
    bc = K.BoundaryCondition(); bc.SetDirichletCriteria(lambda: flags)
    mat = K.FourierConduction(2, k=1.0, area=1.0, density=1.0, heat_capacity_v=1.0, q=0.0)
    solver = K.LinearSolver(linear_solver_type="bicgstab", tol=1e-11, check_every=100)
    fs = K.FEMSolver(verbose=False)
    sol = fs.Solve(formulation=K.HeatAdvectionDiffusion(mesh, velocity=vel), mesh=mesh, material=mat, boundary_condition=bc,
                   solver=solver)
    assert solver.info["status"] == 0, solver.info
    lr = math.log((1 + Pe) / (1 - Pe))
    i = np.arange(NPS, dtype=np.float64)
    with np.errstate(over="ignore"):
        exact = 100.0 + 400.0 * np.exp((i - n) * lr) * (-np.expm1(-i * lr)) / (-np.expm1(-n * lr))
    exact[0] = 100.0
    T = sol.sol_device[:, 0, 0].reshape(NPS, NPS)
    ex = torch.from_numpy(exact).cuda()
    err = float((T - ex[None, :]).abs().max()) / 500.0
    assert err < 1e-8, (err, solver.info)
    # residual of the full system at the returned solution (free rows)
    Kc = K.Assemble(fs, fs.sparsity and K.HeatAdvectionDiffusion(mesh, velocity=vel).function_spaces,
                    K.HeatAdvectionDiffusion(mesh, velocity=vel), mesh, mat, bc)[0]
    r = Kc.dot(T.reshape(-1))
    free = bc.renum >= 0
    assert float(r[free].abs().max()) < 1e-6 * float(Kc.data.abs().max()) * 500.0


def test_transient_heat_conservation_fullsize(big):
    """Config 3/5 at full size through an invariant: with no Dirichlet DOF and no loads, 1^T K = 0 for pure diffusion, so
    the total heat 1^T M T^n is conserved by every theta-step (M dT = -dt K T  =>  1^T M dT = 0), and the solution obeys the
    discrete maximum principle loosely (stays inside the initial range up to the explicit scheme's small overshoot)."""
    K, mesh, _, mat = big
    nn = mesh.nnodes
    g = torch.Generator(device="cuda"); g.manual_seed(3)
    T0 = 100.0 + 50.0 * torch.rand((nn, 1), dtype=torch.float64, device="cuda", generator=g)
    flags = torch.full((nn, 1), float("nan"), dtype=torch.float64, device="cuda")
    import warnings
    for theta in (0.0, 0.5):
        bc = K.BoundaryCondition()
        bc.SetInitialConditions(lambda: T0); bc.SetDirichletCriteria(lambda: flags)
        fs = K.FEMSolver(analysis_type="transient", number_of_increments=6, theta=theta, verbose=False)
        fs.return_device = True
        fs.snapshot_increments = [0, 5]
        form = K.HeatDiffusion(mesh)
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            S = fs.Solve(formulation=form, mesh=mesh, material=mat, boundary_condition=bc, solver=K.LinearSolver(tol=1e-12))
        assert fs.solver_info["status"] == 0, fs.solver_info
        M = K.Assemble(fs, form.function_spaces, form, mesh, mat, bc)[2]
        h0, h5 = float(M.dot(S[:, 0]).sum()), float(M.dot(S[:, 1]).sum())
        assert abs(h5 - h0) <= 1e-9 * abs(h0), (theta, h0, h5, fs.solver_info)
        assert float(S[:, 1].min()) > 90.0 and float(S[:, 1].max()) < 160.0
        assert float((S[:, 1] - S[:, 0]).abs().max()) > 1e-3          # something did happen


def test_next_row_operators_identities_fullsize(big):
    """SURVEY 8(f) n2 / n3 at 16 M DOF through identities: the Robin matrix is h / (rho c_v) times the mass matrix
    (K_e = h int N (x) N), its entries sum to h |Omega| and Flux_e to h T_inf |Omega|; the characteristic-Galerkin matrix
    int (u.grad N)(x)(u.grad N) is symmetric, annihilates constants and, for a uniform velocity u, gives
    x^T K_u x = |u|^2 |Omega| on the linear field x = (u . X)."""
    K, mesh, fs, mat = big
    nn = mesh.nnodes
    form = K.HeatDiffusion(mesh)
    M = K.Assemble(fs, form.function_spaces, form, mesh, mat, K.BoundaryCondition())[2]
    bc = K.BoundaryCondition(); bc.applied_robin_convection = 7.5; bc.ground_robin_convection = 293.0
    Kr, Fr = K.AssemblyRobinForces(fs, form.function_spaces[0], mesh, mat, bc)
    assert float((Kr.data - (7.5 / (2.0 * 1.5)) * M.data).abs().max()) <= 1e-13 * float(Kr.data.abs().max())
    assert abs(float(Kr.data.sum()) - 7.5) < 1e-9 and abs(float(Fr.sum()) - 7.5 * 293.0) < 1e-6
    u = (3.0, -2.0)
    vel = torch.zeros((nn, 2), dtype=torch.float64, device="cuda"); vel[:, 0] = u[0]; vel[:, 1] = u[1]
    formu = K.HeatAdvectionDiffusion(mesh, velocity=vel)
    Ku, Fu = K.AssembleCharacteristicGalerkin(formu.function_spaces, formu, mesh, mat, bc, fem_solver=fs)
    ones = torch.ones(nn, dtype=torch.float64, device="cuda")
    kmax = float(Ku.data.abs().max())
    assert float(Ku.dot(ones).abs().max()) < 1e-11 * kmax
    P = mesh.device_points()
    x = u[0] * P[:, 0] + u[1] * P[:, 1]
    assert abs(float(torch.dot(x, Ku.dot(x))) - (u[0] ** 2 + u[1] ** 2) ** 2) < 1e-8 * (u[0] ** 2 + u[1] ** 2) ** 2
    g = torch.Generator(device="cuda"); g.manual_seed(2)
    a = torch.randn(nn, dtype=torch.float64, device="cuda", generator=g)
    b = torch.randn(nn, dtype=torch.float64, device="cuda", generator=g)
    assert abs(float(torch.dot(a, Ku.dot(b)) - torch.dot(b, Ku.dot(a)))) < 1e-9 * kmax * nn ** 0.5
    # Flux_u = q area int u.grad N: sums to zero (sum_a grad N_a = 0)
    assert abs(float(Fu.sum())) < 1e-9 * float(Fu.abs().max()) * nn ** 0.5


def test_permuted_numbering_at_scale():
    """The generic paths at scale (1 M DOF): the same mesh with RANDOMLY PERMUTED node numbers and shuffled elements has no
    repeating tiles (generic fused kernel, per-(row, element) records from HBM), a band far beyond 16 bits and far more
    than 256 row patterns (32-bit SpMV) -- and must give the permuted solution of the structured solve."""
    import kiltro_b200 as K
    n = 1000
    m = K.Mesh(element_type="quad"); m.Rectangle((0.0, 0.0), (1.0, 1.0), n - 1, n - 1)
    nn = m.nnodes
    g = torch.Generator(device="cuda"); g.manual_seed(11)
    h = 1.0 / (n - 1)
    vel = torch.zeros((nn, 2), dtype=torch.float64, device="cuda"); vel[:, 0] = 0.45 * 2.0 / h
    vel += 0.1 * vel[:, :1].abs().max() * torch.randn((nn, 2), dtype=torch.float64, device="cuda", generator=g)
    flags = torch.full((nn, 1), float("nan"), dtype=torch.float64, device="cuda")
    left = torch.arange(n, device="cuda") * n
    flags[left, 0] = 100.0; flags[left + n - 1, 0] = 500.0
    mat = K.FourierConduction(2, k=1.0, area=1.0, density=1.0, heat_capacity_v=1.0, q=0.0)

    def solve(mesh, v, f):
        bc = K.BoundaryCondition(); bc.SetDirichletCriteria(lambda: f)
        solver = K.LinearSolver(linear_solver_type="bicgstab", tol=1e-11, check_every=50, maxiter=20000)
        fs = K.FEMSolver(verbose=False)
        sol = fs.Solve(formulation=K.HeatAdvectionDiffusion(mesh, velocity=v), mesh=mesh, material=mat,
                       boundary_condition=bc, solver=solver)
        assert solver.info["status"] == 0, solver.info
        return sol.sol_device[:, 0, 0], fs

    ref, fs0 = solve(m, vel, flags)
    assert fs0.assembly_used == "fused_grid"
    perm = torch.randperm(nn, device="cuda", generator=g)          # new id of old node i = perm[i]
    inv = torch.empty_like(perm); inv[perm] = torch.arange(nn, device="cuda")
    P2 = m.device_points()[inv].contiguous()                       # coordinates listed in the new numbering
    el = perm[m.device_elements().long()]
    el = el[torch.randperm(el.shape[0], device="cuda", generator=g)].contiguous()
    m2 = K.Mesh(element_type="quad"); m2.set_arrays(P2, el)
    got, fs2 = solve(m2, vel[inv].contiguous(), flags[inv].contiguous())
    assert fs2.assembly_used in ("fused", "two_pass") and not fs2.sparsity.tile_plan.templated
    err = float((got[perm] - ref).abs().max()) / 500.0
    assert err < 1e-8, err
