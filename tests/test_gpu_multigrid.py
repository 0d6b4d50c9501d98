"""EXPERIMENTAL opt-in multigrid preconditioner (kiltro_b200/multigrid.py): same solution as the prescribed
Jacobi-BiCGSTAB (<= 1e-8 relative L2, the bar of every solve test) in far fewer iterations."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def rel_l2(a, b):
    return float(np.linalg.norm(a - b) / np.linalg.norm(b))


@pytest.mark.parametrize("nx,ny,peclet", [(128, 96, 0.45), (99, 81, 0.45), (128, 128, 0.0), (257, 190, 0.3)])
def test_multigrid_preconditioned_solve_matches_jacobi_bicgstab(nx, ny, peclet):
    import kiltro_b200 as K
    from kiltro_b200.multigrid import GridMultigrid
    m = K.Mesh(element_type="quad"); m.Rectangle((0.0, 0.0), (1.0, 0.8), nx, ny)
    P = m.device_points()
    nn = m.nnodes
    g = torch.Generator(device="cuda"); g.manual_seed(1)
    flags = torch.full((nn, 1), float("nan"), dtype=torch.float64, device="cuda")
    flags[P[:, 0] == 0.0, 0] = 100.0
    flags[P[:, 0] == 1.0, 0] = 500.0
    mat = K.FourierConduction(2, k=1.0, area=1.0, density=1.0, heat_capacity_v=1.0, q=0.0)
    if peclet > 0.0:
        U = peclet * 2.0 * nx
        vel = torch.zeros((nn, 2), dtype=torch.float64, device="cuda"); vel[:, 0] = U
        vel += 0.1 * U * torch.randn((nn, 2), dtype=torch.float64, device="cuda", generator=g)
        form = K.HeatAdvectionDiffusion(m, velocity=vel)
    else:
        form = K.HeatDiffusion(m)

    def solve(solver):
        bc = K.BoundaryCondition(); bc.SetDirichletCriteria(lambda: flags)
        out = K.FEMSolver(verbose=False).Solve(formulation=form, mesh=m, material=mat, boundary_condition=bc, solver=solver)
        return out.sol.ravel().copy()

    ref_solver = K.LinearSolver(tol=1e-12)
    ref = solve(ref_solver)
    bc0 = K.BoundaryCondition(); bc0.SetDirichletCriteria(lambda: flags)
    mg = GridMultigrid(m, form, mat, bc0)
    assert len(mg.levels) >= 3
    mg_solver = K.LinearSolver(tol=1e-11, preconditioner=mg)
    got = solve(mg_solver)
    info = mg_solver.info
    assert info["status"] == 0 and info["method"] == "bicgstab+multigrid", info
    assert rel_l2(got, ref) < 1e-8, (info, ref_solver.info)
    assert info["iterations"] <= 40 and info["iterations"] * 4 < ref_solver.info["iterations"], (info, ref_solver.info)
