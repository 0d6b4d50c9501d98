"""Global assembly, mirroring kiltro/FiniteElements/Assembly.py:8-92.

Assemble(fem_solver, function_spaces, formulation, mesh, material, boundary_condition) -> (K, Flux, M) with K / M as
DeviceCSR on the pattern cached on fem_solver (FEMSolver.ComputeSparsityFEM) and Flux a device (ndof, 1) tensor.  The
per-element Python loop of the reference (Assembly.py:46-72) becomes either
  * the fused kernels (element matrices on chip): csrc/assemble_grid.cu for quad meshes numbered like Mesh.Rectangle
    (element matrices in registers, exchanged by warp shuffles; "auto" picks it, "fused_grid" demands it), else the tile
    kernels -- csrc/assemble_tmpl.cu when the tiles of the mesh repeat, the generic csrc/assemble_fused.cu otherwise
    (fem_solver.assembly = "fused" / "fused_generic" force them), or
  * the two-step seam K1 (kb_element_matrices) + K3 (kb_assemble_from_elements) when fem_solver.assembly == "two_pass".
The dense AssemblyRobinForces pass whose result the reference discards (Assembly.py:90,121-175) is not run."""
import numpy as np
import torch

from .. import _lib, device as D
from ..sparse import DeviceCSR

GRID_IN_AUTO = True         # assembly="auto" picks the grid kernel (csrc/assemble_grid.cu) on Mesh.Rectangle-numbered meshes

__all__ = ['Assemble', 'AssemblyRobinForces', 'AssembleCharacteristicGalerkin']


def Assemble(fem_solver, function_spaces, formulation, mesh, material, boundary_condition):
    if getattr(fem_solver, "recompute_sparsity_pattern", False):
        raise ValueError("not implemented yet")                  # Assembly.py:54
    if getattr(fem_solver, "squeeze_sparsity_pattern", False):
        raise ValueError("not implemented yet")                  # Assembly.py:57
    lib = _lib.load()
    sp = fem_solver.ComputeSparsityFEM(mesh, formulation)
    transient = fem_solver.analysis_type != 'steady'             # Assembly.py:39-40
    nnode = int(mesh.nnodes)
    npe = sp.npe
    K = D.empty(sp.nnz)
    M = D.empty(sp.nnz) if transient else None
    Flux = D.empty((nnode * formulation.nvar, 1))
    method = getattr(fem_solver, "assembly", "auto")
    if method not in ("auto", "fused", "fused_grid", "fused_generic", "two_pass"):
        raise ValueError("assembly must be 'auto', 'fused', 'fused_grid', 'fused_generic' or 'two_pass'")
    use_fused = use_grid = False
    if method in (("auto", "fused_grid") if GRID_IN_AUTO else ("fused_grid",)) and formulation.nvar == 1 and formulation.ndim == 2:
        from .fused import assemble_grid, grid_shape
        use_grid = grid_shape(sp, mesh) is not None              # Mesh.Rectangle numbering, verified on device (cached)
    use_line = False
    if method in (("auto", "fused_grid") if GRID_IN_AUTO else ("fused_grid",)) and formulation.nvar == 1 and formulation.ndim == 1:
        from .fused import assemble_line, line_numbered
        use_grid = use_line = bool(line_numbered(sp, mesh))      # Mesh.Line numbering: the 1-D grid form
    if method == "fused_grid" and not use_grid:
        raise _lib.KiltroB200Error("assembly='fused_grid' needs a quad-4 mesh numbered like Mesh.Rectangle or a line mesh "
                                   "numbered like Mesh.Line")
    if not use_grid and method != "two_pass" and formulation.nvar == 1:
        from .fused import assemble_fused, build_tile_plan
        if sp.tile_plan is None:
            sp.tile_plan = build_tile_plan(sp, mesh)             # symbolic, cached with the pattern
        use_fused = sp.tile_plan.fits
        if method in ("fused", "fused_generic") and not use_fused:
            raise _lib.KiltroB200Error("the tile plan of this mesh does not fit the fused assembly kernel: %r"
                                       % (sp.tile_plan.info,))
    fem_solver.assembly_used = "fused" if use_fused else "two_pass"
    if use_line:
        assemble_line(fem_solver, sp, formulation, mesh, material, K, M, Flux)
    elif use_grid:
        assemble_grid(fem_solver, sp, formulation, mesh, material, K, M, Flux)
    elif use_fused:
        assemble_fused(fem_solver, sp, formulation, mesh, material, K, M, Flux)
    else:
        Ke, Fe, Me = formulation.GetAllElementalMatrices(mesh, material, want_mass=transient)
        _lib.check(lib.kb_assemble_from_elements(npe, nnode, sp.nnz, D.ptr(sp.seg_ptr), D.ptr(sp.gather_src),
                                                 D.ptr(sp.diag_pos), D.ptr(Ke), D.ptr(Me), D.ptr(Fe), D.ptr(K), D.ptr(M),
                                                 D.ptr(Flux), D.stream_ptr()))
    shape = (nnode * formulation.nvar, nnode * formulation.nvar)
    convection = DeviceCSR(K, sp.indices, sp.indptr, shape, symmetric=formulation.is_symmetric)
    # storage details the reference stashes on the solver (Assembly.py:82-83), in MiB
    fem_solver.spmat = K.numel() * 8 / 1024. / 1024.
    fem_solver.ijv = (sp.indptr.numel() * 4 + sp.indices.numel() * 4 + K.numel() * 8) / 1024. / 1024.
    mass = DeviceCSR(M, sp.indices, sp.indptr, shape, symmetric=True) if transient else []
    return convection, Flux, mass


# ------------------------------------------------------------------------------------------------ SURVEY 8(f) rows n2, n3
def _assemble_operator(fem_solver, mesh, op, p0, p1, velocity):
    """Element operator (csrc/operators.cu) -> CSR values on the cached pattern (csrc/assemble_gather.cu)."""
    lib = _lib.load()
    if fem_solver.sparsity is None:                              # same cache as ComputeSparsityFEM (FEMSolver.py:313-325)
        from .ComputeSparsityPattern import ComputeSparsityPatternDevice
        fem_solver.sparsity = ComputeSparsityPatternDevice(mesh, 1)
        fem_solver.is_sparsity_pattern_computed = True
    sp = fem_solver.sparsity
    if sp.nvar != 1:
        raise ValueError("only nvar = 1 is available")
    pts, el = mesh.device_points(), mesh.device_elements()
    ne, npe = int(el.shape[0]), int(el.shape[1])
    nnode = int(mesh.nnodes)
    Ke = D.empty((ne, npe * npe)); Fe = D.empty((ne, npe))
    _lib.check(lib.kb_element_operator(int(mesh.ndim), op, D.ptr(pts), D.ptr(el), ne, D.ptr(velocity), float(p0),
                                       float(p1), D.ptr(Ke), D.ptr(Fe), D.stream_ptr()))
    V = D.empty(sp.nnz); F = D.empty((nnode, 1))
    _lib.check(lib.kb_assemble_from_elements(npe, nnode, sp.nnz, D.ptr(sp.seg_ptr), D.ptr(sp.gather_src),
                                             D.ptr(sp.diag_pos), D.ptr(Ke), D.ptr(None), D.ptr(Fe), D.ptr(V), D.ptr(None),
                                             D.ptr(F), D.stream_ptr()))
    return DeviceCSR(V, sp.indices, sp.indptr, (nnode, nnode), symmetric=True), F


def AssemblyRobinForces(fem_solver, function_space, mesh, material, boundary_condition):
    """kiltro/FiniteElements/Assembly.py:121-175: the Newton-cooling (lateral convection) term, K_e = sum_e h int N (x) N dV
    and Flux_e (as shipped, see csrc/element.cuh), with h = boundary_condition.applied_robin_convection and
    T_inf = boundary_condition.ground_robin_convection.  Sparse on the cached pattern instead of dense (npoints^2)."""
    h, tinf = boundary_condition.applied_robin_convection, boundary_condition.ground_robin_convection
    if not np.isscalar(h) or not np.isscalar(tinf):
        raise ValueError("applied_robin_convection / ground_robin_convection must be scalars (as in the reference)")
    return _assemble_operator(fem_solver, mesh, _lib.KB_OP_ROBIN, h, tinf, None)


def AssembleCharacteristicGalerkin(function_spaces, formulation, mesh, material, boundary_condition, fem_solver=None):
    """kiltro/FiniteElements/Assembly.py:95-117 over GetElementalCharacteristicGalerkin (VariationalPrinciple.py:264-308):
    K_u = sum_e int (u.grad N) (x) (u.grad N) dV, Flux_u = sum_e q area int (u.grad N) dV.  Sparse instead of dense.
    fem_solver (optional, extension): the solver whose cached sparsity pattern is reused."""
    if fem_solver is None:
        from ..FEMSolver import FEMSolver
        fem_solver = FEMSolver(verbose=False)
    return _assemble_operator(fem_solver, mesh, _lib.KB_OP_CHARGALERKIN, material.q * material.area, 0.0,
                              formulation.device_velocity())
