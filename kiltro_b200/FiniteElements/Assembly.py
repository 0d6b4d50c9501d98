"""Global assembly, mirroring kiltro/FiniteElements/Assembly.py:8-92.

Assemble(fem_solver, function_spaces, formulation, mesh, material, boundary_condition) -> (K, Flux, M) with K / M as
DeviceCSR on the pattern cached on fem_solver (FEMSolver.ComputeSparsityFEM) and Flux a device (ndof, 1) tensor.  The
per-element Python loop of the reference (Assembly.py:46-72) becomes either
  * the fused kernel (element matrices on chip; csrc/assemble_tmpl.cu when the tiles of the mesh repeat -- every
    Mesh.Rectangle mesh -- else the generic csrc/assemble_fused.cu; fem_solver.assembly = "fused_generic" forces the
    latter), or
  * the two-step seam K1 (kb_element_matrices) + K3 (kb_assemble_from_elements) when fem_solver.assembly == "two_pass".
The dense AssemblyRobinForces pass whose result the reference discards (Assembly.py:90,121-175) is not run."""
import torch

from .. import _lib, device as D
from ..sparse import DeviceCSR

__all__ = ['Assemble']


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
    if method not in ("auto", "fused", "fused_generic", "two_pass"):
        raise ValueError("assembly must be 'auto', 'fused', 'fused_generic' or 'two_pass'")
    use_fused = False
    if method != "two_pass" and formulation.nvar == 1:
        from .fused import assemble_fused, build_tile_plan
        if sp.tile_plan is None:
            sp.tile_plan = build_tile_plan(sp, mesh)             # symbolic, cached with the pattern
        use_fused = sp.tile_plan.fits
        if method in ("fused", "fused_generic") and not use_fused:
            raise _lib.KiltroB200Error("the tile plan of this mesh does not fit the fused assembly kernel: %r"
                                       % (sp.tile_plan.info,))
    fem_solver.assembly_used = "fused" if use_fused else "two_pass"
    if use_fused:
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
