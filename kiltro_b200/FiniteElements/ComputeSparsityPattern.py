"""Symbolic phase, mirroring kiltro/FiniteElements/ComputeSparsityPattern.pyx:44-112 (+ ComputeSparsityPattern.h).

ComputeSparsityPattern(mesh, nvar) returns the reference's four int32 arrays (indices, indptr, data_local_indices,
data_global_indices) as NumPy arrays; ComputeSparsityPatternDevice returns a SparsityPattern whose arrays stay in HBM and
which also carries the row-owner gather map of the numeric phase (csrc/pattern.cu)."""
import numpy as np
import torch

from .. import _lib, device as D


class SparsityPattern(object):
    """Device-resident result of the symbolic phase (K2).  On a mesh numbered like Mesh.Rectangle indptr / indices /
    diag_pos are written down from the shape (kb_pattern_grid); the reference's two scatter maps and the gather map of the
    two-pass seam, which only the API seam and the non-grid assembly paths read, are then built by the sort-based phase
    on first access."""
    _LAZY = ("data_local_indices", "data_global_indices", "seg_ptr", "gather_src")

    def __init__(self):
        self.indices = self.indptr = self.diag_pos = None
        self._maps = {}
        self._maps_builder = None
        self.nnz = 0
        self.nrows = 0
        self.npe = 0
        self.nvar = 1
        self.tile_plan = None
        self.grid = None          # None: not checked; False / (nxn, nyn): Mesh.Rectangle numbering (fused.grid_shape)
        self.grid_scratch = None
        self.line = None          # None: not checked; True / False: Mesh.Line numbering (fused.line_numbered)

    def _map(self, name):
        if name not in self._maps and self._maps_builder is not None:
            builder, self._maps_builder = self._maps_builder, None
            builder(self)
        return self._maps.get(name)

    data_local_indices = property(lambda self: self._map("data_local_indices"))
    data_global_indices = property(lambda self: self._map("data_global_indices"))
    seg_ptr = property(lambda self: self._map("seg_ptr"))
    gather_src = property(lambda self: self._map("gather_src"))

    def to_numpy(self):
        return (D.to_numpy(self.indices), D.to_numpy(self.indptr), D.to_numpy(self.data_local_indices),
                D.to_numpy(self.data_global_indices))


def _sorted_build(sp, mesh, nvar, keep_pattern):
    """The sort-based symbolic phase (kb_pattern_build): everything, or (keep_pattern) only the maps of a pattern that
    kb_pattern_grid has already written."""
    lib = _lib.load()
    el = mesh.device_elements()
    nelem, npe = int(el.shape[0]), int(el.shape[1])
    nnode = int(mesh.nnodes)
    nent = nelem * (npe * nvar) ** 2
    nrows = nnode * nvar
    indptr = D.empty(nrows + 1, torch.int32)
    indices = D.empty(max(nent, 1), torch.int32)
    dl = D.empty(max(nent, 1), torch.int32); dg = D.empty(max(nent, 1), torch.int32)
    seg = D.empty(nent + 1, torch.int32); src = D.empty(max(nent, 1), torch.int32)
    diag = D.empty(nrows, torch.int32)
    nnz = _lib.I64(0)
    _lib.check(lib.kb_pattern_build(D.ptr(el), nelem, npe, nnode, nvar, D.ptr(indptr), D.ptr(indices), D.ptr(dl),
                                    D.ptr(dg), D.ptr(seg), D.ptr(src), D.ptr(diag), nnz, D.stream_ptr()))
    if keep_pattern:
        if int(nnz.value) != sp.nnz:
            raise _lib.KiltroB200Error("the sort-based pattern disagrees with the grid pattern (nnz %d vs %d)" % (int(nnz.value), sp.nnz))
    else:
        sp.nnz = int(nnz.value)
        sp.indptr = indptr
        sp.indices = indices[:sp.nnz].clone()        # release the over-allocation (.pyx:91)
        sp.diag_pos = diag
    sp._maps = {"data_local_indices": dl[:nent], "data_global_indices": dg[:nent], "seg_ptr": seg[:sp.nnz + 1].clone(),
                "gather_src": src[:nent]}


GRID_SHORTCUT = True        # Mesh.Rectangle-numbered meshes: pattern from the shape, maps on first use


def ComputeSparsityPatternDevice(mesh, nvar=1):
    el = mesh.device_elements()
    nelem, npe = int(el.shape[0]), int(el.shape[1])
    nnode = int(mesh.nnodes)
    sp = SparsityPattern()
    sp.nrows, sp.npe, sp.nvar = nnode * nvar, npe, nvar
    if GRID_SHORTCUT and nvar == 1 and npe == 4 and nelem > 0:
        from .fused import grid_shape
        shape = grid_shape(sp, mesh)                 # verified on device, cached on sp
        if shape is not None:
            lib = _lib.load()
            nxn, nyn = shape
            nnz = (3 * nxn - 2) * (3 * nyn - 2)
            sp.indptr = D.empty(nnode + 1, torch.int32)
            sp.indices = D.empty(nnz, torch.int32)
            sp.diag_pos = D.empty(nnode, torch.int32)
            out = _lib.I64(0)
            _lib.check(lib.kb_pattern_grid(nxn, nyn, D.ptr(sp.indptr), D.ptr(sp.indices), D.ptr(sp.diag_pos), out, D.stream_ptr()))
            sp.nnz = int(out.value)
            sp._maps_builder = lambda s, m=mesh, v=nvar: _sorted_build(s, m, v, True)
            return sp
    _sorted_build(sp, mesh, nvar, False)
    return sp


def ComputeSparsityPattern(mesh, nvar=1):
    """Drop-in for the reference's Cython entry point (.pyx:44): four int32 NumPy arrays."""
    return ComputeSparsityPatternDevice(mesh, nvar).to_numpy()
