"""Symbolic phase, mirroring kiltro/FiniteElements/ComputeSparsityPattern.pyx:44-112 (+ ComputeSparsityPattern.h).

ComputeSparsityPattern(mesh, nvar) returns the reference's four int32 arrays (indices, indptr, data_local_indices,
data_global_indices) as NumPy arrays; ComputeSparsityPatternDevice returns a SparsityPattern whose arrays stay in HBM and
which also carries the row-owner gather map of the numeric phase (csrc/pattern.cu)."""
import numpy as np
import torch

from .. import _lib, device as D


class SparsityPattern(object):
    """Device-resident result of the symbolic phase (K2)."""

    def __init__(self):
        self.indices = self.indptr = self.data_local_indices = self.data_global_indices = None
        self.seg_ptr = self.gather_src = self.diag_pos = None
        self.nnz = 0
        self.nrows = 0
        self.npe = 0
        self.nvar = 1
        self.tile_plan = None
        self.grid = None          # None: not checked; False / (nxn, nyn): Mesh.Rectangle numbering (fused.grid_shape)
        self.grid_scratch = None
        self.line = None          # None: not checked; True / False: Mesh.Line numbering (fused.line_numbered)

    def to_numpy(self):
        return (D.to_numpy(self.indices), D.to_numpy(self.indptr), D.to_numpy(self.data_local_indices),
                D.to_numpy(self.data_global_indices))


def ComputeSparsityPatternDevice(mesh, nvar=1):
    lib = _lib.load()
    el = mesh.device_elements()
    nelem, npe = int(el.shape[0]), int(el.shape[1])
    nnode = int(mesh.nnodes)
    nent = nelem * (npe * nvar) ** 2
    nrows = nnode * nvar
    sp = SparsityPattern()
    sp.nrows, sp.npe, sp.nvar = nrows, npe, nvar
    indptr = D.empty(nrows + 1, torch.int32)
    indices = D.empty(max(nent, 1), torch.int32)
    dl = D.empty(max(nent, 1), torch.int32); dg = D.empty(max(nent, 1), torch.int32)
    seg = D.empty(nent + 1, torch.int32); src = D.empty(max(nent, 1), torch.int32)
    diag = D.empty(nrows, torch.int32)
    nnz = _lib.I64(0)
    _lib.check(lib.kb_pattern_build(D.ptr(el), nelem, npe, nnode, nvar, D.ptr(indptr), D.ptr(indices), D.ptr(dl),
                                    D.ptr(dg), D.ptr(seg), D.ptr(src), D.ptr(diag), nnz, D.stream_ptr()))
    sp.nnz = int(nnz.value)
    sp.indptr = indptr
    sp.indices = indices[:sp.nnz].clone()            # release the over-allocation (.pyx:91)
    sp.data_local_indices = dl[:nent]
    sp.data_global_indices = dg[:nent]
    sp.seg_ptr = seg[:sp.nnz + 1].clone()
    sp.gather_src = src[:nent]
    sp.diag_pos = diag
    return sp


def ComputeSparsityPattern(mesh, nvar=1):
    """Drop-in for the reference's Cython entry point (.pyx:44): four int32 NumPy arrays."""
    return ComputeSparsityPatternDevice(mesh, nvar).to_numpy()
