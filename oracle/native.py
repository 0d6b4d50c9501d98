"""ctypes access to the compiled oracle pieces -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

  * oracle/_build/libkiltro_oracle.so      : plain-C restatement (oracle/pattern_oracle.c)
  * oracle/_ref/libkiltro_ref_pattern.so   : the REFERENCE's own header-only C++ compiled from
                                             /root/reference (oracle/ref_shim.cpp, oracle/Makefile `ref`)
The NumPy preparation around the reference's native calls restates ComputeSparsityPattern.pyx:46-62,80-91.
"""
import ctypes
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
_C = None
_REF = None


def build(ref=True):
    """Compile the C restatement, and the reference shim when /root/reference is present."""
    subprocess.check_call(["make", "-s", "-C", HERE, "all"])
    if ref and os.path.isfile("/root/reference/kiltro/FiniteElements/ComputeSparsityPattern.h"):
        subprocess.check_call(["make", "-s", "-C", HERE, "ref"])


def _p(a, t):
    return a.ctypes.data_as(ctypes.POINTER(t))


def c_lib():
    global _C
    if _C is None:
        path = os.path.join(HERE, "_build", "libkiltro_oracle.so")
        if not os.path.isfile(path):
            build(ref=False)
        _C = ctypes.CDLL(path)
        _C.oracle_sparsity_pattern.restype = ctypes.c_int64
    return _C


def ref_lib():
    """The reference's native code, or None when it was never built (no /root/reference)."""
    global _REF
    if _REF is None:
        path = os.path.join(HERE, "_ref", "libkiltro_ref_pattern.so")
        if not os.path.isfile(path):
            return None
        _REF = ctypes.CDLL(path)
    return _REF


def c_sparsity_pattern(elements, nnodes, nvar=1):
    lib = c_lib()
    el = np.ascontiguousarray(elements, dtype=np.int64)
    ne, npe = el.shape
    cap = (npe * nvar) ** 2 * ne
    indptr = np.zeros(nnodes * nvar + 1, dtype=np.int32)
    indices = np.zeros(max(cap, 1), dtype=np.int32)
    nnz = lib.oracle_sparsity_pattern(_p(el, ctypes.c_int64), ctypes.c_int64(ne), ctypes.c_int32(npe),
                                      ctypes.c_int64(nnodes), ctypes.c_int32(nvar),
                                      _p(indptr, ctypes.c_int32), _p(indices, ctypes.c_int32))
    if nnz < 0:
        raise MemoryError("oracle_sparsity_pattern")
    indices = indices[:nnz].copy()
    dl = np.zeros(max(cap, 1), dtype=np.int32); dg = np.zeros_like(dl)
    lib.oracle_data_indices(_p(el, ctypes.c_int64), ctypes.c_int64(ne), ctypes.c_int32(npe), ctypes.c_int32(nvar),
                            _p(indptr, ctypes.c_int32), _p(indices, ctypes.c_int32),
                            _p(dl, ctypes.c_int32), _p(dg, ctypes.c_int32))
    return indices, indptr, dl[:cap], dg[:cap]


def ref_sparsity_pattern(elements, nnodes, nvar=1):
    """Drives the reference's C++ exactly as its Cython wrapper does (.pyx:46-112)."""
    lib = ref_lib()
    if lib is None:
        raise RuntimeError("oracle/_ref not built")
    el64 = np.ascontiguousarray(elements, dtype=np.int64)
    ne, npe = el64.shape
    sorter = np.ascontiguousarray(np.argsort(el64, axis=1).astype(np.int64))
    srt = np.ascontiguousarray(np.take_along_axis(el64, sorter, axis=1).astype(np.int32))
    flat = srt.ravel()
    idx_sort = np.argsort(flat).astype(np.int32)
    elem_container = np.ascontiguousarray(idx_sort // npe).astype(np.int32)
    idx_start = np.zeros(nnodes + 1, dtype=np.int32)
    idx_start[:-1] = np.unique(flat[idx_sort], return_index=True)[1].astype(np.int32)
    idx_start[-1] = elem_container.shape[0]
    counts = np.zeros(nnodes, dtype=np.int32)
    indices = np.zeros((nvar * npe) ** 2 * ne, dtype=np.int32)
    I = ctypes.c_int
    nnz = lib.ref_compute_sparsity_pattern(_p(flat, I), _p(idx_start, I), _p(elem_container, I),
                                           I(nvar), I(nnodes), I(ne), I(npe), I(nnodes + 1),
                                           _p(counts, I), _p(indices, I))
    counts = np.repeat(counts, nvar)
    indptr = np.zeros(nnodes * nvar + 1, dtype=np.int32)
    indptr[1:] = np.cumsum(np.minimum(counts * nvar, nnodes * nvar))
    indices = np.ascontiguousarray(indices[:nnz])
    cap = (npe * nvar) ** 2 * ne
    dl = np.zeros(cap, dtype=np.int32); dg = np.zeros(cap, dtype=np.int32)
    lib.ref_compute_data_indices(_p(indices, I), _p(indptr, I), I(ne), I(nvar), I(npe), _p(srt, I),
                                 _p(sorter, ctypes.c_long), _p(dl, I), _p(dg, I))
    return indices, indptr, dl, dg
