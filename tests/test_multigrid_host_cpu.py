"""Host-side pieces of the opt-in multigrid preconditioner (csrc/multigrid.cu, kiltro_b200/multigrid.py): the level plan
and the 1-D transfer tables (pure host arithmetic of the C-ABI library: runs without a GPU)."""
import ctypes

import numpy as np
import pytest

from kiltro_b200 import _lib
from kiltro_b200.multigrid import GridMultigrid


def _dense_P(nf, nc):
    j, w, start = GridMultigrid.transfer_table(nf, nc)
    P = np.zeros((nf, nc))
    for i in range(nf):
        P[i, j[i]] += 1.0 - float(w[i])
        P[i, j[i] + 1] += float(w[i])
    return P, j, w, start


@pytest.mark.parametrize("nf,nc", [(9, 5), (129, 65), (100, 51), (4000, 2001), (7, 4), (5, 5), (3, 3), (2, 2)])
def test_transfer_table_reproduces_linear_functions(nf, nc):
    P, j, w, start = _dense_P(nf, nc)
    assert j.min() >= 0 and j.max() <= nc - 2 and w.min() >= 0.0 and w.max() <= 1.0
    assert np.abs(P.sum(axis=1) - 1.0).max() < 1e-6                                   # partition of unity (fp32 weights)
    xc, xf = np.linspace(0.0, 1.0, nc), np.linspace(0.0, 1.0, nf)
    assert np.abs(P @ (3.0 * xc - 1.0) - (3.0 * xf - 1.0)).max() < 1e-6               # exact for linear functions
    assert P[0, 0] == 1.0 and P[nf - 1, nc - 1] == 1.0                                # end points coincide
    if nf == 2 * nc - 1 and nc > 2:                                                   # nested: the classical stencil
        assert np.array_equal(P[2:nf:2], np.eye(nc)[1:])
        assert np.allclose(P[1, :2], [0.5, 0.5])
    # independent restatement: floor / fractional part of i (nc-1)/(nf-1)
    s = np.arange(nf) * (nc - 1) / (nf - 1)
    jj = np.minimum(np.floor(s + 1e-12).astype(int), nc - 2)
    assert np.array_equal(j, jj) and np.abs(w - (s - jj)).max() < 1e-6


@pytest.mark.parametrize("nf,nc", [(9, 5), (100, 51), (7, 4), (4000, 2001), (5, 5)])
def test_transfer_table_start_gives_the_transpose(nf, nc):
    """The restriction kernel walks fine nodes [start[c-1], start[c+1]) for coarse node c with weight (1-w) where j == c and
    w where j == c-1: that must be exactly column c of P."""
    P, j, w, start = _dense_P(nf, nc)
    assert start[0] == 0 and start[nc] == nf and np.all(np.diff(start) >= 0)
    for c in range(nc):
        col = np.zeros(nf)
        for i in range(start[max(c - 1, 0)], start[min(c + 1, nc)]):
            col[i] = (1.0 - float(w[i])) if j[i] == c else float(w[i])
            assert j[i] in (c - 1, c)
        assert np.array_equal(col, P[:, c])


@pytest.mark.parametrize("nx,ny,coarsest", [(4000, 4000, 300), (129, 65, 300), (4000, 32000, 300), (10, 500, 64), (3, 3, 4)])
def test_level_plan(nx, ny, coarsest):
    lib = _lib.load()
    h = ctypes.c_void_p()
    _lib.check(lib.kb_mg_create(nx, ny, coarsest, ctypes.byref(h)))
    try:
        shapes = []
        for l in range(lib.kb_mg_num_levels(h)):
            a, b = _lib.I64(), _lib.I64()
            _lib.check(lib.kb_mg_level_shape(h, l, ctypes.byref(a), ctypes.byref(b)))
            shapes.append((a.value, b.value))
        assert shapes[0] == (nx, ny)
        for (fx, fy), (cx, cy) in zip(shapes, shapes[1:]):
            assert cx == (fx // 2 + 1 if fx > 3 else fx) and cy == (fy // 2 + 1 if fy > 3 else fy)
            assert (cx, cy) != (fx, fy)
        assert shapes[-1][0] * shapes[-1][1] <= max(coarsest, 9)
        assert lib.kb_mg_workspace_bytes(h) > 40 * nx * ny            # 8 coefficients + diagonal + 4 vectors in fp32, level 0
    finally:
        lib.kb_mg_destroy(h)


def test_bad_arguments():
    lib = _lib.load()
    h = ctypes.c_void_p()
    assert lib.kb_mg_create(1, 5, 300, ctypes.byref(h)) == -1
    assert lib.kb_mg_transfer_table(5, 9, None, None, None) == -1
