"""CPU restatement of the entry mapping of the grid assembly kernel (kiltro_b200/csrc/assemble_grid.cu): which entries of
the four element matrices around a node make up its CSR row, in which order, and where boundary rows are shorter.  Checked
against the reference's own scatter (Assembly.py:46,60-61 restated with scipy) on random element matrices."""
import numpy as np
import pytest
import scipy.sparse as sp


def grid_rows(Ke, nxn, nyn):
    """Ke[j, i] = 4x4 matrix of element (i, j), local nodes [BL, BR, TR, TL] (Mesh.py:46-52).  Returns CSR data."""
    nex, ney = nxn - 1, nyn - 1

    def mat(c, j):
        return Ke[j, c] if (0 <= c < nex and 0 <= j < ney) else np.zeros((4, 4))

    out = []
    for j in range(nyn):
        for n in range(nxn):
            c = n - 1                                        # the lane's element column (left of the node)
            p2, b3 = mat(c, j - 1)[2], mat(c + 1, j - 1)[3]  # row TR of the lower-left, row TL of the lower-right element
            m, d0 = mat(c, j).ravel(), mat(c + 1, j)[0]      # upper-left element (row BR = m[4:8]), row BL of the upper-right
            e = [p2[0], p2[1] + b3[0], b3[1],
                 p2[3] + m[4], ((p2[2] + b3[3]) + m[5]) + d0[0], b3[2] + d0[1],
                 m[7], m[6] + d0[3], d0[2]]
            xl, xr, yd, yu = n > 0, n < nxn - 1, j > 0, j < nyn - 1
            keep = [yd and xl, yd, yd and xr, xl, True, xr, yu and xl, yu, yu and xr]
            out += [v for v, k in zip(e, keep) if k]
    return np.array(out)


@pytest.mark.parametrize("nx,ny", [(1, 1), (2, 1), (1, 3), (7, 5), (33, 4)])
def test_grid_entry_mapping_matches_scatter(nx, ny):
    rng = np.random.default_rng(nx * 100 + ny)
    nxn, nyn = nx + 1, ny + 1
    Ke = rng.standard_normal((ny, nx, 4, 4))
    rows, cols, vals = [], [], []
    for j in range(ny):
        for i in range(nx):
            c = j * nxn + i
            nodes = [c, c + 1, c + nxn + 1, c + nxn]
            for a in range(4):
                for b in range(4):
                    rows.append(nodes[a]); cols.append(nodes[b]); vals.append(Ke[j, i, a, b])
    A = sp.coo_matrix((vals, (rows, cols)), shape=(nxn * nyn, nxn * nyn)).tocsr()
    A.sort_indices()
    got = grid_rows(Ke, nxn, nyn)
    assert got.shape[0] == A.nnz == (3 * nxn - 2) * (3 * nyn - 2)
    assert np.allclose(got, A.data, rtol=0, atol=1e-14)


def grid_pattern(nxn, nyn):
    """CPU restatement of kb_pattern_grid (kiltro_b200/csrc/pattern.cu): indptr / indices / diagonal positions of a mesh
    numbered like Mesh.Rectangle, written down from the shape with the kernel's closed-form row offsets."""
    nn = nxn * nyn
    rowlen = 3 * nxn - 2
    indptr = np.zeros(nn + 1, np.int64); indices = []; diag = np.zeros(nn, np.int64)
    for p in range(nn):
        i, j = p % nxn, p // nxn
        cy = (j > 0) + 1 + (j < nyn - 1)
        q = rowlen * (3 * j - 1 if j > 0 else 0) + cy * (3 * i - 1 if i > 0 else 0)
        assert q == len(indices)                             # the closed form IS the running count
        indptr[p] = q
        for dj in range(-1 if j > 0 else 0, (1 if j < nyn - 1 else 0) + 1):
            for di in range(-1 if i > 0 else 0, (1 if i < nxn - 1 else 0) + 1):
                if dj == 0 and di == 0:
                    diag[p] = len(indices)
                indices.append(p + dj * nxn + di)
    indptr[nn] = len(indices)
    return indptr, np.array(indices), diag


@pytest.mark.parametrize("nx,ny", [(1, 1), (2, 1), (1, 4), (7, 5), (33, 4), (20, 31)])
def test_grid_pattern_formula_matches_the_oracle(nx, ny):
    """The shape-only pattern equals the oracle's restatement of the reference's sort + unique (ComputeSparsityPattern.h:22-62),
    and nnz = (3 nxn - 2)(3 nyn - 2)."""
    from oracle import kiltro_oracle as O
    el = O.mesh_rectangle((0.0, 0.0), (1.0, 1.0), nx, ny)[1]
    nxn, nyn = nx + 1, ny + 1
    ind, ptr = O.sparsity_pattern_fast(el, nxn * nyn)
    indptr, indices, diag = grid_pattern(nxn, nyn)
    assert indptr[-1] == (3 * nxn - 2) * (3 * nyn - 2) == len(ind)
    assert np.array_equal(indptr, ptr) and np.array_equal(indices, ind)
    assert np.array_equal(indices[diag], np.arange(nxn * nyn))
