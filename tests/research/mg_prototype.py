#!/usr/bin/env python
"""CPU feasibility prototype for DESIGN 6c item 1 (NOT product code, NOT the oracle's job -- it only borrows the oracle's
assembly to build the S16-style test matrix): Jacobi-BiCGSTAB, as prescribed by north_star, against BiCGSTAB
preconditioned by one geometric-multigrid V-cycle (bilinear transfer on the Mesh.Rectangle hierarchy, Galerkin coarse
operators R A P, damped-Jacobi smoothing = the SpMV the B200 path already has).  Counts operator applications.

    python tests/research/mg_prototype.py [elements_per_side = 256]
"""
import os
import sys
import time

import numpy as np
import scipy.sparse as sp
import scipy.sparse.linalg as spla

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", ".."))
from oracle import kiltro_oracle as O


def problem(ne, peclet=0.45, seed=0):
    nps = ne + 1
    pts, el = O.mesh_rectangle((0.0, 0.0), (1.0, 1.0), ne, ne)
    h = 1.0 / ne
    U = peclet * 2.0 / h
    rng = np.random.default_rng(seed)
    vel = np.zeros(pts.shape); vel[:, 0] = U; vel += 0.1 * U * rng.standard_normal(pts.shape)
    d = np.full((pts.shape[0], 1), np.nan); d[::nps] = 100.0; d[nps - 1::nps] = 500.0
    A = O.assemble(pts, el, vel, (1.0, 1.0, 1.0, 1.0, 0.0), "consistent", False, False)
    nn = pts.shape[0]
    cin, cout, td = O.dirichlet_split(d)
    kb = O.apply_dirichlet_reduce(A["indices"], A["indptr"], A["K"], np.zeros(nn), cin, cout, td)
    Kb = sp.csr_matrix((kb[0], kb[1], kb[2]), shape=(len(cin), len(cin)))
    return Kb, kb[3], nps, cin


def interp_1d(nc):
    """Linear interpolation from nc coarse nodes to 2 nc - 1 fine nodes."""
    nf = 2 * nc - 1
    rows, cols, vals = [], [], []
    for i in range(nc):
        rows.append(2 * i); cols.append(i); vals.append(1.0)
        if i + 1 < nc:
            rows += [2 * i + 1, 2 * i + 1]; cols += [i, i + 1]; vals += [0.5, 0.5]
    return sp.csr_matrix((vals, (rows, cols)), shape=(nf, nc))


def hierarchy(A, nps, free_mask, levels):
    """Galerkin hierarchy on the full node grid restricted to free DOFs (Dirichlet columns x = 0, 1 removed)."""
    ops, Ps, smooth = [A], [], []
    n, mask = nps, free_mask
    for _ in range(levels):
        if (n - 1) % 2 or n < 9:
            break
        nc = (n + 1) // 2
        P1 = interp_1d(nc)
        P = sp.kron(P1, P1, format="csr")                       # y-major node numbering, x fastest
        maskc = np.ones((nc, nc), bool); maskc[:, 0] = False; maskc[:, -1] = False
        P = P[mask.ravel()][:, maskc.ravel()]
        Ac = (P.T @ ops[-1] @ P).tocsr()
        Ps.append(P); ops.append(Ac)
        n, mask = nc, maskc
    return ops, Ps


def hierarchy_rediscretized(ne, peclet=0.45, seed=0, petrov_coarse=True):
    """Every level assembled from its own Mesh.Rectangle with the nodal velocities injected; the coarse levels use the
    Petrov-Galerkin (streamline-upwind) weights the path already has (TemperaturePG / HeatAdvectionDiffusionPG,
    VariationalPrinciple.py:216-261,311-342) because their cell Peclet number doubles per level.  Returns (ops, Ps, b)."""
    nps = ne + 1
    h = 1.0 / ne
    U = peclet * 2.0 / h
    rng = np.random.default_rng(seed)
    vel = np.zeros((nps * nps, 2)); vel[:, 0] = U; vel += 0.1 * U * rng.standard_normal(vel.shape)
    v = vel.reshape(nps, nps, 2)
    ops, Ps, b, n, lvl = [], [], None, nps, 0
    while True:
        p, e = O.mesh_rectangle((0.0, 0.0), (1.0, 1.0), n - 1, n - 1)
        A = O.assemble(p, e, v.reshape(-1, 2), (1.0, 1.0, 1.0, 1.0, 0.0), "consistent", lvl > 0 and petrov_coarse, False)
        d = np.full((n * n, 1), np.nan); d[::n] = 100.0; d[n - 1::n] = 500.0
        cin, cout, td = O.dirichlet_split(d)
        kb = O.apply_dirichlet_reduce(A["indices"], A["indptr"], A["K"], np.zeros(n * n), cin, cout, td)
        ops.append(sp.csr_matrix((kb[0], kb[1], kb[2]), shape=(len(cin), len(cin))))
        if lvl == 0:
            b = kb[3]
        if (n - 1) % 2 or n < 9:
            break
        nc = (n + 1) // 2
        P1 = interp_1d(nc)
        mf = np.ones((n, n), bool); mf[:, 0] = False; mf[:, -1] = False
        mc = np.ones((nc, nc), bool); mc[:, 0] = False; mc[:, -1] = False
        Ps.append(sp.kron(P1, P1, format="csr")[mf.ravel()][:, mc.ravel()])
        v, n, lvl = v[::2, ::2], nc, lvl + 1
    return ops, Ps, b


def vcycle(ops, Ps, b, lvl=0, nu=2, omega=0.8, counter=None):
    A = ops[lvl]
    if lvl == len(ops) - 1:
        return spla.spsolve(A.tocsc(), b)
    Dinv = 1.0 / A.diagonal()
    x = np.zeros_like(b)
    for _ in range(nu):
        x += omega * Dinv * (b - A @ x)
        if counter is not None: counter[lvl] += 1
    r = b - A @ x
    if counter is not None: counter[lvl] += 1
    x += Ps[lvl] @ vcycle(ops, Ps, Ps[lvl].T @ r, lvl + 1, nu, omega, counter)
    for _ in range(nu):
        x += omega * Dinv * (b - A @ x)
        if counter is not None: counter[lvl] += 1
    return x


def run(ne):
    t0 = time.time()
    A, b, nps, cin = problem(ne)
    n = A.shape[0]
    free = np.ones((nps, nps), bool); free[:, 0] = False; free[:, -1] = False
    xref = spla.spsolve(A.tocsc(), b)
    res = {"ndof": n}
    # north_star's method: right-Jacobi BiCGSTAB
    Dinv = 1.0 / A.diagonal()
    it = [0]
    Aj = spla.LinearOperator((n, n), matvec=lambda v: (it.__setitem__(0, it[0] + 1), A @ (Dinv * v))[1])
    y, info = spla.bicgstab(Aj, b, rtol=1e-10, atol=0.0, maxiter=50000)
    res["jacobi_bicgstab"] = {"spmv": it[0], "info": info,
                              "rel_err": float(np.linalg.norm(Dinv * y - xref) / np.linalg.norm(xref))}
    # multigrid-preconditioned BiCGSTAB.  Galerkin coarse operators R A P of the central (Galerkin) fine operator DIVERGE
    # for cell Pe 0.45 (the coarse cell Peclet number passes 1 and damped Jacobi amplifies): `galerkin_rap` reports that;
    # the rediscretised Petrov-Galerkin hierarchy is the one that works.
    ops_g, Ps_g = hierarchy(A, nps, free, 12)
    xg = np.zeros_like(b)
    for _ in range(4):
        xg += vcycle(ops_g, Ps_g, b - A @ xg)
    res["galerkin_rap_residual_after_4_vcycles"] = float(np.linalg.norm(b - A @ xg) / np.linalg.norm(b))
    ops, Ps, b2 = hierarchy_rediscretized(ne)
    assert np.array_equal(b2, b) and abs(ops[0] - A).max() == 0.0
    counter = [0] * len(ops)
    it = [0]
    Am = spla.LinearOperator((n, n), matvec=lambda v: (it.__setitem__(0, it[0] + 1), A @ v)[1])
    M = spla.LinearOperator((n, n), matvec=lambda v: vcycle(ops, Ps, v, counter=counter))
    x, info = spla.bicgstab(Am, b, rtol=1e-10, atol=0.0, maxiter=500, M=M)
    work = it[0] + sum(c * ops[l].nnz / ops[0].nnz for l, c in enumerate(counter))
    res["gmg_bicgstab"] = {"levels": len(ops), "outer_spmv": it[0], "smoother_spmv_per_level": counter,
                           "fine_spmv_equivalents": round(work, 1), "info": info,
                           "rel_err": float(np.linalg.norm(x - xref) / np.linalg.norm(xref)),
                           "coarse_nnz_ratio": [round(o.nnz / ops[0].nnz, 4) for o in ops]}
    res["seconds"] = round(time.time() - t0, 1)
    return res


if __name__ == "__main__":
    import json
    ne = int(sys.argv[1]) if len(sys.argv) > 1 else 256
    print(json.dumps(run(ne)))
