"""Host-side pieces of the experimental multigrid preconditioner (kiltro_b200/multigrid.py): the transfer operators."""
import numpy as np
import pytest

from kiltro_b200.multigrid import GridMultigrid, _interp_1d


@pytest.mark.parametrize("nf,nc", [(9, 5), (129, 65), (100, 51), (4000, 2001), (7, 4)])
def test_interp_1d_reproduces_linear_functions(nf, nc):
    P = _interp_1d(nf, nc)
    assert P.shape == (nf, nc) and P.data.min() > 0.0
    assert np.abs(np.asarray(P.sum(axis=1)).ravel() - 1.0).max() < 1e-14            # partition of unity
    xc, xf = np.linspace(0.0, 1.0, nc), np.linspace(0.0, 1.0, nf)
    assert np.abs(P @ (3.0 * xc - 1.0) - (3.0 * xf - 1.0)).max() < 1e-13             # exact for linear functions
    assert P[0, 0] == 1.0 and P[nf - 1, nc - 1] == 1.0                                # end points coincide
    if nf == 2 * nc - 1:                                                              # nested: the classical stencil
        assert np.array_equal(P[2:nf:2].toarray(), np.eye(nc)[1:])
        assert np.allclose(P[1].toarray().ravel()[:2], [0.5, 0.5])


def test_transfer_restricted_to_free_dofs():
    nxf, nyf, nxc, nyc = 9, 7, 5, 4
    ff = np.ones((nyf, nxf), bool); ff[:, 0] = False; ff[:, -1] = False
    cf = np.ones((nyc, nxc), bool); cf[:, 0] = False; cf[:, -1] = False
    P = GridMultigrid._transfer((nxf, nyf), ff, (nxc, nyc), cf)
    assert P.shape == (int(ff.sum()), int(cf.sum()))
    # a field that is bilinear AND vanishes on the removed (Dirichlet) columns is interpolated exactly
    X, Y = np.meshgrid(np.linspace(0, 1, nxc), np.linspace(0, 1, nyc))
    Xf, Yf = np.meshgrid(np.linspace(0, 1, nxf), np.linspace(0, 1, nyf))
    # piecewise-bilinear hat in x that is zero on x = 0 and x = 1 at coarse nodes: use the coarse nodal values directly
    uc = (np.minimum(X, 1 - X) * (1.0 + 2.0 * Y))
    uf = P @ uc[cf]
    # compare with full (unrestricted) interpolation of the same coarse field
    import scipy.sparse as sp
    full = sp.kron(_interp_1d(nyf, nyc), _interp_1d(nxf, nxc), format="csr") @ uc.ravel()
    assert np.abs(uf - full.reshape(nyf, nxf)[ff]).max() < 1e-14
