"""CPU tier: the element math of the CUDA kernels (csrc/element.cuh, __host__ __device__) compiled for the host with nvcc
and checked against the reference-generated golden element matrices and the NumPy oracle -- the same formulas the GPU
runs, without a GPU.  Covers the sum-factorised production routine (quad4), the general routine it falls back to
(quad4_general), clockwise elements, non-convex elements (mixed-sign Jacobians) and pure diffusion."""
import ctypes
import os
import shutil
import subprocess

import numpy as np
import pytest

from conftest import Golden, rel_max

HERE = os.path.dirname(os.path.abspath(__file__))
ASM = Golden("asm")
VAL_TOL = 1e-12
MODE = {"reference": 0, "consistent": 1}


@pytest.fixture(scope="module")
def host(tmp_path_factory):
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.isfile(nvcc):
        pytest.skip("nvcc not available")
    so = str(tmp_path_factory.mktemp("element_host") / "element_host.so")
    subprocess.check_call([nvcc, "-O2", "-std=c++17", "-gencode", "arch=compute_100a,code=sm_100a",
                           "--expt-relaxed-constexpr", "-shared", "-Xcompiler", "-fPIC", "-o", so,
                           os.path.join(HERE, "host", "element_host.cu")])
    return ctypes.CDLL(so)


def _p(a):
    return a.ctypes.data_as(ctypes.c_void_p) if a is not None else None


def run_quad4(lib, which, points, elements, velocity, material, mode, petrov, want_mass=True):
    pts = np.ascontiguousarray(points, dtype=np.float64)
    el = np.ascontiguousarray(elements, dtype=np.int32)
    vel = None if velocity is None else np.ascontiguousarray(velocity, dtype=np.float64)
    mat = np.ascontiguousarray(material, dtype=np.float64)
    ne = el.shape[0]
    Ke = np.zeros((ne, 16)); Fe = np.zeros((ne, 4)); Me = np.zeros((ne, 16))
    lib.host_quad4(ctypes.c_int(which), _p(pts), _p(el), ctypes.c_long(ne), _p(vel), _p(mat), ctypes.c_int(MODE[mode]),
                   ctypes.c_int(int(petrov)), ctypes.c_int(int(want_mass)), _p(Ke), _p(Fe), _p(Me))
    return Ke, Fe, Me


def _case_mode(name):
    if name.endswith("_pg"):
        return "reference", True
    if name.endswith("_cons"):
        return "consistent", False
    return "reference", False


@pytest.mark.parametrize("which", [0, 1])
@pytest.mark.parametrize("name", [c for c in ASM.cases() if c.startswith("quad")])
def test_quad4_host_matches_reference_golden(host, name, which):
    c = ASM.case(name)
    mode, petrov = _case_mode(name)
    Ke, Fe, Me = run_quad4(host, which, c["points"], c["elements"], c["velocity"], c["material"], mode, petrov)
    assert rel_max(Ke, c["Ke"]) < VAL_TOL and rel_max(Fe, c["Fe"]) < VAL_TOL and rel_max(Me, c["Me"]) < VAL_TOL


def _random_quads(seed, ne, kind):
    rng = np.random.default_rng(seed)
    pts = np.zeros((4 * ne, 2)); el = np.arange(4 * ne).reshape(ne, 4)
    base = np.array([[0.0, 0.0], [1.0, 0.0], [1.0, 1.0], [0.0, 1.0]])
    for e in range(ne):
        A = np.eye(2) + 0.3 * rng.standard_normal((2, 2))
        q = base @ A.T * rng.uniform(0.01, 3.0) + rng.standard_normal(2)
        if not kind.startswith("affine"):
            q = q + 0.12 * rng.standard_normal((4, 2))
        elif e % 3 == 0:                               # axis-parallel rectangles (what Mesh.Rectangle generates)
            q = base * np.array([rng.uniform(0.01, 2.0), rng.uniform(0.01, 2.0)]) + rng.standard_normal(2)
        if kind in ("clockwise", "affine_cw"):
            q = q[[0, 3, 2, 1]]
        elif kind == "nonconvex":                      # pull node 2 through the diagonal: det changes sign inside
            q = base.copy(); q[2] = [0.3, 0.3] + 0.05 * rng.standard_normal(2)
        pts[4 * e:4 * e + 4] = q
    vel = rng.standard_normal((4 * ne, 2)) * rng.uniform(0.1, 50.0)
    return pts, el, vel


@pytest.mark.parametrize("kind", ["ccw", "clockwise", "nonconvex", "affine", "affine_cw"])
@pytest.mark.parametrize("mode,petrov", [("reference", False), ("reference", True), ("consistent", False),
                                         ("consistent", True)])
def test_quad4_host_matches_oracle_random(host, kind, mode, petrov):
    from oracle import kiltro_oracle as O
    pts, el, vel = _random_quads(11, 64, kind)
    mat = (2.5, 1.3, 0.7, 1.7, 0.9)
    rKe, rFe, rMe = O.element_matrices(pts, el, vel, mat, mode=mode, petrov=petrov, want_mass=True)
    for which in (0, 1):
        Ke, Fe, Me = run_quad4(host, which, pts, el, vel, mat, mode, petrov)
        # per element: relative to that element's largest entry
        for e in range(el.shape[0]):
            assert rel_max(Ke[e], rKe[e]) < VAL_TOL, (which, e)
        assert rel_max(Fe, rFe) < VAL_TOL and rel_max(Me, rMe) < VAL_TOL


def test_quad4_host_pure_diffusion_and_zero_source(host):
    from oracle import kiltro_oracle as O
    pts, el, _ = _random_quads(5, 32, "ccw")
    mat = (3.0, 1.0, 1.0, 2.0, 0.0)
    rKe, rFe, _ = O.element_matrices(pts, el, None, mat, mode="consistent", petrov=True, want_mass=False)
    Ke, Fe, Me = run_quad4(host, 0, pts, el, None, mat, "consistent", True, want_mass=False)
    assert rel_max(Ke, rKe) < VAL_TOL and np.all(Fe == 0.0) and np.all(Me == 0.0)
    assert np.abs(Ke.reshape(-1, 4, 4).sum(axis=2)).max() < 1e-12 * np.abs(Ke).max()      # constants in the null space
    assert np.abs(Ke.reshape(-1, 4, 4) - Ke.reshape(-1, 4, 4).transpose(0, 2, 1)).max() == 0.0


@pytest.mark.parametrize("name", [c for c in ASM.cases() if c.startswith("line")])
def test_line2_host_matches_reference_golden(host, name):
    c = ASM.case(name)
    mode, petrov = _case_mode(name)
    pts = np.ascontiguousarray(c["points"], dtype=np.float64).ravel()
    el = np.ascontiguousarray(c["elements"], dtype=np.int32)
    vel = np.ascontiguousarray(c["velocity"], dtype=np.float64).ravel()
    mat = np.ascontiguousarray(c["material"], dtype=np.float64)
    ne = el.shape[0]
    Ke = np.zeros((ne, 4)); Fe = np.zeros((ne, 2)); Me = np.zeros((ne, 4))
    host.host_line2(_p(pts), _p(el), ctypes.c_long(ne), _p(vel), _p(mat), ctypes.c_int(MODE[mode]),
                    ctypes.c_int(int(petrov)), ctypes.c_int(1), _p(Ke), _p(Fe), _p(Me))
    assert rel_max(Ke, c["Ke"]) < VAL_TOL and rel_max(Fe, c["Fe"]) < VAL_TOL and rel_max(Me, c["Me"]) < VAL_TOL


NEXT = Golden("next")


@pytest.mark.parametrize("name", NEXT.cases())
def test_operator_routines_host_match_reference_golden(host, name):
    """quad4_robin / quad4_chargalerkin / line2_*  compiled for the host, assembled with the oracle's
    scatter, against the dense matrices of the reference's AssemblyRobinForces / AssembleCharacteristicGalerkin."""
    import scipy.sparse as sp
    from oracle import kiltro_oracle as O
    c = NEXT.case(name)
    el = np.ascontiguousarray(c["elements"], dtype=np.int32)
    ne, npe = el.shape
    pts = np.ascontiguousarray(c["points"], dtype=np.float64).ravel()
    vel = np.ascontiguousarray(c["velocity"], dtype=np.float64).ravel()
    nn = c["points"].shape[0]
    k, rho, cv, area, q = c["material"]
    for op, p0, p1, Kref, Fref in ((1, float(c["robin_h"]), float(c["robin_tinf"]), c["K_robin"], c["F_robin"]),
                                   (2, q * area, 0.0, c["K_cg"], c["F_cg"])):
        Ke = np.zeros((ne, npe * npe)); Fe = np.zeros((ne, npe))
        host.host_operator(ctypes.c_int(npe), ctypes.c_int(op), _p(pts), _p(el), ctypes.c_long(ne), _p(vel),
                           ctypes.c_double(p0), ctypes.c_double(p1), _p(Ke), _p(Fe))
        A = O.assemble_values(c["elements"], nn, Ke, Fe)
        Kd = sp.csr_matrix((A["data"], A["indices"], A["indptr"]), shape=(nn, nn)).toarray()
        assert rel_max(Kd, Kref) < VAL_TOL and rel_max(A["F"], Fref) < VAL_TOL, op
