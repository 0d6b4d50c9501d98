"""CPU-side checks (no GPU, no compute): the C-ABI library loads and exports every symbol include/kiltro_b200.h declares,
argument validation works, host-side mirrors of the reference interface behave, and the product path refuses to run
without a CUDA device instead of falling back to anything."""
import ctypes
import os
import re

import numpy as np
import pytest

from conftest import ROOT, Golden

HEADER = os.path.join(ROOT, "include", "kiltro_b200.h")


@pytest.fixture(scope="module")
def lib():
    from kiltro_b200 import build, _lib
    build.build_library()
    return _lib.load()


def declared_functions():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(kb_[a-z0-9_]+)\s*\(", src)))


def test_header_symbols_exported(lib):
    from kiltro_b200 import _lib
    names = declared_functions()
    assert len(names) >= 25
    for n in names:
        assert hasattr(lib, n), "library does not export %s" % n
        assert n in _lib.SIGNATURES, "ctypes binding missing for %s" % n
    for n in _lib.SIGNATURES:
        assert n in names, "binding %s is not declared in the header" % n


def test_version_and_error_strings(lib):
    assert lib.kb_version() >= 100
    assert lib.kb_error_string(0) == b"success"
    assert b"invalid" in lib.kb_error_string(-1)
    assert b"int32" in lib.kb_error_string(-2)
    assert lib.kb_error_string(-5) and lib.kb_error_string(2)      # a cudaError_t decodes too


def test_argument_validation_without_gpu(lib):
    # every call below must be rejected before touching the device
    assert lib.kb_mesh_line(0.0, 1.0, 0, None, None, None) == -1
    assert lib.kb_mesh_rectangle(0.0, 0.0, 1.0, 1.0, 4, 4, None, None, None) == -1
    assert lib.kb_mesh_rectangle(0.0, 0.0, 1.0, 1.0, 70000, 70000, 16, 16, None) == -2      # > int32 nodes
    assert lib.kb_element_matrices(3, None, None, 1, None, None, 0, 0, None, None, None, None) == -1
    assert lib.kb_spmv(4, 4, None, None, None, None, None, None, None) == -1
    from kiltro_b200 import _lib
    info = _lib.kb_solve_info()
    assert lib.kb_cg_solve(4, 4, None, None, None, None, None, 1e-8, 10, 1, None, 0, info, None) == -1
    # entry points added with the template assembly, the row-pattern SpMV and the SURVEY 8(f) operators
    mat = _lib.kb_material()
    assert lib.kb_assemble_fused_tmpl(None, 4, None, mat, 0, 0, None, 1, None, None, None, None, None, None, None, None,
                                      None, None, None) == -1
    i3 = (ctypes.c_int64 * 3)()
    assert lib.kb_tile_tmpl_build(4, 16, 1, None, None, None, None, None, None, None, None, None, None, i3, None) == -1
    assert lib.kb_element_operator(2, 7, None, None, 1, None, 1.0, 0.0, None, None, None) == -1      # NULLs / unknown operator
    assert lib.kb_axpby(4, 1.0, None, 1.0, None, None, None) == -1
    assert lib.kb_make_row_patterns(4, None, None, None, None, None, None, None) == -1
    assert lib.kb_spmv_dots(4, 4, None, None, None, None, None, None, None, None, None, None, None, None, None) == -1


def test_host_side_size_queries(lib):
    assert lib.kb_spmv_num_blocks(0) == 1
    assert lib.kb_spmv_num_blocks(2304) == 1 and lib.kb_spmv_num_blocks(2305) == 2
    n, nnz = 16_000_000, 143_952_004
    ws = lib.kb_krylov_workspace_bytes(n, nnz)
    assert 10 * nnz + 7 * 8 * n < ws < 10 * nnz + 7 * 8 * n + (64 << 20)   # scaled values + 16-bit offsets + 7 vectors
    assert lib.kb_transient_workspace_bytes(n, nnz) > ws
    sizes = (ctypes.c_int64 * 4)()
    assert lib.kb_tile_plan_sizes(4000 * 4000, 3999 * 3999, 4, 4000, 4000, sizes) == 0
    assert sizes[0] == 267 * 267 and sizes[3] == 1                 # 15 x 15 node patches
    assert lib.kb_tile_plan_sizes(1001, 1000, 2, 0, 0, sizes) == 0 and sizes[3] == 0 and sizes[0] == 5
    t3 = (ctypes.c_int64 * 3)()
    assert lib.kb_tile_tmpl_sizes(t3) == 0
    assert t3[0] % 16 == 0 and 20_000 < t3[0] < 32_768 and t3[1] >= 9 and t3[2] >= 15    # one template fits shared memory


def test_no_cpu_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    import kiltro_b200 as K
    from kiltro_b200._lib import KiltroB200Error
    m = K.Mesh(element_type="quad")
    with pytest.raises(KiltroB200Error):
        m.Rectangle((0, 0), (1, 1), 2, 2)
    m.points = np.zeros((4, 2)); m.elements = np.array([[0, 1, 3, 2]])
    with pytest.raises(KiltroB200Error):
        K.ComputeSparsityPattern(m, 1)


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "kiltro_b200")
    for dp, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh")):
                src = open(os.path.join(dp, f)).read()
                assert "import oracle" not in src and "from oracle" not in src and "oracle." not in src, f
    # development scripts and examples run the product only; whatever needs the oracle lives under tests/
    for sub in ("scripts", "examples"):
        for dp, _, files in os.walk(os.path.join(ROOT, sub)):
            for f in files:
                if f.endswith((".py", ".sh")):
                    src = open(os.path.join(dp, f)).read()
                    assert "import oracle" not in src and "from oracle" not in src, f
    # bench.py touches it only inside its CPU legs (cpu_baseline / --impl reference)
    src = open(os.path.join(ROOT, "bench.py")).read()
    assert src.count("from oracle") == 1 and "def cpu_oracle_run" in "".join(src.split("from oracle")[0].splitlines()[-6:])


def test_function_space_tables_match_reference():
    import kiltro_b200 as K
    fs = Golden("fs")

    class M1: ndim = 1
    class M2: ndim = 2
    for kind, mm in (("line", M1), ("quad", M2)):
        c = fs.case(kind)
        f = K.FunctionSpace(mm)
        assert np.array_equal(f.Bases, c["Bases"]) and np.array_equal(f.gBases, c["gBases"])
        assert np.array_equal(f.AllGauss, c["AllGauss"])


def test_material_and_solver_interface():
    import kiltro_b200 as K
    mat = K.FourierConduction(2, k=1.0e3, area=1.0, density=1.0, c_v=1.0, q=0.0)       # examples/steady_2d.py:40
    assert (mat.k, mat.rho, mat.c_v, mat.area, mat.q, mat.nvar) == (1.0e3, 1.0, 1.0, 1.0, 0.0, 1)
    mat = K.FourierConduction(1, k=10.0, area=1.0, density=10.0e6, heat_capacity_v=1.0)  # examples/transient_1d.py
    assert mat.q == 0.0 and mat.c_v == 1.0
    s = mat.as_struct()
    assert (s.k, s.rho, s.c_v, s.area, s.q) == (10.0, 10.0e6, 1.0, 1.0, 0.0)
    with pytest.raises(ValueError):
        K.FourierConduction(2, area=1.0)
    fs = K.FEMSolver(analysis_type="transient", number_of_increments=61)               # examples/transient_2d.py:73
    assert fs.number_of_increments == 61 and fs.analysis_nature == "linear" and fs.number_of_load_increments == 1
    with pytest.raises(ValueError):
        fs.Solve()                                                                      # FEMSolver.py:37-38
    from kiltro_b200.transient import snapshot_indices
    assert snapshot_indices(61) == [0, 1, 20, 40, 60] and snapshot_indices(151) == [0, 1, 50, 100, 150]
    ls = K.LinearSolver(linear_solver="direct", linear_solver_type="lapack")           # the reference's defaults
    with pytest.raises(ValueError):
        ls.Solve(np.eye(3), np.ones(3))                                                 # FEMSolver.py:355-356


def test_boundary_condition_setters():
    import kiltro_b200 as K
    bc = K.BoundaryCondition()
    d = np.full((5, 1), np.nan); d[0] = 1.0
    assert bc.SetDirichletCriteria(lambda a: a, d) is d and bc.dirichlet_flags is d
    n = np.full((5, 1), np.nan)
    assert bc.SetNeumannCriteria(lambda: n) is n
    t = bc.SetNeumannCriteria(lambda: (n, d))
    assert bc.neumann_flags is n and bc.applied_neumann is d and t == (n, d)
    with pytest.raises(ValueError):
        bc.SetRobinCriteria(lambda: {"type": "Spring", "flags": n, "data": 1.0})
    bc.SetRobinCriteria(lambda: {"type": "Convection", "flags": n, "data": 2.0})
    assert bc.applied_robin_convection == 2.0
    bc2 = K.BoundaryCondition()
    with pytest.raises(RuntimeError):
        class F: nvar = 1; ndim = 1
        class Me: nnodes = 5
        bc2.GetDirichletBoundaryConditions(F, Me)
