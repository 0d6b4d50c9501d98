"""CPU tier: PostProcess.WriteVTK (kiltro/PostProcess.py:26-76) writes pyevtk-layout .vtu files without
pyevtk.  The files are parsed back here (XML header + raw appended block) and compared with what went in."""
import os
import re
import struct

import numpy as np
import pytest


def read_vtu(path):
    raw = open(path, "rb").read()
    cut = raw.index(b'<AppendedData encoding="raw">')
    xml = raw[:cut].decode("ascii")
    blob = raw[raw.index(b"_", cut) + 1:]
    assert 'byte_order="LittleEndian"' in xml and 'header_type="UInt64"' in xml and 'type="UnstructuredGrid"' in xml
    npts, ncell = [int(v) for v in re.search(r'NumberOfPoints="(\d+)" NumberOfCells="(\d+)"', xml).groups()]
    out = {"npts": npts, "ncell": ncell}
    dt = {"Float64": "<f8", "Int64": "<i8", "UInt8": "u1"}
    for name, ncomp, typ, off in re.findall(r'<DataArray Name="(\w+)" NumberOfComponents="(\d+)" type="(\w+)" format="appended" offset="(\d+)"/>', xml):
        off = int(off)
        nbytes = struct.unpack("<Q", blob[off:off + 8])[0]
        a = np.frombuffer(blob[off + 8:off + 8 + nbytes], dtype=dt[typ])
        out[name] = a.reshape(-1, int(ncomp)) if int(ncomp) > 1 else a
    assert raw.rstrip().endswith(b"</VTKFile>")
    return out


@pytest.mark.parametrize("kind", ["quad", "line"])
def test_write_vtk_roundtrip(tmp_path, kind):
    from kiltro_b200.MeshGeneration import Mesh
    from kiltro_b200.PostProcess import PostProcess
    rng = np.random.default_rng(0)
    m = Mesh(element_type=kind)
    if kind == "quad":
        nx, ny = 4, 3
        x, y = np.meshgrid(np.linspace(0, 2, nx + 1), np.linspace(0, 1, ny + 1))
        m.points = np.column_stack([x.ravel(), y.ravel()])
        c = np.array([j * (nx + 1) + i for j in range(ny) for i in range(nx)])
        m.elements = np.column_stack([c, c + 1, c + nx + 2, c + nx + 1])
    else:
        m.points = np.linspace(0.0, 1.0, 8)[:, None]
        m.elements = np.column_stack([np.arange(7), np.arange(1, 8)])
    sol = rng.standard_normal((m.nnodes, 1, 3))
    pp = PostProcess(m.ndim, 1); pp.SetMesh(m); pp.SetSolution(sol)
    files = pp.WriteVTK(str(tmp_path / "out.vtu"))
    # naming convention of the reference: <base>_quantity_<q>_increment_<i>.vtu  (PostProcess.py:68)
    assert [os.path.basename(f) for f in files] == ["out_quantity_0_increment_%d.vtu" % i for i in range(3)]
    for i, f in enumerate(files):
        d = read_vtu(f)
        assert d["npts"] == m.nnodes and d["ncell"] == m.nelem
        assert np.array_equal(d["points"][:, :m.ndim], m.points) and np.all(d["points"][:, m.ndim:] == 0.0)
        assert np.array_equal(d["connectivity"], m.elements.ravel())
        npe = m.elements.shape[1]
        assert np.array_equal(d["offsets"], np.arange(1, m.nelem + 1) * npe)
        assert np.all(d["types"] == (9 if kind == "quad" else 3))
        assert np.array_equal(d["T"], sol[:, 0, i])


def test_write_vtk_errors(tmp_path):
    from kiltro_b200.MeshGeneration import Mesh
    from kiltro_b200.PostProcess import PostProcess
    pp = PostProcess(2, 1)
    with pytest.raises(ValueError):
        pp.WriteVTK(str(tmp_path / "x.vtu"))
    m = Mesh(element_type="tri"); m.points = np.zeros((3, 2)); m.elements = np.array([[0, 1, 2]])
    pp.SetMesh(m); pp.SetSolution(np.zeros((3, 1, 1)))
    with pytest.raises(ValueError):
        pp.WriteVTK(str(tmp_path / "x.vtu"))
