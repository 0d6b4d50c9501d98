"""Result holder, mirroring kiltro/PostProcess.py:6-23.  `.sol` is (nnodes, nvar, nincr) float64 like the reference's; it
is kept in HBM (`.sol_device`) and copied to the host on first access.  WriteVTK (PostProcess.py:26-76)
writes the same files the reference hands to pyevtk -- one VTK XML UnstructuredGrid (.vtu, raw appended data) per quantity
and increment -- without pyevtk; fields resident on the device are streamed to the file in slices."""
import os
import struct
import sys
from warnings import warn

import numpy as np

from . import device as D


class PostProcess(object):

    def __init__(self, ndim, nvar):
        self.ndim = ndim
        self.nvar = nvar
        self.mesh = None
        self._sol_h = None
        self.sol_device = None
        self.solver_info = None
        self.converged = True          # False when an iterative solve behind this result did not meet its tolerance

    def SetMesh(self, mesh):
        self.mesh = mesh

    def SetSolution(self, sol):
        if isinstance(sol, np.ndarray):
            self._sol_h, self.sol_device = sol, None
        else:
            self._sol_h, self.sol_device = None, sol

    @property
    def sol(self):
        if self._sol_h is None and self.sol_device is not None:
            self._sol_h = D.to_numpy(self.sol_device)
        return self._sol_h

    # ---------------------------------------------------------------------------------- PostProcess.py:26-76
    def WriteVTK(self, filename=None, quantity=0):
        """Writes results to VTK files for Paraview: <base>_quantity_<q>_increment_<i>.vtu, point field "T"."""
        if self.mesh is None or (self._sol_h is None and self.sol_device is None):
            raise ValueError("PostProcess needs a mesh and a solution before WriteVTK")
        if filename is None:
            warn("file name not specified. I am going to write in the current directory")       # :36-38
            filename = os.path.dirname(os.path.realpath(sys.argv[0])) + "/output.vtu"
        mesh = self.mesh
        if mesh.element_type == 'line':                          # :44-51  (VTK_LINE / VTK_QUAD)
            cellflag, offset = 3, 2
        elif mesh.element_type == 'quad':
            cellflag, offset = 9, 4
        else:
            raise ValueError("element type %r cannot be written" % (mesh.element_type,))
        pts = mesh.points
        points = np.zeros((pts.shape[0], 3))                     # :59-65: padded to three coordinates
        points[:, :pts.shape[1]] = pts
        conn = np.ascontiguousarray(mesh.elements, dtype=np.int64).ravel()
        offsets = np.arange(0, offset * mesh.nelem, offset, dtype=np.int64) + offset
        types = np.full(mesh.nelem, cellflag, dtype=np.uint8)
        nincr = int(self.sol_device.shape[2]) if self._sol_h is None else int(self._sol_h.shape[2])
        base = filename.split('.')[0]                            # :68 (the reference cuts at the FIRST dot)
        written = []
        for Increment in range(nincr):                           # :56-73
            name = base + '_quantity_' + str(quantity) + '_increment_' + str(Increment) + '.vtu'
            _write_vtu(name, points, conn, offsets, types, "T", self._field(quantity, Increment))
            written.append(name)
        return written

    def _field(self, quantity, increment):
        """Iterator over contiguous float64 slices of sol[:, quantity, increment] (device fields come over in pieces)."""
        if self._sol_h is not None:
            yield np.ascontiguousarray(self._sol_h[:, quantity, increment], dtype=np.float64)
            return
        col = self.sol_device[:, quantity, increment]
        step = 1 << 22                                           # 32 MB per slice
        for a in range(0, int(col.shape[0]), step):
            yield col[a:a + step].contiguous().cpu().numpy()


def _write_vtu(path, points, conn, offsets, types, field_name, field_slices):
    """VTK XML UnstructuredGrid, version 1.0, little endian, UInt64 headers, raw appended data (pyevtk's layout): every
    array in the appended block is preceded by its byte count."""
    npts, ncell = int(points.shape[0]), int(offsets.shape[0])
    sizes = [points.nbytes, conn.nbytes, offsets.nbytes, types.nbytes, 8 * npts]
    offs = [0]
    for b in sizes[:-1]:
        offs.append(offs[-1] + 8 + b)
    head = ('<?xml version="1.0"?>\n'
            '<VTKFile type="UnstructuredGrid" version="1.0" byte_order="LittleEndian" header_type="UInt64">\n'
            '<UnstructuredGrid>\n<Piece NumberOfPoints="%d" NumberOfCells="%d">\n'
            '<Points>\n<DataArray Name="points" NumberOfComponents="3" type="Float64" format="appended" offset="%d"/>\n</Points>\n'
            '<Cells>\n<DataArray Name="connectivity" NumberOfComponents="1" type="Int64" format="appended" offset="%d"/>\n'
            '<DataArray Name="offsets" NumberOfComponents="1" type="Int64" format="appended" offset="%d"/>\n'
            '<DataArray Name="types" NumberOfComponents="1" type="UInt8" format="appended" offset="%d"/>\n</Cells>\n'
            '<PointData scalars="%s">\n<DataArray Name="%s" NumberOfComponents="1" type="Float64" format="appended" offset="%d"/>\n'
            '</PointData>\n</Piece>\n</UnstructuredGrid>\n<AppendedData encoding="raw">\n_'
            % (npts, ncell, offs[0], offs[1], offs[2], offs[3], field_name, field_name, offs[4]))
    with open(path, "wb") as f:
        f.write(head.encode("ascii"))
        for arr in (points, conn, offsets, types):
            f.write(struct.pack("<Q", arr.nbytes))
            np.ascontiguousarray(arr).astype(arr.dtype.newbyteorder("<"), copy=False).tofile(f)
        f.write(struct.pack("<Q", 8 * npts))
        n = 0
        for piece in field_slices:
            piece.astype("<f8", copy=False).tofile(f)
            n += piece.shape[0]
        if n != npts:
            raise ValueError("field has %d values, mesh has %d points" % (n, npts))
        f.write(b"\n</AppendedData>\n</VTKFile>\n")
