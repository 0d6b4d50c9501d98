"""Mesh container + structured generators, mirroring kiltro/MeshGeneration/Mesh.py (reference file:line cited per method).

points / elements live in HBM (fp64 / int32); the NumPy attributes `mesh.points` (float64) and `mesh.elements` (int64,
like the reference) are materialised lazily on first access.  Once a host copy has been handed out it is treated as the
source of truth (user scripts edit it in place), and it is re-uploaded whenever the device copy is requested."""
import numpy as np
import torch

from .. import _lib, device as D


class Mesh(object):

    def __init__(self, element_type=None):                       # Mesh.py:5-11
        self._points_h = None
        self._elements_h = None
        self._points_d = None
        self._elements_d = None
        self.nnodes = None
        self.nelem = None
        self.ndim = None
        self.element_type = element_type
        self.structured_shape = None      # (nx_nodes, ny_nodes) for generator-made meshes (tiling hint only)
        self.partition = None             # StripPartition when this is one rank's strip (+ halo rows) of a larger mesh

    # ---------------------------------------------------------------- host views (lazy)
    @property
    def points(self):
        if self._points_h is None and self._points_d is not None:
            self._points_h = D.to_numpy(self._points_d)
        return self._points_h

    @points.setter
    def points(self, value):
        self._points_h = None if value is None else np.asarray(value, dtype=np.float64)
        self._points_d = None
        if value is not None:
            self.nnodes = self._points_h.shape[0]
            self.ndim = self._points_h.shape[1]
            self.structured_shape = None

    @property
    def elements(self):
        if self._elements_h is None and self._elements_d is not None:
            self._elements_h = D.to_numpy(self._elements_d).astype(np.int64)
        return self._elements_h

    @elements.setter
    def elements(self, value):
        self._elements_h = None if value is None else np.asarray(value)
        self._elements_d = None
        if value is not None:
            self.nelem = self._elements_h.shape[0]
            self.structured_shape = None

    # ---------------------------------------------------------------- device views
    def device_points(self):
        if self._points_h is not None:                           # host copy exists -> it is the truth
            return D.to_device(self._points_h, torch.float64)
        return self._points_d

    def device_elements(self):
        if self._elements_h is not None:
            return D.to_device(self._elements_h, torch.int32)
        return self._elements_d

    def host_is_authoritative(self):
        return self._points_h is not None or self._elements_h is not None

    # ---------------------------------------------------------------- generators (K8)
    def Line(self, left_point=0.0, right_point=1.0, n=10):      # Mesh.py:14-30
        left_point = float(left_point); right_point = float(right_point); n = int(n)
        lib = _lib.load()
        pts = D.empty((n + 1, 1)); el = D.empty((n, 2), torch.int32)
        _lib.check(lib.kb_mesh_line(left_point, right_point, n, D.ptr(pts), D.ptr(el), D.stream_ptr()))
        self._set_device(pts, el, 1)
        if self.element_type is None:
            self.element_type = "line"
        self.structured_shape = (n + 1, 1)

    def Rectangle(self, lower_left_point=(0, 0), upper_right_point=(1, 1), nx=5, ny=5, partition=None):      # Mesh.py:33-58
        """partition (extension): a kiltro_b200.StripPartition of the (nx+1) x (ny+1) node grid -- the mesh then
        holds only this rank's mesh rows plus one halo row per interior side, in local numbering (coordinates bit-identical
        to the global mesh); use with FEMSolver(partition="rows")."""
        nx, ny = int(nx), int(ny)
        lib = _lib.load()
        if partition is not None:
            part = partition
            if nx + 1 != part.nxn or ny + 1 != part.nyn:
                raise ValueError("partition does not match the rectangle")
            pts = D.empty((part.n_loc, 2)); el = D.empty(((part.nrows_loc - 1) * nx, 4), torch.int32)
            _lib.check(lib.kb_mesh_rectangle_strip(float(lower_left_point[0]), float(lower_left_point[1]),
                                                   float(upper_right_point[0]), float(upper_right_point[1]), nx, ny,
                                                   part.j_first, part.nrows_loc, D.ptr(pts), D.ptr(el), D.stream_ptr()))
            self._set_device(pts, el, 2)
            if self.element_type is None:
                self.element_type = "quad"
            self.structured_shape = (part.nxn, part.nrows_loc)
            self.partition = part
            return
        pts = D.empty(((nx + 1) * (ny + 1), 2)); el = D.empty((nx * ny, 4), torch.int32)
        _lib.check(lib.kb_mesh_rectangle(float(lower_left_point[0]), float(lower_left_point[1]),
                                         float(upper_right_point[0]), float(upper_right_point[1]), nx, ny,
                                         D.ptr(pts), D.ptr(el), D.stream_ptr()))
        self._set_device(pts, el, 2)
        if self.element_type is None:
            self.element_type = "quad"
        self.structured_shape = (nx + 1, ny + 1)

    def global_node_ids(self):
        """int64 device tensor: global node id of every local node (arange for an unpartitioned mesh)."""
        off = 0 if self.partition is None else self.partition.j_first * self.partition.nxn
        return off + torch.arange(self.nnodes, dtype=torch.int64, device="cuda")

    def set_arrays(self, points, elements, partition=None):
        """Adopt caller-provided arrays (NumPy, torch host/pinned or torch CUDA) by uploading them once: the device
        copies become the source of truth (use the `points` / `elements` setters for host arrays you intend to edit).
        partition: the arrays are one rank's strip (+ halo rows, local numbering) of a partitioned mesh."""
        pts = D.to_device(points, torch.float64)
        el = D.to_device(elements, torch.int32)
        self._set_device(pts, el, int(pts.shape[1]))
        if self.element_type is None:
            self.element_type = "quad" if el.shape[1] == 4 else "line"
        self.structured_shape = None
        self.partition = partition

    def _set_device(self, pts, el, ndim):
        self._points_d, self._elements_d = pts, el
        self._points_h = self._elements_h = None
        self.nnodes, self.nelem, self.ndim = int(pts.shape[0]), int(el.shape[0]), ndim

    # ---------------------------------------------------------------- queries
    def InferSpatialDimension(self):                             # Mesh.py:61-69
        assert self.nnodes is not None
        return self.ndim if self._points_h is None else self._points_h.shape[1]

    def InferNumberOfNodesPerElement(self, p=None, element_type=None):      # Mesh.py:72-92
        if p is not None and element_type is not None:
            if element_type == "line":
                return int(p + 1)
            elif element_type == "tri":
                return int((p + 1) * (p + 2) / 2)
            elif element_type == "quad":
                return int((p + 1) ** 2)
            elif element_type == "tet":
                return int((p + 1) * (p + 2) * (p + 3) / 6)
            elif element_type == "hex":
                return int((p + 1) ** 3)
            else:
                raise ValueError("Did not understand element type")
        assert self.nelem is not None
        return int((self._elements_h if self._elements_h is not None else self._elements_d).shape[1])

    def __deepcopy__(self, memo):
        m = Mesh(self.element_type)
        m._points_h = None if self._points_h is None else self._points_h.copy()
        m._elements_h = None if self._elements_h is None else self._elements_h.copy()
        m._points_d, m._elements_d = self._points_d, self._elements_d      # immutable on device: share
        m.nnodes, m.nelem, m.ndim, m.structured_shape = self.nnodes, self.nelem, self.ndim, self.structured_shape
        m.partition = self.partition
        return m
