"""CSR matrix resident in HBM.  Stands in for the scipy.sparse.csr_matrix objects the reference passes between
Assemble, BoundaryCondition and LinearSolver (FiniteElements/Assembly.py:78-87, BoundaryCondition.py:156-200)."""
import torch

from . import _lib, device as D


class DeviceCSR(object):
    """data (fp64), indices (int32), indptr (int32) as device tensors; shape like scipy's."""

    def __init__(self, data, indices, indptr, shape, symmetric=None):
        self.data, self.indices, self.indptr = data, indices, indptr
        self.shape = (int(shape[0]), int(shape[1]))
        self.symmetric = symmetric          # True / False / None (unknown): picks CG vs BiCGSTAB
        self._rowblk = None

    @property
    def nnz(self):
        return int(self.data.shape[0])

    def row_blocks(self):
        """Row partition with ~equal nnz per CTA for the SpMV kernel (cached)."""
        if self._rowblk is None:
            lib = _lib.load()
            nblk = lib.kb_spmv_num_blocks(self.nnz)
            rb = D.empty(nblk + 1, torch.int32)
            _lib.check(lib.kb_spmv_partition(self.shape[0], self.nnz, D.ptr(self.indptr), D.ptr(rb), D.stream_ptr()))
            self._rowblk = rb
        return self._rowblk

    def dot(self, x):
        """y = A x on device (K5)."""
        lib = _lib.load()
        x = D.to_device(x, torch.float64).reshape(-1)
        if x.shape[0] != self.shape[1]:
            raise ValueError("dimension mismatch")
        y = D.empty(self.shape[0])
        _lib.check(lib.kb_spmv(self.shape[0], self.nnz, D.ptr(self.indptr), D.ptr(self.indices), D.ptr(self.data),
                               D.ptr(self.row_blocks()), D.ptr(x), D.ptr(y), D.stream_ptr()))
        return y

    __matmul__ = dot

    def tocsr(self):
        """Copy to the host as scipy.sparse.csr_matrix (tests / interoperability; not used by the solve path)."""
        from scipy.sparse import csr_matrix
        return csr_matrix((D.to_numpy(self.data), D.to_numpy(self.indices), D.to_numpy(self.indptr)), shape=self.shape)

    toscipy = tocsr

    def toarray(self):
        return self.tocsr().toarray()


def issparse(A):
    return isinstance(A, DeviceCSR)
