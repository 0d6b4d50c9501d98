"""Result holder, mirroring kiltro/PostProcess.py:6-23.  `.sol` is (nnodes, nvar, nincr) float64 like the reference's; it
is kept in HBM (`.sol_device`) and copied to the host on first access.  VTK output (PostProcess.py:26-76) is outside the
hot path (out of scope) and raises."""
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

    def WriteVTK(self, filename=None, quantity=0):
        raise NotImplementedError("WriteVTK (kiltro/PostProcess.py:26-76, needs pyevtk) is outside the B200 hot path; "
                                  "use solution.sol (NumPy) with the writer of your choice")
