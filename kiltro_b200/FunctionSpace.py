"""Interpolation tables, mirroring kiltro/FunctionSpace.py:5-62.

These tables are what the CUDA element kernels evaluate in registers (csrc/element.cuh builds the same numbers from the
same truncated Gauss abscissa 0.57735027); the NumPy copies here exist because the reference exposes them as
attributes (Bases, gBases, AllGauss) and formulations read their shapes."""
import numpy as np

GAUSS_ABSCISSA = 0.57735027          # FunctionSpace.py:30,43 -- the truncated literal, not 1/sqrt(3)


class FunctionSpace(object):

    def __init__(self, mesh):                                    # FunctionSpace.py:5-16
        ndim = mesh.ndim
        if ndim == 1:
            Bases, gBases, w = self.GetBases1D()
        else:
            Bases, gBases, w = self.GetBases2D()
        self.Bases = Bases
        self.gBases = gBases
        self.AllGauss = w

    def Lagrange(self, eta):                                     # FunctionSpace.py:20-24
        return np.array([0.5 * (1 - eta), 0.5 * (1 + eta)]), np.array([-0.5, 0.5])

    def GetBases1D(self):                                        # FunctionSpace.py:27-37
        z = np.array([-GAUSS_ABSCISSA, GAUSS_ABSCISSA])
        Bases = np.zeros((2, 2)); gBases = np.zeros((1, 2, 2))
        for g in range(2):
            Bases[:, g], gBases[0, :, g] = self.Lagrange(z[g])
        return Bases, gBases, np.array([1.0, 1.0])

    def GetBases2D(self):                                        # FunctionSpace.py:40-62
        arranger = np.array([0, 1, 3, 2])
        z = np.array([-GAUSS_ABSCISSA, GAUSS_ABSCISSA])
        Bases = np.zeros((4, 4)); gBases = np.zeros((2, 4, 4)); w = np.zeros(4)
        g = 0
        for i in range(2):
            for j in range(2):
                Ne, dNe = self.Lagrange(z[i]); Nz, dNz = self.Lagrange(z[j])
                Bases[:, g] = np.outer(Ne, Nz).flatten()[arranger]
                gBases[0, :, g] = np.outer(Ne, dNz).flatten()[arranger]
                gBases[1, :, g] = np.outer(dNe, Nz).flatten()[arranger]
                w[g] = 1.0
                g += 1
        return Bases, gBases, w
