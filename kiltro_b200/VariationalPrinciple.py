"""Variational formulations, mirroring kiltro/VariationalPrinciple.py.

Temperature / TemperaturePG are the library's classes (scalar-broadcast advection term as shipped: SURVEY F3,
VariationalPrinciple.py:152,156,164-173).  HeatDiffusion / HeatAdvectionDiffusion are the names the reference's example
scripts use (examples/steady_2d.py:59,80); they select the consistent form W (x) u.gradN that survives in
MaterialBase.py:42-58 and reproduces the shipped figures.  The arithmetic runs in csrc/element.cuh."""
import numpy as np
import torch

from . import _lib, device as D
from .FunctionSpace import FunctionSpace


class VariationalPrinciple(object):

    mode = _lib.KB_MODE_REFERENCE
    petrov = False

    def __init__(self, mesh, function_spaces=None):              # VariationalPrinciple.py:9-13
        self.nvar = None
        self.ndim = mesh.ndim
        self.function_spaces = function_spaces

    def GetFunctionSpaces(self, mesh, function_spaces=None):     # :16-39
        if function_spaces is None and self.function_spaces is None:
            self.function_spaces = (FunctionSpace(mesh),)
        else:
            self.function_spaces = function_spaces if function_spaces is not None else self.function_spaces
        local_size = self.function_spaces[0].Bases.shape[0] * self.nvar
        self.local_rows = np.repeat(np.arange(0, local_size), local_size, axis=0)
        self.local_columns = np.tile(np.arange(0, local_size), local_size)
        self.local_size = local_size
        local_size_m = self.function_spaces[0].Bases.shape[0] * self.ndim
        self.local_rows_mass = np.repeat(np.arange(0, local_size_m), local_size_m, axis=0)
        self.local_columns_mass = np.tile(np.arange(0, local_size_m), local_size_m)
        self.local_size_m = local_size_m

    def FindIndices(self, A):                                    # :72-73
        return self.local_rows, self.local_columns, A.ravel()

    # -------------------------------------------------------------------------------- device data
    def device_velocity(self):
        """(nnodes, ndim) fp64 device tensor, or None for an identically zero field (skips the loads in the kernel)."""
        v = self._velocity
        if v is None:
            return None
        return D.to_device(v, torch.float64)

    @property
    def is_symmetric(self):
        """The assembled matrix is symmetric for the as-shipped (scalar-broadcast) advection term and for pure
        diffusion; the consistent advection term makes it non-symmetric (-> BiCGSTAB)."""
        return self.mode == _lib.KB_MODE_REFERENCE or self._velocity is None

    # -------------------------------------------------------------------------------- K1 (batched, device)
    def GetAllElementalMatrices(self, mesh, material, want_mass=True, elements=None):
        """All element matrices in one launch: (Ke (ne,npe^2), Fe (ne,npe), Me (ne,npe^2)|None) device tensors."""
        lib = _lib.load()
        pts = mesh.device_points()
        el = mesh.device_elements() if elements is None else elements
        ne, npe = int(el.shape[0]), int(el.shape[1])
        Ke = D.empty((ne, npe * npe)); Fe = D.empty((ne, npe))
        Me = D.empty((ne, npe * npe)) if want_mass else None
        vel = self.device_velocity()
        mat = material.as_struct()
        _lib.check(lib.kb_element_matrices(self.ndim, D.ptr(pts), D.ptr(el), ne, D.ptr(vel), mat, int(self.mode),
                                           int(bool(self.petrov)), D.ptr(Ke), D.ptr(Fe), D.ptr(Me), D.stream_ptr()))
        return Ke, Fe, Me

    # -------------------------------------------------------------------------------- reference per-element seam
    def _one(self, elem, mesh, material, want_mass):
        el = mesh.device_elements()[elem:elem + 1].contiguous()
        Ke, Fe, Me = self.GetAllElementalMatrices(mesh, material, want_mass, elements=el)
        n = int(el.shape[1])
        return (D.to_numpy(Ke).reshape(n, n), D.to_numpy(Fe).reshape(n, 1),
                D.to_numpy(Me).reshape(n, n) if want_mass else None)

    def GetLocalConvection(self, elem, function_space, mesh, material, ElemCoords=None):      # :117-161 / :216-261
        Ke, Fe, _ = self._one(elem, mesh, material, False)
        return Ke, Fe

    def GetLocalMass(self, elem, function_space, mesh, material, ElemCoords=None):            # :42-69
        return self._one(elem, mesh, material, True)[2]

    def GetElementalMatrices(self, elem, function_space, mesh, material, fem_solver):         # :94-113 / :194-213
        transient = fem_solver.analysis_type != 'steady'
        Ke, Fe, Me = self._one(elem, mesh, material, transient)
        I, J, V = self.FindIndices(Ke)
        Im, Jm, Vm = [], [], []
        if transient:
            Im, Jm, Vm = self.FindIndices(Me)
        return I, J, V, Fe, Im, Jm, Vm


def _init_thermal(self, mesh, velocity, function_spaces):
    self.fields = "thermal"
    self.nvar = 1
    if velocity is not None:
        if not velocity.shape[1] == self.ndim:
            raise ValueError("Dimension of velocity does not match with problem dimension")
    self._velocity = velocity
    self.GetFunctionSpaces(mesh, function_spaces=function_spaces)
    self._nnodes = mesh.nnodes


class Temperature(VariationalPrinciple):
    """kiltro/VariationalPrinciple.py:76-173 (Galerkin weights, advection as shipped)."""
    mode = _lib.KB_MODE_REFERENCE
    petrov = False

    def __init__(self, mesh, velocity=None, function_spaces=None):
        super(Temperature, self).__init__(mesh, function_spaces=function_spaces)
        _init_thermal(self, mesh, velocity, function_spaces)

    @property
    def velocity(self):
        """Like the reference's attribute: zeros((nnodes, ndim)) when no velocity was given (VariationalPrinciple.py:
        84-85).  Internally a missing field stays None so the kernels skip its loads."""
        if self._velocity is None:
            return np.zeros((self._nnodes, self.ndim), dtype=np.float64)
        return self._velocity

    @velocity.setter
    def velocity(self, v):
        self._velocity = v


class TemperaturePG(Temperature):
    """kiltro/VariationalPrinciple.py:176-342 (Petrov-Galerkin weights W = N + alpha/|u| u.dN/dxi)."""
    petrov = True


class HeatDiffusion(Temperature):
    """examples/steady_2d.py:59 -- pure conduction (consistent form; identical to Temperature for zero velocity)."""
    mode = _lib.KB_MODE_CONSISTENT

    def __init__(self, mesh, function_spaces=None):
        super(HeatDiffusion, self).__init__(mesh, velocity=None, function_spaces=function_spaces)


class HeatAdvectionDiffusion(Temperature):
    """examples/steady_2d.py:80 -- consistent Galerkin advection-diffusion (MaterialBase.py:42-58)."""
    mode = _lib.KB_MODE_CONSISTENT


class HeatAdvectionDiffusionPG(Temperature):
    """Consistent advection term with the reference's Petrov-Galerkin weights (streamline upwinding)."""
    mode = _lib.KB_MODE_CONSISTENT
    petrov = True
