"""Material constants, mirroring kiltro/MaterialBase.py:4-39.  The per-Gauss-point integrands the reference keeps on
the material (HeatDiffusion/HeatConvection/HeatMass/HeatGeneration, MaterialBase.py:42-77) are evaluated inside the CUDA
element kernel (csrc/element.cuh, KB_MODE_CONSISTENT); only the constants cross the boundary."""
from . import _lib


class Material(object):

    def __init__(self, mtype, ndim, density=None, conductivity=None, heat_capacity_v=None, **kwargs):   # :6-23
        if not isinstance(mtype, str):
            raise TypeError("Type of material model should be given as a string")
        self.k = conductivity
        self.rho = density
        self.c_v = heat_capacity_v
        for i in kwargs.keys():
            if "__" not in i:
                setattr(self, i, kwargs[i])
        self.mtype = mtype
        self.ndim = ndim


class FourierConduction(Material):

    def __init__(self, ndim, **kwargs):                          # MaterialBase.py:28-39
        mtype = type(self).__name__
        super(FourierConduction, self).__init__(mtype, ndim, **kwargs)
        if self.k is None and "conductivity" not in self.__dict__:
            raise ValueError("Thermal conductivity not defined.")
        if "q" not in self.__dict__:
            self.q = 0.0
        self.nvar = 1

    def as_struct(self):
        """The five scalars the kernels need (k, rho, c_v, area, q)."""
        missing = [n for n in ("k", "rho", "c_v") if getattr(self, n, None) is None]
        if missing or not hasattr(self, "area"):
            raise ValueError("material is missing %s" % ", ".join(missing + ([] if hasattr(self, "area") else ["area"])))
        return _lib.kb_material(float(self.k), float(self.rho), float(self.c_v), float(self.area), float(self.q))
