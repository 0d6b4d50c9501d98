"""kiltro_b200 -- B200-native implementation of Kiltro's FEM hot path behind Kiltro's own Python API.

`from kiltro_b200 import *` exposes the names `from kiltro import *` does (kiltro/__init__.py:1-8) that lie on the hot
path: Mesh, BoundaryCondition, FunctionSpace, Assemble, ComputeSparsityPattern, FEMSolver, LinearSolver, Material,
FourierConduction, Temperature, TemperaturePG (+ the example scripts' HeatDiffusion / HeatAdvectionDiffusion),
PostProcess.  Computation happens in libkiltro_b200.so (hand-written sm_100a CUDA); there is no CPU fallback."""
from .MeshGeneration import Mesh
from .BoundaryCondition import BoundaryCondition
from .FunctionSpace import FunctionSpace
from .FiniteElements import Assemble, ComputeSparsityPattern, AssemblyRobinForces, AssembleCharacteristicGalerkin
from .FiniteElements.ComputeSparsityPattern import ComputeSparsityPatternDevice, SparsityPattern
from .FEMSolver import FEMSolver, LinearSolver
from .MaterialBase import Material, FourierConduction
from .VariationalPrinciple import (VariationalPrinciple, Temperature, TemperaturePG, HeatDiffusion,
                                   HeatAdvectionDiffusion, HeatAdvectionDiffusionPG)
from .PostProcess import PostProcess
from .sparse import DeviceCSR
from .device import to_numpy
from .partition import StripPartition

__all__ = ["Mesh", "BoundaryCondition", "FunctionSpace", "Assemble", "ComputeSparsityPattern",
           "ComputeSparsityPatternDevice", "SparsityPattern", "AssemblyRobinForces", "AssembleCharacteristicGalerkin", "FEMSolver", "LinearSolver", "Material",
           "FourierConduction", "VariationalPrinciple", "Temperature", "TemperaturePG", "HeatDiffusion",
           "HeatAdvectionDiffusion", "HeatAdvectionDiffusionPG", "PostProcess", "DeviceCSR", "to_numpy", "StripPartition"]
