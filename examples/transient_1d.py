#!/usr/bin/env python
"""The reference's examples/transient_1d.py (lines 28-87) on the B200 path: 1D line mesh, forward-Euler loop on device
(kb_transient_run); the returned array is TotalSol[:, 0, [0, 1, N/3, 2N/3, -1]] as in FEMSolver.py:297."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from kiltro_b200 import *  # noqa: F401,F403

mesh = Mesh()
mesh.Line(0.0, 0.02, 5)
ndim = mesh.ndim
material = FourierConduction(ndim, k=10.0, area=1.0, density=10.0e6, heat_capacity_v=1.0)


def InitialFunction(mesh):
    boundary_data = np.zeros((mesh.nnodes, 1), dtype=np.float64)
    boundary_data[:, 0] = 200.0
    return boundary_data


def DirichletFunction(mesh):
    boundary_data = np.zeros((mesh.nnodes, 1), dtype=np.float64) + np.nan
    boundary_data[-1, 0] = 0.0
    return boundary_data


def NeumannFunction(mesh):
    boundary_data = np.zeros((mesh.nnodes, 1), dtype=np.float64) + np.nan
    boundary_data[0, 0] = 0.0
    return boundary_data


boundary_condition = BoundaryCondition()
boundary_condition.SetInitialConditions(InitialFunction, mesh)
boundary_condition.SetDirichletCriteria(DirichletFunction, mesh)
boundary_condition.SetNeumannCriteria(NeumannFunction, mesh)

print('\n======================  PURE-DIFFUSION PROBLEM  ======================')
formulation = HeatDiffusion(mesh)
fem_solver = FEMSolver(analysis_type="transient", number_of_increments=151)
TotalSol = fem_solver.Solve(formulation=formulation, mesh=mesh, material=material, boundary_condition=boundary_condition)
print("time increment", fem_solver.time_increment)       # 0.8 with the current factor 0.1 (FEMSolver.py:144)
print(TotalSol.T)                                          # snapshots 0/1/50/100/150

print('\n===================  ADVECTION-DIFFUSION PROBLEM  ===================')
u = np.zeros((mesh.nnodes, 1), dtype=np.float64) - 1.0e-4
formulation = HeatAdvectionDiffusion(mesh, velocity=u)
fem_solver = FEMSolver(analysis_type="transient", number_of_increments=151)
TotalSol = fem_solver.Solve(formulation=formulation, mesh=mesh, material=material, boundary_condition=boundary_condition)
print(TotalSol.T)
