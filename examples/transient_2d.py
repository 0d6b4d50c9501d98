#!/usr/bin/env python
"""BASELINE config 3 -- the reference's examples/transient_2d.py (lines 32-94) on the B200 path: the full time loop runs on
device (kb_transient_run); the returned array is TotalSol[:, 0, [0, 1, N/3, 2N/3, -1]] as in FEMSolver.py:297."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from kiltro_b200 import *  # noqa: F401,F403

x_elems, y_elems = 5, 3
mesh = Mesh()
mesh.Rectangle(lower_left_point=(0, 0), upper_right_point=(0.02, 1.0), nx=x_elems, ny=y_elems)
ndim = mesh.ndim
material = FourierConduction(ndim, k=10.0, area=1.0, density=10.0e6, heat_capacity_v=1.0)

left_border = np.where(mesh.points[:, 0] == 0.0)[0]
right_border = np.where(mesh.points[:, 0] == 0.02)[0]


def InitialFunction(mesh):
    boundary_data = np.zeros((mesh.nnodes, 1), dtype=np.float64)
    boundary_data[:, 0] = 200.0
    return boundary_data


def DirichletFunction(mesh, right_border):
    boundary_data = np.zeros((mesh.nnodes, 1), dtype=np.float64) + np.nan
    boundary_data[right_border, 0] = 0.0
    return boundary_data


def NeumannFunction(mesh, left_border):
    boundary_data = np.zeros((mesh.nnodes, 1), dtype=np.float64) + np.nan
    boundary_data[left_border, 0] = 0.0
    return boundary_data


boundary_condition = BoundaryCondition()
boundary_condition.SetInitialConditions(InitialFunction, mesh)
boundary_condition.SetDirichletCriteria(DirichletFunction, mesh, right_border)
boundary_condition.SetNeumannCriteria(NeumannFunction, mesh, left_border)

print('\n======================  PURE-DIFFUSION PROBLEM  ======================')
formulation = HeatDiffusion(mesh)
fem_solver = FEMSolver(analysis_type="transient", number_of_increments=61)
TotalSol = fem_solver.Solve(formulation=formulation, mesh=mesh, material=material, boundary_condition=boundary_condition)
print("time increment", fem_solver.time_increment)       # 0.8 with the current factor 0.1 (FEMSolver.py:144)
print(TotalSol.reshape(y_elems + 1, x_elems + 1, 5)[1].T)  # row 1, snapshots 0/1/20/40/60
