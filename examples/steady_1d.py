#!/usr/bin/env python
"""BASELINE config 1 -- the reference's examples/steady_1d.py (lines 33-85) on the B200 path.
Only the import line and np.NAN -> np.nan differ from the reference script (plots dropped)."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from kiltro_b200 import *  # noqa: F401,F403

u = 40.0e6
mesh = Mesh(element_type="line")
mesh.Line(left_point=0.0, right_point=0.5, n=5)
ndim = mesh.ndim
material = FourierConduction(ndim, k=1.0e3, area=10.0e3, density=1.0, heat_capacity_v=1.0, q=0.0)


def DirichletFunction(mesh):
    boundary_data = np.zeros((mesh.nnodes, 1), dtype=np.float64) + np.nan
    boundary_data[0, 0] = 100.0
    boundary_data[-1, 0] = 500.0
    return boundary_data


boundary_condition = BoundaryCondition()
boundary_condition.SetDirichletCriteria(DirichletFunction, mesh)

print('\n======================  PURE-DIFFUSION PROBLEM  ======================')
velocity = np.zeros((mesh.nnodes, 1), dtype=np.float64) + u
formulation = Temperature(mesh, velocity=velocity)
fem_solver = FEMSolver(analysis_type="steady")
solution = fem_solver.Solve(formulation=formulation, mesh=mesh, material=material, boundary_condition=boundary_condition)
print(solution.sol[:, 0, 0])          # reference figure steady_1d_diff.pdf: [100 180 260 340 420 500]

print('\n===================  ADVECTION-DIFFUSION PROBLEM  ===================')
formulation = TemperaturePG(mesh, velocity=velocity)
fem_solver = FEMSolver(analysis_type="steady")
solution = fem_solver.Solve(formulation=formulation, mesh=mesh, material=material, boundary_condition=boundary_condition)
print(solution.sol[:, 0, 0])          # steady_1d_addi.pdf: [100 23.72 10.74 26.52 112.98 500]
