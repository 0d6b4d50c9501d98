#!/usr/bin/env python
"""BASELINE config 2 -- the reference's examples/steady_2d.py (lines 33-94) on the B200 path."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from kiltro_b200 import *  # noqa: F401,F403

x_elems, y_elems = 5, 3
mesh = Mesh(element_type="quad")
mesh.Rectangle(lower_left_point=(0, 0), upper_right_point=(0.5, 0.01), nx=x_elems, ny=y_elems)
ndim = mesh.ndim
material = FourierConduction(ndim, k=1.0e3, area=1.0, density=1.0, c_v=1.0, q=0.0)

left_border = np.where(mesh.points[:, 0] == 0.0)[0]
right_border = np.where(mesh.points[:, 0] == 0.5)[0]


def DirichletFunction(mesh, left_border, right_border):
    boundary_data = np.zeros((mesh.nnodes, 1), dtype=np.float64) + np.nan
    boundary_data[left_border, 0] = 100.0
    boundary_data[right_border, 0] = 500.0
    return boundary_data


boundary_condition = BoundaryCondition()
boundary_condition.SetDirichletCriteria(DirichletFunction, mesh, left_border, right_border)

print('\n======================  PURE-DIFFUSION PROBLEM  ======================')
formulation = HeatDiffusion(mesh)
fem_solver = FEMSolver(analysis_type="steady")
solution = fem_solver.Solve(formulation=formulation, mesh=mesh, material=material, boundary_condition=boundary_condition)
print(solution.sol.reshape(y_elems + 1, x_elems + 1))     # steady_2d_diff.pdf: rows [100 180 260 340 420 500]

print('\n===================  ADVECTION-DIFFUSION PROBLEM  ===================')
u = np.zeros((mesh.nnodes, 2), dtype=np.float64)
u[:, 0] += 9.0e3
formulation = HeatAdvectionDiffusion(mesh, velocity=u)
fem_solver = FEMSolver(analysis_type="steady")
solution = fem_solver.Solve(formulation=formulation, mesh=mesh, material=material, boundary_condition=boundary_condition)
print(solution.sol.reshape(y_elems + 1, x_elems + 1))     # steady_2d_addi.pdf: [100 105.16 118.92 154.91 249.82 500]
