#!/usr/bin/env python
"""Golden fixtures for the SURVEY section 8(f) "next" rows, again by RUNNING THE REFERENCE ITSELF.

  n2  AssemblyRobinForces              kiltro/FiniteElements/Assembly.py:121-175 (runs inside Assemble, result discarded)
  n3  AssembleCharacteristicGalerkin   kiltro/FiniteElements/Assembly.py:95-117 + TemperaturePG.
      GetElementalCharacteristicGalerkin, kiltro/VariationalPrinciple.py:264-308 (never called by the solver)
Both routines are importable and are called here directly on small meshes; their dense outputs are stored.  Same scratch
build of the reference as make_golden.py (test infrastructure only; /root/reference is not needed to run the tests).

Usage:  python tests/golden/make_golden_next.py
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import make_golden as MG                                                     # noqa: E402


def main():
    MG.build_reference()
    import kiltro as K
    from kiltro.FiniteElements.Assembly import AssemblyRobinForces, AssembleCharacteristicGalerkin

    def rect(ll, ur, nx, ny):
        m = K.Mesh(element_type="quad"); m.Rectangle(lower_left_point=ll, upper_right_point=ur, nx=nx, ny=ny)
        return m

    def line(l, r, n):
        m = K.Mesh(element_type="line"); m.Line(l, r, n)
        return m

    def distort(m, rng, amp):
        if m.ndim == 1:
            h = np.min(np.diff(m.points[:, 0]))
            m.points[1:-1, 0] += amp * h * rng.uniform(-1, 1, m.nnodes - 2)
            return
        x, y = m.points[:, 0], m.points[:, 1]
        hx = (x.max() - x.min()) / (len(np.unique(x)) - 1)
        hy = (y.max() - y.min()) / (len(np.unique(y)) - 1)
        inner = (x > x.min()) & (x < x.max()) & (y > y.min()) & (y < y.max())
        m.points[inner, 0] += amp * hx * rng.uniform(-1, 1, inner.sum())
        m.points[inner, 1] += amp * hy * rng.uniform(-1, 1, inner.sum())

    out = {}
    rng = np.random.default_rng(2024)
    cases = {
        "quad": (rect((0.0, 0.0), (1.5, 1.0), 7, 5), K.FourierConduction(2, k=2.5, area=1.3, density=1.7, heat_capacity_v=0.9, q=3.0)),
        "quad_small": (rect((0.0, 0.0), (0.5, 0.01), 5, 3), K.FourierConduction(2, k=1.0e3, area=1.0, density=1.0, heat_capacity_v=1.0, q=0.0)),
        "line": (line(0.0, 2.0, 11), K.FourierConduction(1, k=0.8, area=2.0, density=1.1, heat_capacity_v=1.4, q=-2.0)),
    }
    distort(cases["quad"][0], rng, 0.22)
    distort(cases["line"][0], rng, 0.3)
    for name, (m, mat) in cases.items():
        vel = (np.array([1.2, -0.4])[:m.ndim] * 6.0 + 2.0 * rng.standard_normal((m.nnodes, m.ndim)))
        form = K.TemperaturePG(m, velocity=vel)
        fs = K.FEMSolver(analysis_type="steady")
        bc = K.BoundaryCondition()
        h, tinf = 7.5, 293.0
        bc.applied_robin_convection = h                    # what RobinSelector stores (BoundaryCondition.py:88-91)
        bc.ground_robin_convection = tinf
        with MG.quiet():
            Kr, Fr = AssemblyRobinForces(fs, form.function_spaces[0], m, mat, bc)
            Ku, Fu = AssembleCharacteristicGalerkin(form.function_spaces, form, m, mat, bc)
        out[f"next/{name}/points"] = m.points.copy()
        out[f"next/{name}/elements"] = m.elements.astype(np.int64)
        out[f"next/{name}/velocity"] = vel
        out[f"next/{name}/material"] = np.array([mat.k, mat.rho, mat.c_v, mat.area, mat.q])
        out[f"next/{name}/robin_h"] = np.float64(h); out[f"next/{name}/robin_tinf"] = np.float64(tinf)
        out[f"next/{name}/K_robin"] = Kr; out[f"next/{name}/F_robin"] = Fr.ravel()
        out[f"next/{name}/K_cg"] = Ku; out[f"next/{name}/F_cg"] = Fu.ravel()
    path = os.path.join(HERE, "next.npz")
    np.savez_compressed(path, **{k.split("/", 1)[1].replace("/", "__"): v for k, v in out.items()})
    print("wrote", path, os.path.getsize(path) / 1024, "KiB")


if __name__ == "__main__":
    main()
