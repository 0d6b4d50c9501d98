#!/usr/bin/env python
"""Generate the golden fixtures under tests/golden/ by RUNNING THE REFERENCE ITSELF.

Test infrastructure only.  This script is the one place that touches /root/reference:
it copies the reference package to a scratch directory (the reference tree is read-only
and its shipped ComputeSparsityPattern.so needs libpython3.8), rebuilds the Cython/C++
extension from the reference's own .pyx/.h exactly as kiltro/FiniteElements/Makefile:11-16
does, imports it, and dumps inputs + outputs of every seam of the hot path as small .npz
files.  The .npz files are committed; /root/reference does not exist on the GPU box.

Declared shims:
  * stdout -> /dev/null while the reference runs (it prints once per element,
    VariationalPrinciple.py:125,224,328);
  * mode="consistent": subclass of the reference's Temperature whose GetLocalConvection
    sums the reference's own MaterialBase.HeatDiffusion + HeatConvection (MaterialBase.py:42-58);
  * transient: FEMSolver.TransientSolver is unrunnable as shipped; the loop below
    restates FEMSolver.py:253-297 on top of the reference's importable Assemble (K and M).

Usage:  python tests/golden/make_golden.py      (needs /root/reference, cython, g++)
"""
import contextlib
import io
import os
import shutil
import subprocess
import sys
import sysconfig

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REF = "/root/reference"
SCRATCH = "/tmp/kiltro_ref_scratch"


def build_reference():
    if not os.path.isdir(REF):
        raise SystemExit("reference tree not present; golden fixtures are already committed")
    if os.path.isdir(SCRATCH):
        shutil.rmtree(SCRATCH)
    os.makedirs(SCRATCH)
    shutil.copytree(os.path.join(REF, "kiltro"), os.path.join(SCRATCH, "kiltro"))
    fe = os.path.join(SCRATCH, "kiltro", "FiniteElements")
    for f in ("ComputeSparsityPattern.so", "ComputeSparsityPattern.cpp"):
        p = os.path.join(fe, f)
        if os.path.exists(p):
            os.remove(p)
    subprocess.check_call(["cython", "--cplus", "-3", "ComputeSparsityPattern.pyx"], cwd=fe)
    subprocess.check_call(
        ["g++", "-fPIC", "-shared", "-pthread", "-O3", "-fno-strict-aliasing", "-DNDEBUG",
         "ComputeSparsityPattern.cpp", "-o", "ComputeSparsityPattern.so", "-I.",
         "-I" + sysconfig.get_paths()["include"], "-I" + np.get_include()], cwd=fe)
    sys.path.insert(0, SCRATCH)


@contextlib.contextmanager
def quiet():
    with contextlib.redirect_stdout(io.StringIO()):
        yield


def main():
    build_reference()
    import kiltro as K
    from kiltro.FiniteElements.ComputeSparsityPattern import ComputeSparsityPattern as CSP
    from kiltro.FiniteElements.Assembly import Assemble
    import kiltro.FiniteElements.Assembly as AsmMod
    from scipy.sparse.linalg import spsolve

    # ------------------------------------------------------------------ consistent-mode shim (F4)
    class TemperatureConsistent(K.Temperature):
        petrov = False

        def GetLocalConvection(self, elem, function_space, mesh, material, ElemCoords):
            Bases, gBases, w = function_space.Bases, function_space.gBases, function_space.AllGauss
            u = np.mean(self.velocity[mesh.elements[elem]], axis=0)
            PG = np.einsum('ijk,jl->kil', gBases, ElemCoords)
            G = np.einsum('ijk,kli->ijl', np.linalg.inv(PG), gBases)
            dV = np.einsum('i,i->i', w, np.abs(np.linalg.det(PG)))
            n = Bases.shape[0]
            Ke = np.zeros((n, n)); fe = np.zeros((n, 1))
            for g in range(dV.shape[0]):
                DB = material.HeatDiffusion(G, gcounter=g) + material.HeatConvection(Bases, G, u, gcounter=g)
                Ke += DB * dV[g]
                fe += (material.area * material.q * Bases[:, g])[:, None] * dV[g]
            return Ke, fe

    def make_mesh(kind, **kw):
        if kind == "line":
            m = K.Mesh(element_type="line"); m.Line(kw["l"], kw["r"], kw["n"])
        else:
            m = K.Mesh(element_type="quad")
            m.Rectangle(lower_left_point=kw["ll"], upper_right_point=kw["ur"], nx=kw["nx"], ny=kw["ny"])
        return m

    def distort(m, rng, amp):
        """Move interior nodes of a Rectangle mesh by amp*h (keeps elements convex for amp<0.25)."""
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

    def scramble(m, rng):
        """Unstructured-looking input: permute element order, rotate local node order (keeps CCW)."""
        perm = rng.permutation(m.nelem)
        el = m.elements[perm].copy()
        if el.shape[1] == 4:
            for e in range(el.shape[0]):
                el[e] = np.roll(el[e], rng.integers(0, 4))
        m.elements = np.ascontiguousarray(el)

    out = {}

    # ------------------------------------------------------------------ 1. sparsity pattern (native piece)
    rng = np.random.default_rng(1234)
    pat_cases = {
        "line3": make_mesh("line", l=0.0, r=1.0, n=3),
        "line17": make_mesh("line", l=-1.0, r=2.0, n=17),
        "rect2x2": make_mesh("quad", ll=(0, 0), ur=(1, 1), nx=2, ny=2),
        "rect5x3": make_mesh("quad", ll=(0, 0), ur=(0.5, 0.01), nx=5, ny=3),
        "rect1x1": make_mesh("quad", ll=(0, 0), ur=(1, 1), nx=1, ny=1),
        "rect9x1": make_mesh("quad", ll=(0, 0), ur=(3, 1), nx=9, ny=1),
        "rect7x6s": make_mesh("quad", ll=(0, 0), ur=(2, 1), nx=7, ny=6),
    }
    scramble(pat_cases["rect7x6s"], rng)
    for name, m in pat_cases.items():
        ind, ptr, dl, dg = CSP(m, 1)
        out[f"pattern/{name}/elements"] = m.elements.astype(np.int64)
        out[f"pattern/{name}/nnodes"] = np.int64(m.nnodes)
        out[f"pattern/{name}/indices"] = ind
        out[f"pattern/{name}/indptr"] = ptr
        out[f"pattern/{name}/data_local"] = dl
        out[f"pattern/{name}/data_global"] = dg
    # nvar = 2 on one case (the native routine is generic in nvar)
    ind, ptr, dl, dg = CSP(pat_cases["rect2x2"], 2)
    out["pattern/rect2x2_nvar2/indices"] = ind
    out["pattern/rect2x2_nvar2/indptr"] = ptr
    out["pattern/rect2x2_nvar2/data_local"] = dl
    out["pattern/rect2x2_nvar2/data_global"] = dg

    # ------------------------------------------------------------------ 2./3. element matrices + Assemble
    def assemble_case(name, m, material, velocity, cls, transient):
        form = cls(m, velocity=velocity)
        fs = K.FEMSolver(analysis_type="transient" if transient else "steady")
        bc = K.BoundaryCondition()
        with quiet():
            fs.ComputeSparsityFEM(m, form)
            Kc, Flux, Mm = Assemble(fs, form.function_spaces, form, m, material, bc)
            Ke = np.zeros((m.nelem, m.elements.shape[1] ** 2)); Fe = np.zeros((m.nelem, m.elements.shape[1]))
            Me = np.zeros_like(Ke)
            for e in range(m.nelem):
                X = m.points[m.elements[e]]
                k_, f_ = form.GetLocalConvection(e, form.function_spaces[0], m, material, X)
                Ke[e] = k_.ravel(); Fe[e] = f_.ravel()
                Me[e] = form.GetLocalMass(e, form.function_spaces[0], m, material, X).ravel()
        out[f"{name}/points"] = m.points.copy()
        out[f"{name}/elements"] = m.elements.astype(np.int64)
        out[f"{name}/velocity"] = form.velocity.copy()
        out[f"{name}/material"] = np.array([material.k, material.rho, material.c_v, material.area, material.q])
        out[f"{name}/Ke"] = Ke; out[f"{name}/Fe"] = Fe; out[f"{name}/Me"] = Me
        out[f"{name}/K_data"] = Kc.data.copy(); out[f"{name}/K_indices"] = Kc.indices.copy()
        out[f"{name}/K_indptr"] = Kc.indptr.copy(); out[f"{name}/Flux"] = Flux.ravel().copy()
        if transient:
            out[f"{name}/M_data"] = Mm.data.copy()
        return form, fs, Kc, Flux, Mm

    rng = np.random.default_rng(7)
    mq = make_mesh("quad", ll=(0.0, 0.0), ur=(1.5, 1.0), nx=9, ny=6); distort(mq, rng, 0.22)
    matq = K.FourierConduction(2, k=2.5, area=1.3, density=1.7, heat_capacity_v=0.9, q=3.0)
    vq = np.array([1.2, -0.4]) * 14.0 + 3.0 * rng.standard_normal((mq.nnodes, 2))
    for cname, cls in (("ref", K.Temperature), ("pg", K.TemperaturePG), ("cons", TemperatureConsistent)):
        assemble_case(f"asm/quad_{cname}", mq, matq, vq, cls, transient=True)
    # scrambled connectivity (unstructured-style input)
    mqs = make_mesh("quad", ll=(0.0, 0.0), ur=(1.0, 1.0), nx=6, ny=5); distort(mqs, rng, 0.2); scramble(mqs, rng)
    vqs = 9.0 * rng.standard_normal((mqs.nnodes, 2))
    assemble_case("asm/quad_scrambled_pg", mqs, matq, vqs, K.TemperaturePG, transient=True)
    # zero velocity (alpha = 0 branch of PetrovGalerkinBases)
    assemble_case("asm/quad_zero_u_pg", mqs, matq, None, K.TemperaturePG, transient=False)

    ml = make_mesh("line", l=0.0, r=2.0, n=11); distort(ml, rng, 0.3)
    matl = K.FourierConduction(1, k=0.8, area=2.0, density=1.1, heat_capacity_v=1.4, q=-2.0)
    vl = 5.0 + rng.standard_normal((ml.nnodes, 1))
    for cname, cls in (("ref", K.Temperature), ("pg", K.TemperaturePG), ("cons", TemperatureConsistent)):
        assemble_case(f"asm/line_{cname}", ml, matl, vl, cls, transient=True)

    # ------------------------------------------------------------------ 4./5. Dirichlet reduction + full steady solves
    def nanflags(m):
        return np.zeros((m.nnodes, 1)) + np.nan

    def steady_case(name, m, material, velocity, cls, dflags, nflags=None):
        form = cls(m, velocity=velocity)
        bc = K.BoundaryCondition()
        bc.SetDirichletCriteria(lambda: dflags.copy())
        if nflags is not None:
            bc.SetNeumannCriteria(lambda: nflags.copy())
        fs = K.FEMSolver(analysis_type="steady")
        with quiet():
            sol = fs.Solve(formulation=form, mesh=m, material=material, boundary_condition=bc)
            # seam-level: reduced system exactly as SteadyLinearDiffusionSolver builds it
            Kc = Assemble(fs, form.function_spaces, form, m, material, bc)[0]
            Res = -bc.ComputeNeumannFluxes(m, material)
            K_b, F_b, F = bc.ApplyDirichletGetReducedMatrices(Kc, Res, bc.applied_dirichlet)
        out[f"{name}/points"] = m.points.copy(); out[f"{name}/elements"] = m.elements.astype(np.int64)
        out[f"{name}/velocity"] = form.velocity.copy()
        out[f"{name}/material"] = np.array([material.k, material.rho, material.c_v, material.area, material.q])
        out[f"{name}/dirichlet_flags"] = dflags
        if nflags is not None:
            out[f"{name}/neumann_flags"] = nflags
        out[f"{name}/columns_in"] = bc.columns_in.astype(np.int64)
        out[f"{name}/columns_out"] = bc.columns_out.astype(np.int64)
        out[f"{name}/Kb_data"] = K_b.data.copy(); out[f"{name}/Kb_indices"] = K_b.indices.copy()
        out[f"{name}/Kb_indptr"] = K_b.indptr.copy(); out[f"{name}/F_b"] = F_b.copy()
        out[f"{name}/sol"] = sol.sol.copy()
        return sol.sol

    # config 1: examples/steady_1d.py
    m1 = make_mesh("line", l=0.0, r=0.5, n=5)
    mat1 = K.FourierConduction(1, k=1.0e3, area=10.0e3, density=1.0, heat_capacity_v=1.0, q=0.0)
    d1 = nanflags(m1); d1[0, 0] = 100.0; d1[-1, 0] = 500.0
    u1 = np.zeros((m1.nnodes, 1)) + 4.0e7
    steady_case("steady/ex1d_galerkin", m1, mat1, u1, K.Temperature, d1)
    steady_case("steady/ex1d_pg", m1, mat1, u1, K.TemperaturePG, d1)
    # config 2: examples/steady_2d.py
    m2 = make_mesh("quad", ll=(0, 0), ur=(0.5, 0.01), nx=5, ny=3)
    mat2 = K.FourierConduction(2, k=1.0e3, area=1.0, density=1.0, c_v=1.0, q=0.0)
    d2 = nanflags(m2); d2[m2.points[:, 0] == 0.0, 0] = 100.0; d2[m2.points[:, 0] == 0.5, 0] = 500.0
    u2 = np.zeros((m2.nnodes, 2)); u2[:, 0] += 9.0e3
    steady_case("steady/ex2d_diffusion", m2, mat2, None, K.Temperature, d2)
    steady_case("steady/ex2d_ref_u", m2, mat2, u2, K.Temperature, d2)
    steady_case("steady/ex2d_pg", m2, mat2, u2, K.TemperaturePG, d2)
    steady_case("steady/ex2d_consistent", m2, mat2, u2, TemperatureConsistent, d2)
    # medium distorted mesh, mixed Dirichlet (one zero-valued -> np.isclose branch) + Neumann loads
    rng = np.random.default_rng(99)
    m3 = make_mesh("quad", ll=(0.0, 0.0), ur=(2.0, 1.0), nx=23, ny=13); distort(m3, rng, 0.2)
    mat3 = K.FourierConduction(2, k=1.5, area=1.0, density=1.2, heat_capacity_v=0.8, q=0.0)
    d3 = nanflags(m3)
    d3[m3.points[:, 0] == 0.0, 0] = 0.0          # exercised by ~np.isclose(T_D, 0)
    d3[m3.points[:, 0] == 2.0, 0] = 350.0
    d3[m3.points[:, 1] == 0.0, 0] = 5e-9         # |T_D| <= 1e-8: dropped from the RHS by np.isclose, still imposed
    n3 = nanflags(m3); top = np.where(m3.points[:, 1] == 1.0)[0]; n3[top[3:-3], 0] = 12.5
    u3 = np.array([6.0, 2.0]) + 0.6 * rng.standard_normal((m3.nnodes, 2))
    steady_case("steady/mixed_pg", m3, mat3, u3, K.TemperaturePG, d3, n3)
    steady_case("steady/mixed_consistent", m3, mat3, u3, TemperatureConsistent, d3, n3)
    steady_case("steady/mixed_ref", m3, mat3, u3, K.Temperature, d3, n3)
    ml3 = make_mesh("line", l=0.0, r=1.0, n=40); distort(ml3, rng, 0.3)
    matl3 = K.FourierConduction(1, k=0.5, area=1.0, density=1.0, heat_capacity_v=1.0, q=0.0)
    dl3 = nanflags(ml3); dl3[0, 0] = 10.0; dl3[-1, 0] = -4.0
    nl3 = nanflags(ml3); nl3[17, 0] = 3.0
    ul3 = np.zeros((ml3.nnodes, 1)) + 30.0
    steady_case("steady/line_consistent", ml3, matl3, ul3, TemperatureConsistent, dl3, nl3)
    steady_case("steady/line_pg", ml3, matl3, ul3, K.TemperaturePG, dl3, nl3)

    # ------------------------------------------------------------------ 6. transient restatement (FEMSolver.py:253-297, :138-145)
    def transient_case(name, m, material, velocity, cls, dflags, nflags, init, nincr, dt_factor):
        form = cls(m, velocity=velocity)
        bc = K.BoundaryCondition()
        bc.SetInitialConditions(lambda: init.copy())
        bc.SetDirichletCriteria(lambda: dflags.copy())
        if nflags is not None:
            bc.SetNeumannCriteria(lambda: nflags.copy())
        fs = K.FEMSolver(analysis_type="transient")
        with quiet():
            fs.ComputeSparsityFEM(m, form)
            bc.GetDirichletBoundaryConditions(form, m, material, fs)
            Fn = bc.ComputeNeumannFluxes(m, material)
            Kc, _, Mm = Assemble(fs, form.function_spaces, form, m, material, bc)
        # time step rule, FEMSolver.py:138-144 (factor 0.1 in the code; 0.25 reproduces the shipped 2D figure)
        dts = np.empty(m.nelem)
        for e in range(m.nelem):
            X = m.points[m.elements[e]]
            dts[e] = material.rho * material.c_v * np.linalg.norm(X[1] - X[0]) ** 2 / (2.0 * material.k)
        dt = dt_factor * dts.min()
        cin, cout, td = bc.columns_in, bc.columns_out, bc.applied_dirichlet
        invM = np.linalg.inv(Mm.toarray())                       # :262 (full M, dense)
        Kd = Kc.toarray()
        nzc = ~np.isclose(td, 0.0)
        # :270-272: NodalFluxes = -(0 - K[in,out] T_D) ; NodalFluxes += -F_neumann
        F = np.zeros((m.nnodes, 1)); F[cin, 0] = Kd[np.ix_(cin, cout[nzc])] @ td[nzc]; F -= Fn
        T = np.zeros((m.nnodes, nincr)); T[:, 0] = init[:, 0]
        T[cout, 0] += td                                           # :274 (added on top of the initial field)
        K_b = Kd[np.ix_(cin, cin)]; F_b = F[cin, 0]; invM_b = invM[np.ix_(cin, cin)]
        for n in range(nincr - 1):
            T[cout, n + 1] += td                                   # :277
            dTdt = K_b @ T[cin, n] + F_b                           # :281
            T[cin, n + 1] += T[cin, n] - dt * (invM_b @ dTdt)      # :283-285
        out[f"{name}/points"] = m.points.copy(); out[f"{name}/elements"] = m.elements.astype(np.int64)
        out[f"{name}/velocity"] = form.velocity.copy()
        out[f"{name}/material"] = np.array([material.k, material.rho, material.c_v, material.area, material.q])
        out[f"{name}/dirichlet_flags"] = dflags; out[f"{name}/initial"] = init
        if nflags is not None:
            out[f"{name}/neumann_flags"] = nflags
        out[f"{name}/dt"] = np.float64(dt); out[f"{name}/nincr"] = np.int64(nincr)
        out[f"{name}/T_all"] = T
        out[f"{name}/snapshots"] = T[:, [0, 1, int(nincr / 3), int(2 * nincr / 3), -1]]   # :297

    # config 3: examples/transient_2d.py (current code: factor 0.1) and the shipped-figure variant (0.25)
    mt = make_mesh("quad", ll=(0, 0), ur=(0.02, 1.0), nx=5, ny=3)
    matt = K.FourierConduction(2, k=10.0, area=1.0, density=10.0e6, heat_capacity_v=1.0)
    it = np.zeros((mt.nnodes, 1)) + 200.0
    dt_ = nanflags(mt); dt_[mt.points[:, 0] == 0.02, 0] = 0.0
    nt_ = nanflags(mt); nt_[mt.points[:, 0] == 0.0, 0] = 0.0
    transient_case("transient/ex2d_dt010", mt, matt, None, K.Temperature, dt_, nt_, it, 61, 0.1)
    transient_case("transient/ex2d_dt025", mt, matt, None, K.Temperature, dt_, nt_, it, 61, 0.25)
    mt1 = make_mesh("line", l=0.0, r=0.02, n=5)
    matt1 = K.FourierConduction(1, k=10.0, area=1.0, density=10.0e6, heat_capacity_v=1.0)
    it1 = np.zeros((mt1.nnodes, 1)) + 200.0
    dt1 = nanflags(mt1); dt1[-1, 0] = 0.0
    nt1 = nanflags(mt1); nt1[0, 0] = 0.0
    transient_case("transient/ex1d", mt1, matt1, None, K.Temperature, dt1, nt1, it1, 151, 0.1)
    # non-trivial: distorted mesh, nonzero Dirichlet, Neumann load, consistent advection
    rng = np.random.default_rng(5)
    mt3 = make_mesh("quad", ll=(0.0, 0.0), ur=(1.0, 0.6), nx=11, ny=7); distort(mt3, rng, 0.15)
    matt3 = K.FourierConduction(2, k=0.7, area=1.0, density=2.0, heat_capacity_v=1.5, q=0.0)
    it3 = 20.0 + 5.0 * rng.standard_normal((mt3.nnodes, 1))
    dt3 = nanflags(mt3); dt3[mt3.points[:, 0] == 0.0, 0] = 80.0; dt3[mt3.points[:, 0] == 1.0, 0] = 0.0
    nt3 = nanflags(mt3); nt3[np.where(mt3.points[:, 1] == 0.6)[0][2:-2], 0] = -1.5
    ut3 = np.array([0.8, 0.3]) + 0.05 * rng.standard_normal((mt3.nnodes, 2))
    transient_case("transient/distorted_consistent", mt3, matt3, ut3, TemperatureConsistent, dt3, nt3, it3, 41, 0.1)

    # ------------------------------------------------------------------ function-space tables
    fs2 = K.FunctionSpace(m2); fs1 = K.FunctionSpace(m1)
    out["fs/quad/Bases"] = fs2.Bases; out["fs/quad/gBases"] = fs2.gBases; out["fs/quad/AllGauss"] = fs2.AllGauss
    out["fs/line/Bases"] = fs1.Bases; out["fs/line/gBases"] = fs1.gBases; out["fs/line/AllGauss"] = fs1.AllGauss

    # group by top-level prefix -> one npz per group
    groups = {}
    for k, v in out.items():
        g, rest = k.split("/", 1)
        groups.setdefault(g, {})[rest.replace("/", "__")] = v
    for g, d in groups.items():
        path = os.path.join(HERE, f"{g}.npz")
        np.savez_compressed(path, **d)
        print(f"wrote {path}: {len(d)} arrays, {os.path.getsize(path)/1024:.1f} KiB")


if __name__ == "__main__":
    main()
