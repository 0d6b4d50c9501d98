"""Boundary conditions, mirroring kiltro/BoundaryCondition.py (reference file:line cited per method).

Flag arrays follow the reference's convention: (nnodes, nvar) float64, NaN = unconstrained.  They may be NumPy arrays
(uploaded once) or torch CUDA tensors (large synthetic cases).  Index sets (columns_in / columns_out / applied_dirichlet)
are computed on device by stream compaction and exposed as lazy NumPy attributes for drop-in compatibility."""
from warnings import warn

import numpy as np
import torch

from . import _lib, device as D
from .sparse import DeviceCSR


class BoundaryCondition(object):

    def __init__(self):                                          # BoundaryCondition.py:6-25
        self.analysis_type = "steady"
        self.initial_field = None
        self.dirichlet_data_applied_at = 'node'
        self.neumann_data_applied_at = 'node'
        self.dirichlet_flags = None
        self.neumann_flags = None
        self.applied_neumann = None
        self.robin_convection_flags = None
        self.applied_robin_convection = 0.0
        self.ground_robin_convection = 0.0
        # device state produced by GetDirichletBoundaryConditions
        self.renum = None                 # int32 (ndof): index in the reduced system, or -1-k for the k-th fixed DOF
        self.columns_in_device = None     # int64 (n_free)
        self.columns_out_device = None    # int64 (n_fixed)
        self.applied_dirichlet_device = None
        self.n_free = None
        self.n_fixed = None
        self._host = {}
        # cached structure (symbolic phases): split buffers, renumbering version, reduced pattern + row masks per CSR pattern
        self._split = None
        self.renum_version = 0
        self._reduce_cache = None

    # ------------------------------------------------------------------ user data (BoundaryCondition.py:27-63)
    def SetInitialConditions(self, func, *args, **kwargs):
        self.initial_field = func(*args, **kwargs)
        return self.initial_field

    def SetDirichletCriteria(self, func, *args, **kwargs):
        self.dirichlet_flags = func(*args, **kwargs)
        return self.dirichlet_flags

    def SetNeumannCriteria(self, func, *args, **kwargs):
        tups = func(*args, **kwargs)
        if not isinstance(tups, tuple) and self.neumann_data_applied_at == "node":
            self.neumann_flags = tups
            return self.neumann_flags
        if len(tups) != 2:
            raise ValueError("User-defined Neumann criterion function {} "
                             "should return one flag and one data array".format(func.__name__))
        self.neumann_flags = tups[0]
        self.applied_neumann = tups[1]
        return tups

    def SetRobinCriteria(self, func, *args, **kwargs):           # :66-97 stores data only (Robin is dead code upstream)
        dics = func(*args, **kwargs)
        if isinstance(dics, dict):
            self.RobinSelector(dics)
        elif isinstance(dics, tuple):
            for d in dics:
                if not isinstance(d, dict):
                    raise ValueError("User-defined Robin criterion function {} "
                                     "should return dictionary or tuple(dict,dict,...)".format(func.__name__))
                self.RobinSelector(d)
        else:
            raise ValueError("User-defined Robin criterion function {} "
                             "should return dictionary or tuple".format(func.__name__))
        return dics

    def RobinSelector(self, tups):
        if tups['type'] == 'Convection':
            self.robin_convection_flags = tups['flags']
            self.applied_robin_convection = tups['data']
        else:
            raise ValueError("Type force {} not understood or not available.".format(tups['type']))

    # ------------------------------------------------------------------ lazy host views of the index sets
    def _lazy(self, name, dev):
        if name not in self._host:
            self._host[name] = D.to_numpy(dev) if dev is not None else None
        return self._host[name]

    @property
    def columns_in(self):
        return self._lazy("in", self.columns_in_device)

    @property
    def columns_out(self):
        return self._lazy("out", self.columns_out_device)

    @property
    def applied_dirichlet(self):
        return self._lazy("td", self.applied_dirichlet_device)

    # ------------------------------------------------------------------ K4
    def GetDirichletBoundaryConditions(self, formulation, mesh, material=None, fem_solver=None):      # :100-122
        if self.dirichlet_flags is None:
            raise RuntimeError("Dirichlet boundary conditions are not set for the analysis")
        lib = _lib.load()
        flags = D.to_device(self.dirichlet_flags, torch.float64).reshape(-1)
        ndof = int(flags.shape[0])
        if ndof != formulation.nvar * mesh.nnodes:
            raise ValueError("dirichlet_flags has %d entries, expected nnodes*nvar = %d" % (ndof, formulation.nvar * mesh.nnodes))
        # buffers are kept across calls (load increments / repeated Solve() reuse them); the renumbering goes to the one of
        # two buffers the previous call did not use, so that the kernel can tell whether anything changed
        sp = self._split
        if sp is None or sp["ndof"] != ndof:
            sp = {"ndof": ndof, "renum": [D.empty(ndof, torch.int32), D.empty(ndof, torch.int32)], "cur": -1,
                  "cin": D.empty(ndof, torch.int64), "cout": D.empty(ndof, torch.int64), "td": D.empty(ndof),
                  "ws": D.empty(int(lib.kb_dirichlet_split_workspace_bytes(ndof)), torch.uint8)}
            self._split = sp
        prev = sp["renum"][sp["cur"]] if sp["cur"] >= 0 else None
        nxt = 0 if sp["cur"] != 0 else 1
        renum = sp["renum"][nxt]
        counts = (_lib.I64 * 3)()
        _lib.check(lib.kb_dirichlet_split(D.ptr(flags), ndof, D.ptr(prev), D.ptr(renum), D.ptr(sp["cin"]), D.ptr(sp["cout"]),
                                          D.ptr(sp["td"]), D.ptr(sp["ws"]), sp["ws"].numel(), counts, D.stream_ptr()))
        self.n_free, self.n_fixed = int(counts[0]), int(counts[1])
        if int(counts[2]) == 1 and self.renum is prev:
            pass                                                 # same Dirichlet set: renum (and what hangs on it) stays
        else:
            sp["cur"] = nxt
            self.renum = renum
            self.renum_version += 1
            self._reduce_cache = None
        self.columns_in_device = sp["cin"][:self.n_free]
        self.columns_out_device = sp["cout"][:self.n_fixed]
        self.applied_dirichlet_device = sp["td"][:self.n_fixed].clone()
        self._host = {}
        if self.n_free == 0:
            warn("Dirichlet boundary conditions have been applied on the entire mesh")
        if self.n_fixed == 0:
            warn("No Dirichlet boundary conditions have been applied. The system is unconstrained")

    def ComputeNeumannFluxes(self, mesh, material):              # :125-140  -> device (ndof, 1)
        ndof = mesh.nnodes * material.nvar
        if self.neumann_flags is None:
            return D.zeros((ndof, 1))
        lib = _lib.load()
        flags = D.to_device(self.neumann_flags, torch.float64).reshape(-1)
        F = D.empty((ndof, 1))
        _lib.check(lib.kb_neumann_fluxes(D.ptr(flags), ndof, D.ptr(F), D.stream_ptr()))
        return F

    def _applied(self, AppliedDirichlet):
        if AppliedDirichlet is None:
            return self.applied_dirichlet_device
        return D.to_device(AppliedDirichlet, torch.float64).reshape(-1)

    def ApplyDirichletGetReducedMatrices(self, K, F, AppliedDirichlet, LoadFactor=1., mass=None, only_residual=False):
        """BoundaryCondition.py:169-200.  K (and mass) are DeviceCSR, F a device (ndof,1)/(ndof,) tensor that is updated
        in place like the reference's.  Returns (K_b, F_b, F) [+ mass_b when self.analysis_type != 'steady']."""
        lib = _lib.load()
        if not isinstance(K, DeviceCSR):
            raise ValueError("K must be a kiltro_b200 DeviceCSR")
        ndof = K.shape[0]
        if not (isinstance(F, torch.Tensor) and F.is_cuda and F.dtype == torch.float64 and F.is_contiguous()):
            raise ValueError("F must be a contiguous float64 CUDA tensor (it is modified in place)")
        applied = self._applied(AppliedDirichlet)
        want_mass = mass is not None and not isinstance(mass, list)
        nfree = self.n_free
        # symbolic phase, cached per (CSR pattern, Dirichlet set): reduced pattern + one survival mask per row
        key = (K.indptr.data_ptr(), K.indices.data_ptr(), K.nnz, ndof, self.renum_version)
        rc = self._reduce_cache
        if rc is None or rc["key"] != key:
            ws = D.empty(int(lib.kb_dirichlet_reduce_workspace_bytes(ndof, nfree)), torch.uint8)
            mask = D.empty(max(ndof, 1), torch.int32)
            kb_ptr = D.empty(nfree + 1, torch.int32)
            kb_ind = D.empty(max(K.nnz, 1), torch.int32)
            h = (_lib.I64 * 2)()
            _lib.check(lib.kb_dirichlet_reduce_symbolic(ndof, D.ptr(K.indptr), D.ptr(K.indices), D.ptr(self.renum), nfree,
                                                        D.ptr(mask), D.ptr(kb_ptr), D.ptr(kb_ind), D.ptr(ws), ws.numel(), h,
                                                        D.stream_ptr()))
            nb = int(h[0])
            rc = {"key": key, "mask": mask, "kb_ptr": kb_ptr, "kb_ind": kb_ind[:nb].clone() if nb < K.nnz // 2 else kb_ind[:nb],
                  "nnz_b": nb, "masks_ok": bool(h[1]), "ws": ws, "indptr": K.indptr, "indices": K.indices}
            self._reduce_cache = rc
        nb = rc["nnz_b"]
        Fb = D.empty(nfree)
        if not rc["masks_ok"]:
            # rows longer than 32 entries: one-shot form (recomputes the structure, synchronises)
            kb_ind = D.empty(K.nnz, torch.int32); kb_dat = D.empty(K.nnz)
            kb_ptr = D.empty(nfree + 1, torch.int32)
            mb_dat = D.empty(K.nnz) if want_mass else None
            nnzb = _lib.I64(0)
            _lib.check(lib.kb_apply_dirichlet_reduce(ndof, D.ptr(K.indptr), D.ptr(K.indices), D.ptr(K.data),
                                                     D.ptr(mass.data) if want_mass else None, D.ptr(self.renum),
                                                     D.ptr(applied), float(LoadFactor), D.ptr(F), nfree, D.ptr(kb_ptr),
                                                     D.ptr(kb_ind), D.ptr(kb_dat), D.ptr(mb_dat), D.ptr(Fb), D.ptr(rc["ws"]),
                                                     rc["ws"].numel(), nnzb, D.stream_ptr()))
            if only_residual:
                return F
            K_b = DeviceCSR(kb_dat[:nb], kb_ind[:nb], kb_ptr, (nfree, nfree), symmetric=K.symmetric)
            if want_mass:
                return K_b, Fb, F, DeviceCSR(mb_dat[:nb], kb_ind[:nb], kb_ptr, (nfree, nfree), symmetric=True)
            return K_b, Fb, F
        kb_dat = None if only_residual else D.empty(nb)
        mb_dat = D.empty(nb) if (want_mass and not only_residual) else None
        _lib.check(lib.kb_dirichlet_reduce_numeric(ndof, D.ptr(K.indptr), D.ptr(K.indices), D.ptr(K.data),
                                                   D.ptr(mass.data) if mb_dat is not None else None, D.ptr(self.renum),
                                                   D.ptr(applied), float(LoadFactor), D.ptr(F), nfree, D.ptr(rc["mask"]),
                                                   D.ptr(rc["kb_ptr"]), D.ptr(kb_dat), D.ptr(mb_dat), D.ptr(Fb),
                                                   D.stream_ptr()))
        if only_residual:
            return F
        K_b = DeviceCSR(kb_dat, rc["kb_ind"], rc["kb_ptr"], (nfree, nfree), symmetric=K.symmetric)
        if want_mass:
            return K_b, Fb, F, DeviceCSR(mb_dat, rc["kb_ind"], rc["kb_ptr"], (nfree, nfree), symmetric=True)
        return K_b, Fb, F

    def GetReducedMatrices(self, K, F, M=None):                  # :156-166 (no RHS elimination)
        zero = torch.zeros_like(self.applied_dirichlet_device)
        F2 = F.clone()
        out = self.ApplyDirichletGetReducedMatrices(K, F2, zero, mass=M)
        if M is not None and not isinstance(M, list):
            return out[0], out[1], out[3]
        return out[0], out[1], np.array([])

    def _total(self, sol, applied, fsize, nvar):
        lib = _lib.load()
        T = D.empty(fsize)
        # keep every buffer referenced by a local until the launch is enqueued (a temporary freed mid-expression could be
        # handed out again by the caching allocator before the kernel reads it)
        sol_d = D.to_device(sol, torch.float64).reshape(-1) if sol is not None else None
        app_d = self._applied(applied) if applied is not False else None
        _lib.check(lib.kb_total_component_sol(fsize, D.ptr(self.renum), D.ptr(sol_d), D.ptr(app_d), D.ptr(T),
                                              D.stream_ptr()))
        return T.reshape(fsize // nvar, nvar)

    def UpdateFixDoFs(self, AppliedDirichletInc, fsize, nvar):   # :203-214
        return self._total(None, AppliedDirichletInc, fsize, nvar)

    def UpdateFreeDoFs(self, sol, fsize, nvar):                  # :217-228
        return self._total(sol, False, fsize, nvar)

    def TotalComponentSol(self, sol, AppliedDirichletInc, Iter, fsize, nvar):      # :231-241
        return self._total(sol, AppliedDirichletInc, fsize, nvar)
