"""K9: row-block (mesh-strip) partitioned solves across GPUs -- new capability (the reference is single-process; SURVEY
section 8e).  One process per GPU; torch.distributed (NCCL) carries the only two exchanges the path has:

  * halo rows of the SpMV input (one mesh row = nx_nodes doubles per neighbour per SpMV),
  * the scalars of the dot products (2 doubles per all-reduce).

Each rank owns a horizontal strip of mesh rows, generates / assembles its strip plus ONE halo row of nodes per side as an
ordinary local mesh (so the CSR rows of the nodes it owns are complete and assembly needs no communication), and iterates
on  A' = [Kff 0; 0 I]  (identity rows for prescribed DOFs: the iterates on the free DOFs are those of the reduced system
the single-GPU path solves).  The solver loops are written SPMD-over-a-list: with NCCL the list has one local part; with
`LocalComm` several parts live in one process on one device and the "collectives" are plain copies/sums -- that is how
the partitioned algorithm is tested against the single-GPU solve without a multi-GPU box, and the CPU (gloo) tests cover
the partition arithmetic and the exchange plumbing."""
import sys

import numpy as np
import torch

from . import _lib, device as D
from .partition import StripPartition
from .MeshGeneration import Mesh
from .BoundaryCondition import BoundaryCondition
from .FEMSolver import FEMSolver
from .FiniteElements.Assembly import Assemble


# Hide the halo exchange behind the interior rows of the SpMV (3 launches instead of 1).  Measured on 2 B200s at 16 M DOF
# per GPU: no gain (the exchange is short compared with the launch overhead it adds), so it is off by default.
OVERLAP_HALO = False
USE_ROW_PATTERNS = True       # row-pattern form of the SpMV (csrc/spmv_pat.cuh) when the local matrix allows it


# ------------------------------------------------------------------------------------------------------ communicators
class DistComm(object):
    """torch.distributed (NCCL on GPUs; gloo works for CPU tensors in the host-logic tests).  One part per process."""

    def __init__(self, group=None):
        import torch.distributed as dist
        self.dist, self.group = dist, group
        self.world, self.rank = dist.get_world_size(group), dist.get_rank(group)

    def halo_exchange(self, parts, vecs):
        part, v = parts[0], vecs[0]
        dist = self.dist
        ops, keep = [], []
        if part.ha:
            sb = v[part.last_own_row()].contiguous(); rb = torch.empty_like(sb); keep.append((rb, part.halo_above()))
            ops += [dist.P2POp(dist.isend, sb, part.rank + 1, self.group), dist.P2POp(dist.irecv, rb, part.rank + 1, self.group)]
        if part.hb:
            sb = v[part.first_own_row()].contiguous(); rb = torch.empty_like(sb); keep.append((rb, part.halo_below()))
            ops += [dist.P2POp(dist.isend, sb, part.rank - 1, self.group), dist.P2POp(dist.irecv, rb, part.rank - 1, self.group)]
        if ops:
            for r in dist.batch_isend_irecv(ops):
                r.wait()
            for rb, sl in keep:
                v[sl] = rb

    def halo_start(self, parts, vecs):
        """Posts the sends/receives of the boundary rows; returns a handle for halo_finish (the transfers run on NCCL's
        stream while the caller computes the rows that do not need the halo)."""
        part, v = parts[0], vecs[0]
        dist = self.dist
        ops, keep = [], []
        if part.ha:
            sb = v[part.last_own_row()].contiguous(); rb = torch.empty_like(sb); keep.append((rb, part.halo_above()))
            ops += [dist.P2POp(dist.isend, sb, part.rank + 1, self.group), dist.P2POp(dist.irecv, rb, part.rank + 1, self.group)]
        if part.hb:
            sb = v[part.first_own_row()].contiguous(); rb = torch.empty_like(sb); keep.append((rb, part.halo_below()))
            ops += [dist.P2POp(dist.isend, sb, part.rank - 1, self.group), dist.P2POp(dist.irecv, rb, part.rank - 1, self.group)]
        works = dist.batch_isend_irecv(ops) if ops else []
        return (works, keep, v)

    def halo_finish(self, handle):
        works, keep, v = handle
        for r in works:
            r.wait()
        for rb, sl in keep:
            v[sl] = rb

    def allreduce(self, outs):
        self.dist.all_reduce(outs[0], op=self.dist.ReduceOp.SUM, group=self.group)

    def barrier(self):
        self.dist.barrier(self.group)


class LocalComm(object):
    """All parts in one process (one device): exchanges are copies, reductions are sums in rank order."""

    def __init__(self, world):
        self.world, self.rank = world, 0

    def halo_exchange(self, parts, vecs):
        for k, part in enumerate(parts):
            if part.ha:
                vecs[k][part.halo_above()] = vecs[k + 1][parts[k + 1].first_own_row()]
            if part.hb:
                vecs[k][part.halo_below()] = vecs[k - 1][parts[k - 1].last_own_row()]

    def halo_start(self, parts, vecs):
        return (parts, vecs)

    def halo_finish(self, handle):
        self.halo_exchange(*handle)

    def allreduce(self, outs):
        total = outs[0].clone()
        for o in outs[1:]:
            total += o
        for o in outs:
            o.copy_(total)

    def barrier(self):
        pass


# ------------------------------------------------------------------------------------------------------ local systems
class LocalSystem(object):
    """The strip of one rank: local mesh, assembled operators (owned rows complete), right-hand side, work vectors."""

    def __init__(self, part):
        self.part = part
        self.lib = _lib.load()
        self.cols16 = None
        self.row_pat = None             # row-pattern form of the SpMV (csrc/spmv_pat.cuh): one byte per local row
        self.pat_tab = None

    def pat_ptr(self, row0):
        """Pattern bytes of a row view starting at local row `row0` (NULL when the form is not in use)."""
        return None if self.row_pat is None else self.row_pat.data_ptr() + row0

    def _view(self, A_vals):
        """Row-range view (owned rows) of the local CSR for the SpMV kernel."""
        return A_vals

    def xptr(self, x, row0):
        """x argument of the SpMV kernels for a row view starting at local row `row0`: with 16-bit column offsets the
        kernel addresses x relative to the first row of the view."""
        return x.data_ptr() + (8 * row0 if (self.cols16 is not None or self.row_pat is not None) else 0)

    def spmv_dots(self, vals, x, y, w, out2):
        p = self.part
        _lib.check(self.lib.kb_spmv_dots(p.n_own, self.nnz_own, self.indptr.data_ptr() + 4 * p.own_off,
                                         D.ptr(self.indices), D.ptr(self.cols16), self.pat_ptr(p.own_off), D.ptr(self.pat_tab),
                                         D.ptr(vals), D.ptr(self.rowblk),
                                         self.xptr(x, p.own_off), y.data_ptr() + 8 * p.own_off,
                                         w.data_ptr() + 8 * p.own_off, D.ptr(out2), D.ptr(self.scratch), D.stream_ptr()))

    def spmv_dots4(self, vals, x, y, w, w2, out4):
        p = self.part
        _lib.check(self.lib.kb_spmv_dots4(p.n_own, self.nnz_own, self.indptr.data_ptr() + 4 * p.own_off,
                                          D.ptr(self.indices), D.ptr(self.cols16), self.pat_ptr(p.own_off), D.ptr(self.pat_tab),
                                          D.ptr(vals), D.ptr(self.rowblk),
                                          self.xptr(x, p.own_off), y.data_ptr() + 8 * p.own_off,
                                          w.data_ptr() + 8 * p.own_off, w2.data_ptr() + 8 * p.own_off, D.ptr(out4),
                                          D.ptr(self.scratch), D.stream_ptr()))

    def spmv(self, vals, x, y):
        p = self.part
        _lib.check(self.lib.kb_spmv(p.n_own, self.nnz_own, self.indptr.data_ptr() + 4 * p.own_off, D.ptr(self.indices),
                                    D.ptr(vals), D.ptr(self.rowblk), D.ptr(x), y.data_ptr() + 8 * p.own_off,
                                    D.stream_ptr()))

    def own_ptr(self, v):
        return v.data_ptr() + 8 * self.part.own_off

    # ---- overlap of the halo exchange with the rows that do not need it
    def build_ranges(self):
        """Splits the owned rows into [first mesh row | interior | last mesh row]; the interior rows only reference owned
        columns, so their SpMV can run while the halo rows are in flight."""
        p, lib = self.part, self.lib
        self.ranges = None
        nrows_mesh = p.j1 - p.j0
        if nrows_mesh < 3:
            return
        cuts = [p.own_off, p.own_off + p.nxn, p.own_off + p.n_own - p.nxn, p.own_off + p.n_own]
        ip = self.indptr[cuts].cpu().tolist()
        self.ranges = []
        for a in range(3):
            r0, nr, nnz = cuts[a], cuts[a + 1] - cuts[a], ip[a + 1] - ip[a]
            rb = D.empty(lib.kb_spmv_num_blocks(nnz) + 1, torch.int32)
            _lib.check(lib.kb_spmv_partition(nr, nnz, self.indptr.data_ptr() + 4 * r0, D.ptr(rb), D.stream_ptr()))
            self.ranges.append((r0, nr, nnz, rb))
        self.dparts = [D.zeros(4) for _ in range(3)]

    def _range_spmv(self, k, vals, x, y, w, w2, out):
        r0, nr, nnz, rb = self.ranges[k]
        ip = self.indptr.data_ptr() + 4 * r0
        yp, wp = y.data_ptr() + 8 * r0, w.data_ptr() + 8 * r0
        if w2 is None:
            _lib.check(self.lib.kb_spmv_dots(nr, nnz, ip, D.ptr(self.indices), D.ptr(self.cols16), self.pat_ptr(r0),
                                             D.ptr(self.pat_tab), D.ptr(vals), D.ptr(rb),
                                             self.xptr(x, r0), yp, wp, D.ptr(out), D.ptr(self.scratch), D.stream_ptr()))
        else:
            _lib.check(self.lib.kb_spmv_dots4(nr, nnz, ip, D.ptr(self.indices), D.ptr(self.cols16), self.pat_ptr(r0),
                                              D.ptr(self.pat_tab), D.ptr(vals), D.ptr(rb),
                                              self.xptr(x, r0), yp, wp, w2.data_ptr() + 8 * r0, D.ptr(out),
                                              D.ptr(self.scratch), D.stream_ptr()))


def spmv_overlapped(systems, comm, vals, x, y, w, w2, out):
    """y = A x on the owned rows of every part with the halo exchange of x hidden behind the interior rows; out[k] gets the
    local partial dots (2 values, or 4 when w2 is given)."""
    parts = [s.part for s in systems]
    nd = 2 if w2 is None else 4
    if not OVERLAP_HALO or any(s.ranges is None for s in systems):
        comm.halo_exchange(parts, x)
        for k, s in enumerate(systems):
            if w2 is None:
                s.spmv_dots(vals[k], x[k], y[k], w[k], out[k][0:2])
            else:
                s.spmv_dots4(vals[k], x[k], y[k], w[k], w2[k], out[k])
        return
    h = comm.halo_start(parts, x)
    for k, s in enumerate(systems):
        s._range_spmv(1, vals[k], x[k], y[k], w[k], None if w2 is None else w2[k], s.dparts[1])
    comm.halo_finish(h)
    for k, s in enumerate(systems):
        s._range_spmv(0, vals[k], x[k], y[k], w[k], None if w2 is None else w2[k], s.dparts[0])
        s._range_spmv(2, vals[k], x[k], y[k], w[k], None if w2 is None else w2[k], s.dparts[2])
        torch.add(s.dparts[0][0:nd], s.dparts[1][0:nd], out=out[k][0:nd])
        out[k][0:nd].add_(s.dparts[2][0:nd])


def local_system_from(mesh, form, bc, fs, material, transient=False, initial_fn=None):
    """LocalSystem of one rank from the ordinary API objects: `mesh` made with Mesh.Rectangle(..., partition=...) (or
    set_arrays(..., partition=...)), `form` / `bc` holding the rank's LOCAL nodal velocity / flags.  Not yet assembled:
    follow with reassemble()."""
    part = mesh.partition
    if part is None:
        raise ValueError("the mesh carries no StripPartition (Mesh.Rectangle(..., partition=StripPartition(...)))")
    if mesh.nnodes != part.n_loc:
        raise ValueError("the mesh has %d nodes, the partition expects %d (owned rows + halo rows)" % (mesh.nnodes, part.n_loc))
    sysm = LocalSystem(part)
    sysm.rowblk = None
    attach_objects(sysm, mesh, form, bc, fs, material, transient, initial_fn)
    return sysm


def attach_objects(sysm, mesh, form, bc, fs, material, transient=False, initial_fn=None):
    """(Re)binds the API objects of a LocalSystem; the symbolic state (SpMV row partition, index forms, peer mappings, the
    sparsity pattern cached on fs) is kept -- like the reference's pattern cache, it assumes the same mesh topology."""
    if mesh.structured_shape is None and getattr(mesh, "partition", None) is not None:
        mesh.structured_shape = (sysm.part.nxn, sysm.part.nrows_loc)
    sysm.mesh, sysm.form, sysm.bc, sysm.fs, sysm.material = mesh, form, bc, fs, material
    sysm.transient, sysm.initial_fn = transient, initial_fn
    sysm.velocity = form.device_velocity()
    sysm.flags = D.to_device(bc.dirichlet_flags, torch.float64) if bc.dirichlet_flags is not None else None
    sysm.symmetric = bool(form.is_symmetric)
    return sysm


def build_local_system(part, rectangle, material, formulation_cls, velocity_fn, dirichlet_fn, neumann_fn=None,
                       transient=False, initial_fn=None):
    """Generates the rank's strip (+ halo rows) of Mesh.Rectangle(*rectangle) on device, assembles it with the ordinary
    single-GPU machinery and prepares the partitioned operators.
    velocity_fn / dirichlet_fn / neumann_fn / initial_fn: callables (points (n_loc, 2) CUDA tensor) -> CUDA tensor
    ((n_loc, 2) velocity or None; (n_loc, 1) flags with NaN = unconstrained; (n_loc, 1) initial field)."""
    (x0, y0), (x1, y1), nx, ny = rectangle
    mesh = Mesh(element_type="quad")
    mesh.Rectangle((x0, y0), (x1, y1), nx, ny, partition=part)
    pts = mesh.device_points()
    vel = velocity_fn(pts) if velocity_fn is not None else None
    form = formulation_cls(mesh, velocity=vel) if vel is not None else formulation_cls(mesh)
    fs = FEMSolver(analysis_type="transient" if transient else "steady", verbose=False)
    bc = BoundaryCondition()
    flags = dirichlet_fn(pts)
    bc.SetDirichletCriteria(lambda: flags)
    if neumann_fn is not None:
        nfl = neumann_fn(pts)
        bc.SetNeumannCriteria(lambda: nfl)
    sysm = local_system_from(mesh, form, bc, fs, material, transient, initial_fn)
    reassemble(sysm)
    return sysm


def host_inputs(sysm):
    """Pinned host copies of the rank's inputs in the reference's array formats (float64 points / velocity / flags, int64
    connectivity): what a caller holding NumPy-side data would hand over.  Returns (points, elements, velocity, flags)."""
    pin = lambda t: t.cpu().pin_memory()
    return (pin(sysm.mesh.device_points()), pin(sysm.mesh.device_elements().to(torch.int64)),
            pin(sysm.velocity) if sysm.velocity is not None else None, pin(sysm.flags))


def refresh_inputs_from_host(sysm, h_points, h_elements, h_velocity, h_flags):
    """Host -> device copy of every input of the rank's numeric phase into the buffers the cached symbolic phase refers
    to (connectivity narrowed int64 -> int32 on device, like Mesh.set_arrays); follow with reassemble().  Returns the
    number of bytes copied."""
    n = 0
    sysm.mesh.device_points().copy_(h_points, non_blocking=True); n += h_points.numel() * 8
    el64 = h_elements.to("cuda", non_blocking=True); n += h_elements.numel() * 8
    sysm.mesh.device_elements().copy_(el64)
    if h_velocity is not None:
        sysm.velocity.copy_(h_velocity, non_blocking=True); n += h_velocity.numel() * 8
    sysm.flags.copy_(h_flags, non_blocking=True); n += h_flags.numel() * 8
    return n


def reassemble(sysm):
    """Numeric phase of one rank (pattern / tile plan are cached on sysm.fs): assembly, Dirichlet/Neumann, right-hand side."""
    lib, part = sysm.lib, sysm.part
    mesh, form, bc, fs, material = sysm.mesh, sysm.form, sysm.bc, sysm.fs, sysm.material
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
    sysm.events = ev
    ev[0].record()
    fs.ComputeSparsityFEM(mesh, form)
    K, Flux, M = Assemble(fs, form.function_spaces, form, mesh, material, bc)
    ev[1].record()
    bc.GetDirichletBoundaryConditions(form, mesh, material, fs)
    Fn = bc.ComputeNeumannFluxes(mesh, material)
    sysm.K, sysm.M = K, M
    sysm.indptr, sysm.indices = K.indptr, K.indices
    if sysm.rowblk is None:                                   # symbolic: SpMV row partition of the owned-row view
        ip = K.indptr[[part.own_off, part.own_off + part.n_own]].cpu()
        sysm.nnz_own = int(ip[1] - ip[0])
        nblk = lib.kb_spmv_num_blocks(sysm.nnz_own)
        sysm.rowblk = D.empty(nblk + 1, torch.int32)
        _lib.check(lib.kb_spmv_partition(part.n_own, sysm.nnz_own, K.indptr.data_ptr() + 4 * part.own_off,
                                         D.ptr(sysm.rowblk), D.stream_ptr()))
        sysm.scratch = D.zeros(lib.kb_blocks_scratch_bytes(sysm.nnz_own), torch.uint8)
        sysm.build_ranges()
        # 16-bit column offsets (10 instead of 12 bytes per non-zero in the SpMV) when the local band fits
        c16 = D.empty(K.nnz + 8, torch.int16); flag = D.zeros(1, torch.int32)
        _lib.check(lib.kb_make_cols16(part.n_loc, D.ptr(K.indptr), D.ptr(K.indices), D.ptr(c16), D.ptr(flag), D.stream_ptr()))
        sysm.cols16 = c16 if (int(flag.item()) == 0 and part.nxn % 4 == 0 and K.indptr.data_ptr() % 16 == 0) else None
        # row patterns (8 bytes per non-zero: no column index is streamed) when the local matrix has <= 256 of them; the
        # row views start at multiples of nxn, which must keep the pattern bytes 16-byte aligned
        if sysm.cols16 is not None and part.nxn % 16 == 0 and USE_ROW_PATTERNS:
            rp = D.empty((part.n_loc + 15) // 16 * 16 + 16, torch.uint8)
            keys = D.empty(256 * 8, torch.uint8); tab = D.empty(256 * 16, torch.int16); pflag = D.zeros(1, torch.int32)
            _lib.check(lib.kb_make_row_patterns(part.n_loc, D.ptr(K.indptr), D.ptr(K.indices), D.ptr(rp), D.ptr(keys),
                                                D.ptr(tab), D.ptr(pflag), D.stream_ptr()))
            if int(pflag.item()) == 0:
                sysm.row_pat, sysm.pat_tab = rp, tab
    n = part.n_loc
    # right-hand side of the steady problem on A' = [Kff 0; 0 I]:  F - K[:,fixed] T_D at free rows, T_D at fixed rows
    F = (-Fn).contiguous()
    K_b, F_b, F = bc.ApplyDirichletGetReducedMatrices(K, F, bc.applied_dirichlet_device)
    sysm.b = bc.TotalComponentSol(F_b, bc.applied_dirichlet_device, 0, n, 1).reshape(-1).contiguous()
    del K_b
    sysm.Fm = None
    if sysm.transient:
        # F = K[in,out] T_D - F_neumann at free DOFs, 0 at fixed DOFs (FEMSolver.py:269-272)
        F2 = D.zeros((n, 1))
        bc.ApplyDirichletGetReducedMatrices(K, F2, bc.applied_dirichlet_device, only_residual=True)
        F2 = (-F2 - Fn).contiguous()
        sysm.Fm = D.empty(n)
        _lib.check(lib.kb_mask_free(n, D.ptr(bc.renum), D.ptr(F2), D.ptr(sysm.Fm), D.stream_ptr()))
        T0 = None
        if sysm.initial_fn is not None:
            T0 = sysm.initial_fn(mesh.device_points()).reshape(-1).to(torch.float64).clone()
        elif bc.initial_field is not None:
            T0 = D.to_device(bc.initial_field, torch.float64).reshape(-1).clone()
        if T0 is not None:
            T0 += bc.UpdateFixDoFs(bc.applied_dirichlet_device, n, 1).reshape(-1)
            sysm.T0 = T0
    ev[2].record()
    return sysm


def _masked(sysm, c_k, c_m, identity_fixed):
    lib = sysm.lib
    out = D.empty(sysm.K.nnz)
    _lib.check(lib.kb_masked_matrix(sysm.part.n_loc, sysm.K.nnz, D.ptr(sysm.indptr), D.ptr(sysm.indices),
                                    D.ptr(sysm.K.data), D.ptr(sysm.M.data) if (c_m != 0.0 and sysm.M != []) else None,
                                    D.ptr(sysm.bc.renum), float(c_k), float(c_m), int(identity_fixed), D.ptr(out),
                                    D.stream_ptr()))
    return out


def _jacobi_scale(systems, comm, mats, mode):
    """scale vectors (halo-consistent) and scaled copies of the local matrices.  mode 0: A D^-1, mode 1: D^-1/2 A D^-1/2."""
    scales, scaled = [], []
    for s, A in zip(systems, mats):
        sc = D.empty(s.part.n_loc)
        _lib.check(s.lib.kb_diag_scale(s.part.n_loc, D.ptr(s.indptr), D.ptr(s.indices), D.ptr(A), mode, D.ptr(sc),
                                       D.stream_ptr()))
        scales.append(sc)
    comm.halo_exchange([s.part for s in systems], scales)      # halo rows are incomplete locally: take the owner's value
    for s, A, sc in zip(systems, mats, scales):
        out = D.empty(A.shape[0])
        _lib.check(s.lib.kb_scale_matrix(s.part.n_loc, D.ptr(s.indptr), D.ptr(s.indices), D.ptr(A), mode, D.ptr(sc),
                                         D.ptr(out), D.stream_ptr()))
        scaled.append(out)
    return scales, scaled


# ------------------------------------------------------------------------------------------------------ Krylov loops
class _BestIterate(object):
    """Copy of the iterate with the smallest recurrence residual seen at a convergence check (one vector copy every
    check_every iterations).  The partitioned loops look at the residual only then, so a breakdown is noticed up to
    check_every iterations late, on an iterate that may already be wrecked: the loops hand back this copy instead and
    solve_steady restarts from it (the single-GPU driver freezes the iterate on device before the bad update).  The
    residual norms are all-reduced, so every rank takes the same decisions."""

    def __init__(self, x, rr, enabled=True):
        self.x = [xk.clone() for xk in x] if enabled else None
        self.rr = rr

    def offer(self, x, rr):
        if self.x is not None and rr < self.rr:
            for keep, xk in zip(self.x, x):
                keep.copy_(xk)
            self.rr = rr

    def restore(self, x):
        if self.x is None:
            return float("nan")
        for keep, xk in zip(self.x, x):
            xk.copy_(keep)
        return self.rr


def dist_bicgstab(systems, comm, Ahat, b, x, rtol=1e-10, maxiter=20000, check_every=50, checkpoint=True):
    """Unpreconditioned BiCGSTAB on the (already Jacobi-scaled) partitioned operator, in the merged-update form of the
    single-GPU solver (csrc/krylov.cu): two all-reduces per iteration (after each SpMV; 2 and 4 doubles), the
    residual norm is reduced only when convergence is checked.  b, x: full local vectors per part (x: initial guess in,
    solution out, owned range).  Returns dict(iterations, relative_residual, status)."""
    parts = [s.part for s in systems]
    P = len(systems)
    z = lambda s: D.zeros(s.part.n_loc)
    r = [z(s) for s in systems]; rhat = [z(s) for s in systems]; p = [z(s) for s in systems]
    v = [z(s) for s in systems]; t = [z(s) for s in systems]
    d = [D.zeros(4) for _ in systems]
    dr = [D.zeros(2) for _ in systems]
    sc = [dict(rho=[D.zeros(1), D.zeros(1)], alpha=D.zeros(1)) for _ in systems]      # rho is double-buffered
    cur = 0
    st = D.stream_ptr
    # r = b - A x ; rhat = p = r
    comm.halo_exchange(parts, x)
    for k, s in enumerate(systems):
        s.spmv(Ahat[k], x[k], v[k])
        _lib.check(s.lib.kb_residual(s.part.n_own, s.own_ptr(b[k]), s.own_ptr(v[k]), s.own_ptr(r[k]), D.ptr(dr[k]),
                                     D.ptr(s.scratch), st()))
    comm.allreduce(dr)
    rr0, bb = float(dr[0][0]), float(dr[0][1])
    tol2 = rtol * rtol * bb
    for k in range(P):
        rhat[k].copy_(r[k]); p[k].copy_(r[k]); sc[k]["rho"][0].copy_(dr[k][0:1])
    it, rr, status = 0, rr0, 0
    if not (rr0 > tol2):
        return {"iterations": 0, "relative_residual": (rr0 / bb) ** 0.5 if bb > 0 else 0.0, "status": 0}
    keep = _BestIterate(x, rr0, checkpoint)
    import os, time
    dbg = os.environ.get("KB_DIST_DEBUG") == "1" and comm.rank == 0
    while it < maxiter:
        if dbg and it == check_every:
            torch.cuda.synchronize(); t_enq0 = time.perf_counter()
        for _ in range(min(check_every, maxiter - it)):
            spmv_overlapped(systems, comm, Ahat, p, v, rhat, None, d)
            comm.allreduce([dk[0:2] for dk in d])
            for k, s in enumerate(systems):                                                # alpha = rho/<rhat,v> ; s = r - alpha v
                _lib.check(s.lib.kb_bicgstab_s2(s.part.n_own, D.ptr(sc[k]["rho"][cur]), D.ptr(d[k]), s.own_ptr(v[k]),
                                                s.own_ptr(r[k]), D.ptr(sc[k]["alpha"]), st()))
            spmv_overlapped(systems, comm, Ahat, r, t, r, rhat, d)                        # {<s,t>,<t,t>,<rhat,t>,<rhat,s>}
            comm.allreduce(d)
            for k, s in enumerate(systems):                                                # omega, rho_new, beta inside the kernel
                _lib.check(s.lib.kb_bicgstab_update_p2(s.part.n_own, D.ptr(sc[k]["rho"][cur]), D.ptr(sc[k]["rho"][cur ^ 1]),
                                                       D.ptr(sc[k]["alpha"]), D.ptr(d[k]), s.own_ptr(t[k]), s.own_ptr(v[k]),
                                                       s.own_ptr(x[k]), s.own_ptr(r[k]), s.own_ptr(p[k]), D.ptr(dr[k]),
                                                       D.ptr(s.scratch), st()))
            cur ^= 1
            it += 1
        if dbg and it == 2 * check_every:
            t_enq1 = time.perf_counter(); torch.cuda.synchronize(); t_done = time.perf_counter()
            print("[dist debug] per iteration: CPU enqueue %.1f us, GPU-complete %.1f us" % (
                (t_enq1 - t_enq0) / check_every * 1e6, (t_done - t_enq0) / check_every * 1e6), file=sys.stderr, flush=True)
        comm.allreduce(dr)                                     # ||r||^2 only when convergence is looked at
        rr = float(dr[0][0])                                   # host sync every check_every iterations
        if not np.isfinite(rr) or rr > 1.0e10 * rr0:           # non-finite, or diverged 1e5-fold (same bound as krylov.cu)
            status = _lib.KB_EBREAKDOWN
            rr = keep.restore(x)                               # the caller restarts from the best checked iterate
            break
        keep.offer(x, rr)
        if rr <= tol2:
            break
    else:
        status = _lib.KB_ENOTCONV
    if status == 0 and rr > tol2:
        status = _lib.KB_ENOTCONV
    return {"iterations": it, "relative_residual": (rr / bb) ** 0.5 if bb > 0 else 0.0, "status": status}


def dist_cg(systems, comm, Ahat, b, x, rtol=1e-10, maxiter=20000, check_every=50, checkpoint=True):
    """CG on the (already symmetrically Jacobi-scaled) partitioned operator."""
    parts = [s.part for s in systems]
    P = len(systems)
    z = lambda s: D.zeros(s.part.n_loc)
    r = [z(s) for s in systems]; p = [z(s) for s in systems]; q = [z(s) for s in systems]
    d = [D.zeros(2) for _ in systems]
    dr = [D.zeros(2) for _ in systems]
    sc = [dict(rho=[D.zeros(1), D.zeros(1)]) for _ in systems]                          # rho is double-buffered
    cur = 0
    st = D.stream_ptr
    comm.halo_exchange(parts, x)
    for k, s in enumerate(systems):
        s.spmv(Ahat[k], x[k], q[k])
        _lib.check(s.lib.kb_residual(s.part.n_own, s.own_ptr(b[k]), s.own_ptr(q[k]), s.own_ptr(r[k]), D.ptr(d[k]),
                                     D.ptr(s.scratch), st()))
    comm.allreduce(d)
    rr, bb = float(d[0][0]), float(d[0][1])
    rr0 = rr
    tol2 = rtol * rtol * bb
    for k in range(P):
        p[k].copy_(r[k]); sc[k]["rho"][0].copy_(d[k][0:1])
    it, status = 0, 0
    if not (rr > tol2):
        return {"iterations": 0, "relative_residual": (rr / bb) ** 0.5 if bb > 0 else 0.0, "status": 0}
    keep = _BestIterate(x, rr0, checkpoint)
    while it < maxiter:
        for _ in range(min(check_every, maxiter - it)):
            spmv_overlapped(systems, comm, Ahat, p, q, p, None, d)
            comm.allreduce(d)
            for k, s in enumerate(systems):                                                # alpha = rho/<p,q>
                _lib.check(s.lib.kb_cg_update2(s.part.n_own, D.ptr(sc[k]["rho"][cur]), D.ptr(d[k]), s.own_ptr(p[k]),
                                               s.own_ptr(q[k]), s.own_ptr(x[k]), s.own_ptr(r[k]), D.ptr(dr[k]),
                                               D.ptr(s.scratch), st()))
            comm.allreduce(dr)
            for k, s in enumerate(systems):                                                # beta = <r,r>/rho ; rho = <r,r>
                _lib.check(s.lib.kb_cg_p2(s.part.n_own, D.ptr(sc[k]["rho"][cur]), D.ptr(sc[k]["rho"][cur ^ 1]), D.ptr(dr[k]),
                                          s.own_ptr(r[k]), s.own_ptr(p[k]), st()))
            cur ^= 1
            it += 1
        rr = float(dr[0][0])
        if not np.isfinite(rr) or rr > 1.0e10 * rr0:           # non-finite, or diverged 1e5-fold (same bound as krylov.cu)
            status = _lib.KB_EBREAKDOWN
            rr = keep.restore(x)
            break
        keep.offer(x, rr)
        if rr <= tol2:
            break
    else:
        status = _lib.KB_ENOTCONV
    if status == 0 and rr > tol2:
        status = _lib.KB_ENOTCONV
    return {"iterations": it, "relative_residual": (rr / bb) ** 0.5 if bb > 0 else 0.0, "status": status}


# ------------------------------------------------------------------------------------------------------ peer-memory path
class _RawCuda(object):
    """Exposes a raw device pointer through __cuda_array_interface__ so torch can view IPC memory without copying."""

    def __init__(self, ptr, n):
        self.__cuda_array_interface__ = {"shape": (int(n),), "typestr": "<f8", "data": (int(ptr), False), "version": 2}


def _align256(x):
    return (x + 255) // 256 * 256


class PeerRegion(object):
    """Peer-mapped memory of one rank -- control block + the three exchanged vectors (p, r, T: full local vectors) -- and
    the mappings of every other rank's region: what the native partitioned solvers (csrc/krylov_peer.cuh) exchange
    through.  Two ways to get one:
      * PeerRegion.create_ipc(part, dist): one process per GPU with NVLink peer access (a B200 HGX node); the collective
        backend only carries the 64-byte IPC handles, once;
      * PeerRegion.create_local(parts): all ranks in ONE process on ONE device (plain device memory), driven by one host
        thread and one stream per rank -- how the kernels' protocol is tested on a single-GPU box.
    self.part is the kb_peer_part the C entry points take; its event / halo sequence numbers run on across calls."""
    CTRL_BYTES = 1024
    NVEC = 3

    def __init__(self, part, bases, n_locs, own_offs, n_owns, owner):
        self.spart, self.lib = part, _lib.load()
        P, rank = part.world, part.rank
        if P > 8:
            raise ValueError("the peer-memory path supports up to 8 ranks")
        self.base, self._owner = bases, owner
        stride = [_align256(n * 8) for n in n_locs]
        self.part = _lib.kb_peer_part()
        kp = self.part
        kp.ctx.nranks, kp.ctx.rank = P, rank
        for r in range(P):
            kp.ctx.ctrl[r] = bases[r]
        kp.n_loc, kp.n_own, kp.own_off, kp.nxn = part.n_loc, part.n_own, part.own_off, part.nxn
        kp.ev = 0

        def slot(r, k, offset_elems):
            return bases[r] + self.CTRL_BYTES + k * stride[r] + 8 * offset_elems

        self.vec = []
        for k in range(self.NVEC):
            kp.d_vec[k] = slot(rank, k, 0)
            # halo-above slot of rank-1 / halo-below slot of rank+1 for vector k
            kp.d_dn_dst[k] = slot(rank - 1, k, own_offs[rank - 1] + n_owns[rank - 1]) if rank > 0 else None
            kp.d_up_dst[k] = slot(rank + 1, k, 0) if rank < P - 1 else None
            kp.hseq[k] = 0
            self.vec.append(torch.as_tensor(_RawCuda(slot(rank, k, 0), part.n_loc), device="cuda"))
        self.closed = False
        self.broken = False

    @staticmethod
    def nbytes(n_loc):
        return PeerRegion.CTRL_BYTES + PeerRegion.NVEC * _align256(n_loc * 8)

    @classmethod
    def create_ipc(cls, part, dist):
        import ctypes
        lib = _lib.load()
        P, rank = part.world, part.rank
        base = ctypes.c_void_p()
        handle = ctypes.create_string_buffer(64)
        _lib.check(lib.kb_peer_alloc(cls.nbytes(part.n_loc), ctypes.byref(base), handle))
        info = [None] * P
        dist.all_gather_object(info, (bytes(handle.raw), part.n_loc, part.own_off, part.n_own))
        bases = [None] * P
        bases[rank] = base.value
        for r in range(P):
            if r != rank:
                ptr = ctypes.c_void_p()
                _lib.check(lib.kb_peer_open(info[r][0], ctypes.byref(ptr)))
                bases[r] = ptr.value
        reg = cls(part, bases, [i[1] for i in info], [i[2] for i in info], [i[3] for i in info], ("ipc", dist))
        dist.barrier()
        return reg

    @classmethod
    def create_local(cls, parts):
        bufs = [torch.zeros(cls.nbytes(p.n_loc), dtype=torch.uint8, device="cuda") for p in parts]
        bases = [b.data_ptr() for b in bufs]
        return [cls(p, bases, [q.n_loc for q in parts], [q.own_off for q in parts], [q.n_own for q in parts], ("local", bufs))
                for p in parts]

    def error(self):
        e = _lib.I64(0)
        _lib.check(self.lib.kb_peer_error(self.part.ctx, e, D.stream_ptr()))
        return int(e.value)

    def close(self):
        """Unmaps the peers' regions and frees the own one (IPC form; collective: every rank must call it)."""
        if self.closed:
            return
        self.closed = True
        self.vec = []
        kind, h = self._owner
        if kind == "ipc":
            torch.cuda.synchronize()
            try:
                h.barrier()                                  # nobody unmaps memory a peer kernel may still write
            except Exception:
                pass
            rank = self.part.ctx.rank
            for r, bptr in enumerate(self.base):
                if r != rank and bptr:
                    self.lib.kb_peer_close(bptr)
            try:
                h.barrier()
            except Exception:
                pass
            self.lib.kb_peer_free(self.base[rank])
        self._owner = (kind, None)

    def __del__(self):
        try:
            if not self.closed and self._owner[0] == "local":
                self.closed = True
        except Exception:
            pass


def _peer_regions(systems, comm):
    """The systems' PeerRegions (created on first use, kept on the LocalSystem: IPC mappings are expensive)."""
    if all(getattr(s, "peer_region", None) is not None and not s.peer_region.closed for s in systems):
        return [s.peer_region for s in systems]
    if isinstance(comm, DistComm):
        regs = [PeerRegion.create_ipc(systems[0].part, comm.dist)]
    else:
        regs = PeerRegion.create_local([s.part for s in systems])
    for s, r in zip(systems, regs):
        s.peer_region = r
    return regs


def close_peer_regions(systems):
    for s in systems:
        r = getattr(s, "peer_region", None)
        if r is not None:
            r.close()
            s.peer_region = None


def _peer_capable(systems, comm):
    """The native peer loops need one part per process (or all parts in one process) and 16-byte aligned row views of the
    local CSR (a multiple of 4 nodes per mesh row: the bulk-copy SpMV kernels); otherwise the host-driven loops run."""
    if not (isinstance(comm, LocalComm) or (len(systems) == 1 and isinstance(comm, DistComm))):
        return False
    return all(s.part.world == 1 or s.cols16 is not None for s in systems)


def _peer_op(sysm, vals):
    """kb_peer_op of the owned-row view of a LocalSystem with values `vals` (returned with the tensors it points into)."""
    op = _lib.kb_peer_op()
    off = sysm.part.own_off
    op.nnz_own = sysm.nnz_own
    op.d_indptr_own = sysm.indptr.data_ptr() + 4 * off
    op.d_indices = sysm.indices.data_ptr()
    op.d_cols16 = sysm.cols16.data_ptr() if sysm.cols16 is not None else None
    op.d_row_pat_own = sysm.pat_ptr(off)
    op.d_pat_tab = sysm.pat_tab.data_ptr() if sysm.pat_tab is not None else None
    op.d_vals = vals.data_ptr()
    op.d_rowblk = sysm.rowblk.data_ptr()
    return op


def _run_ranks(systems, comm, fn):
    """fn(k, system) for every local part: directly for one part per process; one host thread + one stream per part when
    several ranks share a process / device (their kernels wait for each other, so they must be in flight together)."""
    if len(systems) == 1:
        return [fn(0, systems[0])]
    import threading
    out, errs = [None] * len(systems), [None] * len(systems)
    streams = [torch.cuda.Stream() for _ in systems]
    _lib.check(systems[0].lib.kb_peer_preload())          # no lazy kernel load while a peer's kernel is waiting
    torch.cuda.synchronize()

    dev = torch.cuda.current_device()

    def work(k):
        try:
            torch.cuda.set_device(dev)
            with torch.cuda.stream(streams[k]):
                out[k] = fn(k, systems[k])
                streams[k].synchronize()
        except BaseException as e:                             # noqa: B902 -- re-raised below
            errs[k] = e

    th = [threading.Thread(target=work, args=(k,)) for k in range(len(systems))]
    for t in th:
        t.start()
    for t in th:
        t.join()
    torch.cuda.synchronize()
    for e in errs:
        if e is not None:
            raise e
    return out


def _peer_failed_anywhere(comm, regions, rcs):
    """True when ANY rank saw a peer wait time out (all ranks leave together: ADVICE r01)."""
    bad = 1.0 if any(rc == _lib.KB_EPEER for rc in rcs) else 0.0
    if isinstance(comm, DistComm):
        t = torch.tensor([bad], dtype=torch.float64, device="cuda")
        comm.dist.all_reduce(t, op=comm.dist.ReduceOp.MAX, group=comm.group)
        bad = float(t.item())
    if bad:
        for r in regions:
            r.broken = True
    return bool(bad)


def peer_krylov(systems, comm, Ahat, b, x, scales, sym, rtol, maxiter, check_every):
    """Native partitioned CG / BiCGSTAB (kb_peer_cg_solve / kb_peer_bicgstab_solve) on the scaled operators.  b, x: full
    local vectors per part (x: initial guess in / solution out, owned range).  Returns the info dict of rank 0 (every rank
    computes the same one)."""
    regions = _peer_regions(systems, comm)
    if any(r.broken for r in regions):
        raise _lib.KiltroB200Error("the peer regions of this system are unusable after a timed-out exchange")
    # workspaces are allocated up front (an allocation may synchronise the device -- not while a peer's kernel waits)
    wss = [D.empty(s.lib.kb_peer_krylov_workspace_bytes(s.part.n_own, s.nnz_own), torch.uint8) for s in systems]
    ops = [_peer_op(s, Ahat[k]) for k, s in enumerate(systems)]

    def one(k, s):
        lib, op, ws = s.lib, ops[k], wss[k]
        nbytes = ws.numel()
        info = _lib.kb_solve_info()
        if sym:
            rc = lib.kb_peer_cg_solve(regions[k].part, op, s.own_ptr(b[k]), s.own_ptr(x[k]), s.own_ptr(scales[k]), rtol,
                                      int(maxiter), int(check_every), D.ptr(ws), nbytes, info, D.stream_ptr())
        else:
            rc = lib.kb_peer_bicgstab_solve(regions[k].part, op, s.own_ptr(b[k]), s.own_ptr(x[k]), rtol, int(maxiter),
                                            int(check_every), D.ptr(ws), nbytes, info, D.stream_ptr())
        return rc, info

    res = _run_ranks(systems, comm, one)
    if _peer_failed_anywhere(comm, regions, [r[0] for r in res]):
        raise _lib.KiltroB200Error("peer-memory exchange timed out (a rank fell out of the iteration): return codes %r, "
                                   "error words %r, events %r" % ([r[0] for r in res], [g.error() for g in regions],
                                                                  [(int(g.part.ev), list(g.part.hseq)) for g in regions]))
    for rc, _ in res:
        _lib.check(rc)
    info = res[0][1]
    rel = float(info.residual) / float(info.rhs_norm) if info.rhs_norm > 0 else 0.0
    return {"iterations": int(info.iterations), "status": int(info.status), "relative_residual": rel,
            "true_relative_residual": rel, "restarts": int(info.restarts),
            "attainable_accuracy": bool(info.attainable_accuracy), "kernel_launches": int(info.kernel_launches)}


def _run_transient_peer(systems, comm, Kff, Ahat, scales, sym, dt, nsteps, rtol, maxiter):
    regions = _peer_regions(systems, comm)
    if any(r.broken for r in regions):
        raise _lib.KiltroB200Error("the peer regions of this system are unusable after a timed-out exchange")
    for s, reg in zip(systems, regions):
        reg.vec[2].copy_(s.T0)
    torch.cuda.synchronize()
    comm.barrier()
    wss = [D.empty(s.lib.kb_peer_transient_workspace_bytes(s.part.n_own, s.nnz_own), torch.uint8) for s in systems]
    ops = [_peer_op(s, Ahat[k]) for k, s in enumerate(systems)]

    def one(k, s):
        lib, off, op, ws = s.lib, s.part.own_off, ops[k], wss[k]
        nbytes = ws.numel()
        info = _lib.kb_solve_info()
        rc = lib.kb_peer_transient_run(regions[k].part, op, D.ptr(Kff[k]), 1 if sym else 0, s.own_ptr(scales[k]),
                                       s.own_ptr(s.Fm), s.bc.renum.data_ptr() + 4 * off,
                                       D.ptr(s.bc.applied_dirichlet_device), float(dt), int(nsteps), float(rtol),
                                       int(maxiter), D.ptr(ws), nbytes, info, D.stream_ptr())
        return rc, info

    res = _run_ranks(systems, comm, one)
    if _peer_failed_anywhere(comm, regions, [r[0] for r in res]):
        raise _lib.KiltroB200Error("peer-memory exchange timed out (a rank fell out of the time loop): return codes %r, error "
                                   "words %r, events %r" % ([r[0] for r in res], [g.error() for g in regions],
                                                            [(int(g.part.ev), list(g.part.hseq)) for g in regions]))
    for rc, _ in res:
        _lib.check(rc)
    info = res[0][1]
    if info.status != 0:
        from warnings import warn
        warn("partitioned transient loop stopped at step %d of %d: inner solve failed (status %d)" % (
            info.failed_step, nsteps, info.status))
    out = [reg.vec[2][s.part.own()].clone() for s, reg in zip(systems, regions)]
    return out, {"iterations": int(info.iterations), "worst_relative_residual": float(info.residual),
                 "method": "cg" if sym else "bicgstab", "status": int(info.status), "failed_step": int(info.failed_step),
                 "steps_done": int(info.steps_done), "exchange": "peer", "kernel_launches": int(info.kernel_launches)}


# ------------------------------------------------------------------------------------------------------ drivers
def _true_residual(systems, comm, Ahat, b, x):
    """(||b - A x||^2, ||b||^2) of the partitioned system: one halo exchange, SpMV and reduction."""
    comm.halo_exchange([s.part for s in systems], x)
    dres = [D.zeros(2) for _ in systems]
    for k, s in enumerate(systems):
        q = D.zeros(s.part.n_loc)
        s.spmv(Ahat[k], x[k], q)
        _lib.check(s.lib.kb_residual(s.part.n_own, s.own_ptr(b[k]), s.own_ptr(q), s.own_ptr(q), D.ptr(dres[k]),
                                     D.ptr(s.scratch), D.stream_ptr()))
    comm.allreduce(dres)
    return float(dres[0][0]), float(dres[0][1])


def residual_check(sysm, comm, sol_own):
    """||b - A' x|| / ||b|| of a returned solution on the UNSCALED operator A' = [Kff 0; 0 I], computed with the host-driven
    exchange of `comm` (halo rows through the communicator, plain SpMV, all-reduce): an independent check of solutions
    produced by the peer-memory kernels."""
    part = sysm.part
    A = _masked(sysm, 1.0, 0.0, 1)
    x = D.zeros(part.n_loc)
    x[part.own()] = sol_own
    comm.halo_exchange([part], [x])
    q = D.zeros(part.n_loc)
    sysm.spmv(A, x, q)
    d = D.zeros(2)
    _lib.check(sysm.lib.kb_residual(part.n_own, sysm.own_ptr(sysm.b), sysm.own_ptr(q), sysm.own_ptr(q), D.ptr(d),
                                    D.ptr(sysm.scratch), D.stream_ptr()))
    comm.allreduce([d])
    rr, bb = float(d[0]), float(d[1])
    return (rr / bb) ** 0.5 if bb > 0 else 0.0


def solve_steady(systems, comm, rtol=1e-10, maxiter=20000, check_every=50, method="auto", exchange="nccl",
                 max_restarts=16, cycle_iterations=None):
    """Partitioned steady solve.  Returns (list of owned-solution tensors (n_own,), info).
    exchange:
      "peer"  the native loop (csrc/krylov_peer.cuh, CG and BiCGSTAB): halo rows and dot products travel through NVLink
              peer memory inside the kernels, convergence / breakdown are decided on device, the host looks at its own
              device's state every check_every iterations -- no collective call inside the solve.  One part per process
              (DistComm) or several parts in one process on one device (LocalComm: one thread + stream per part);
      "nccl"  host-driven loops of this module: halo rows and dot products through the communicator (torch.distributed).
    Like the single-GPU driver (csrc/krylov.cu), the iterate is judged by its TRUE residual and the Krylov loop is
    restarted from it when the recurrence broke down / diverged or "converged" on an iterate whose true residual does not
    confirm it.  cycle_iterations caps one cycle of the host-driven loops (tests use it to exercise the restart path)."""
    mats = [_masked(s, 1.0, 0.0, 1) for s in systems]
    sym = all(s.symmetric for s in systems) if method == "auto" else (method == "cg")
    mode = 1 if sym else 0
    scales, Ahat = _jacobi_scale(systems, comm, mats, mode)
    del mats
    b, x = [], []
    for s, sc in zip(systems, scales):
        bb = s.b.clone()
        if mode == 1:
            _lib.check(s.lib.kb_vec_mul(s.part.n_loc, D.ptr(s.b), D.ptr(sc), D.ptr(bb), D.stream_ptr()))
        b.append(bb); x.append(D.zeros(s.part.n_loc))
    peer = exchange == "peer" and _peer_capable(systems, comm)
    if peer:
        info = peer_krylov(systems, comm, Ahat, b, x, scales, sym, rtol, maxiter, check_every)
        info["exchange"] = "peer"
        info["method"] = "cg" if sym else "bicgstab"
    else:
        info = _solve_steady_host_loops(systems, comm, Ahat, b, x, sym, rtol, maxiter, check_every, max_restarts,
                                        cycle_iterations)
    sols = []
    for s, sc, xh in zip(systems, scales, x):
        out = D.empty(s.part.n_loc)
        _lib.check(s.lib.kb_vec_mul(s.part.n_loc, D.ptr(xh), D.ptr(sc), D.ptr(out), D.stream_ptr()))   # x = D^-1 xhat (or D^-1/2)
        sols.append(out[s.part.own()].clone())
    return sols, info


def _solve_steady_host_loops(systems, comm, Ahat, b, x, sym, rtol, maxiter, check_every, max_restarts, cycle_iterations):
    info = {"iterations": 0, "status": _lib.KB_ENOTCONV, "restarts": 0, "attainable_accuracy": False}
    tol2_factor = rtol * rtol
    prev_true = float("inf")
    rr_true, bb_true = float("nan"), 0.0
    for cycle in range(max_restarts + 1):
        left = maxiter - info["iterations"]
        if left <= 0:
            break
        budget = left if cycle_iterations is None else min(left, cycle_iterations)
        c = (dist_cg if sym else dist_bicgstab)(systems, comm, Ahat, b, x, rtol, budget, check_every)
        info["iterations"] += c["iterations"]
        info["status"] = c["status"]
        info["restarts"] = cycle
        rr_true, bb_true = _true_residual(systems, comm, Ahat, b, x)
        if not np.isfinite(rr_true):
            info["status"] = _lib.KB_EBREAKDOWN           # nothing sane to restart from
            break
        if rr_true <= tol2_factor * bb_true:
            info["status"] = 0
            break
        # attainable accuracy (same rule as csrc/krylov.cu): a whole cycle did not halve the TRUE residual norm while it
        # sits within 10x of the tolerance, or the recurrence converged while the true residual stalls
        stalled = not (rr_true < 0.25 * prev_true)
        if (cycle > 0 and stalled and rr_true <= 100.0 * tol2_factor * bb_true) or \
                (c["status"] == 0 and stalled and rr_true <= 1.0e6 * tol2_factor * bb_true):
            info["attainable_accuracy"] = True
            break
        if c["status"] == _lib.KB_ENOTCONV and cycle_iterations is None:
            break                                          # maxiter exhausted
        prev_true = rr_true
    info["exchange"] = "nccl" if isinstance(comm, DistComm) else "local"
    info["method"] = "cg" if sym else "bicgstab"
    # the verdict is taken on the TRUE residual only (for CG: of the symmetrically scaled system the loop iterates on)
    info["relative_residual"] = info["true_relative_residual"] = (rr_true / bb_true) ** 0.5 if bb_true > 0 else 0.0
    if bb_true == 0.0 and rr_true == 0.0:
        info["status"] = 0
    elif not np.isfinite(rr_true):
        info["status"] = _lib.KB_EBREAKDOWN
    elif rr_true <= tol2_factor * bb_true:
        info["status"] = 0
    elif info["attainable_accuracy"]:
        info["status"] = 0 if rr_true <= 100.0 * tol2_factor * bb_true else _lib.KB_ESTAGNATED
    elif info["status"] == 0:
        info["status"] = _lib.KB_ENOTCONV
    return info


def run_transient(systems, comm, dt, nsteps, theta=0.0, rtol=1e-10, maxiter=2000, check_every=4, exchange="nccl"):
    """Partitioned theta-method loop (same scheme as kb_transient_run).  Returns (list of owned T^nsteps, info).
    exchange="peer": the whole time loop runs inside kb_peer_transient_run (csrc/krylov_peer.cuh) -- halo rows of T and of
    the inner Krylov vectors and every dot product through peer memory, one or two host round trips per step; "nccl":
    the host-driven loop below."""
    parts = [s.part for s in systems]
    Kff = [_masked(s, 1.0, 0.0, 0) for s in systems]
    A = [_masked(s, theta * dt, 1.0, 0) for s in systems]                 # M + theta dt Kff  (rows of fixed DOFs: M only)
    sym = theta == 0.0 or all(s.symmetric for s in systems)
    mode = 1 if sym else 0
    scales, Ahat = _jacobi_scale(systems, comm, A, mode)
    del A
    if exchange == "peer" and _peer_capable(systems, comm):
        return _run_transient_peer(systems, comm, Kff, Ahat, scales, sym, dt, nsteps, rtol, maxiter)
    T = [s.T0.clone() for s in systems]
    yh = [D.zeros(s.part.n_loc) for s in systems]
    kt = [D.zeros(s.part.n_loc) for s in systems]
    rhs = [D.zeros(s.part.n_loc) for s in systems]
    total = 0
    worst = 0.0
    status, failed_step, steps_done = 0, 0, 0
    for step in range(nsteps):
        comm.halo_exchange(parts, T)
        for k, s in enumerate(systems):
            s.spmv(Kff[k], T[k], kt[k])
            torch.add(kt[k], s.Fm, out=rhs[k])
            if mode == 1:
                _lib.check(s.lib.kb_vec_mul(s.part.n_loc, D.ptr(rhs[k]), D.ptr(scales[k]), D.ptr(rhs[k]), D.stream_ptr()))
        info = (dist_cg if sym else dist_bicgstab)(systems, comm, Ahat, rhs, yh, rtol, maxiter, check_every,
                                                       checkpoint=False)   # short inner solves: no restart
        total += info["iterations"]
        rel = info["relative_residual"]
        if info["status"] != 0 or not np.isfinite(rel):
            # a failed inner solve stops the time loop: T keeps the last good state and the caller learns which step failed
            status = info["status"] if info["status"] != 0 else _lib.KB_EBREAKDOWN
            failed_step = step + 1
            worst = rel if not np.isfinite(rel) else max(worst, rel)
            break
        worst = max(worst, rel)
        steps_done = step + 1
        for k, s in enumerate(systems):
            _lib.check(s.lib.kb_vec_mul(s.part.n_loc, D.ptr(yh[k]), D.ptr(scales[k]), D.ptr(kt[k]), D.stream_ptr()))   # y = D^-1 yhat
            _lib.check(s.lib.kb_transient_update(s.part.n_loc, D.ptr(s.bc.renum), D.ptr(s.bc.applied_dirichlet_device),
                                                 D.ptr(kt[k]), float(dt), D.ptr(T[k]), D.stream_ptr()))
    if status != 0:
        from warnings import warn
        warn("partitioned transient loop stopped at step %d of %d: inner solve failed (status %d)" % (failed_step, nsteps, status))
    return [Tk[s.part.own()].clone() for Tk, s in zip(T, systems)], {"iterations": total, "worst_relative_residual": worst,
                                                                     "method": "cg" if sym else "bicgstab", "status": status,
                                                                     "failed_step": failed_step, "steps_done": steps_done}
