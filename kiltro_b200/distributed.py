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
from .MeshGeneration import Mesh
from .BoundaryCondition import BoundaryCondition
from .FEMSolver import FEMSolver
from .FiniteElements.Assembly import Assemble


# Hide the halo exchange behind the interior rows of the SpMV (3 launches instead of 1).  Measured on 2 B200s at 16 M DOF
# per GPU: no gain (the exchange is short compared with the launch overhead it adds), so it is off by default.
OVERLAP_HALO = False
USE_ROW_PATTERNS = True       # row-pattern form of the SpMV (csrc/spmv_pat.cuh) when the local matrix allows it


class StripPartition(object):
    """Mesh rows [j0, j1) of an (nx_nodes x ny_nodes) node grid owned by `rank`, plus one halo row per interior side."""

    def __init__(self, nx_nodes, ny_nodes, world, rank):
        if world < 1 or not (0 <= rank < world):
            raise ValueError("bad rank/world")
        if ny_nodes < 2 * world:
            raise ValueError("need at least two mesh rows per rank")
        self.nxn, self.nyn, self.world, self.rank = int(nx_nodes), int(ny_nodes), int(world), int(rank)
        base, rem = divmod(self.nyn, self.world)
        self.j0 = rank * base + min(rank, rem)
        self.j1 = self.j0 + base + (1 if rank < rem else 0)
        self.hb = 1 if rank > 0 else 0                  # halo row below (owned by rank-1)
        self.ha = 1 if rank < world - 1 else 0          # halo row above (owned by rank+1)
        self.j_first = self.j0 - self.hb
        self.nrows_loc = self.j1 - self.j0 + self.hb + self.ha
        self.n_own = (self.j1 - self.j0) * self.nxn
        self.n_loc = self.nrows_loc * self.nxn
        self.own_off = self.hb * self.nxn
        self.global_off = self.j0 * self.nxn            # global node id of the first owned node

    # slices of a full local vector
    def own(self):
        return slice(self.own_off, self.own_off + self.n_own)

    def first_own_row(self):
        return slice(self.own_off, self.own_off + self.nxn)

    def last_own_row(self):
        return slice(self.own_off + self.n_own - self.nxn, self.own_off + self.n_own)

    def halo_below(self):
        return slice(0, self.nxn) if self.hb else None

    def halo_above(self):
        return slice(self.own_off + self.n_own, self.own_off + self.n_own + self.nxn) if self.ha else None


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


def build_local_system(part, rectangle, material, formulation_cls, velocity_fn, dirichlet_fn, neumann_fn=None,
                       transient=False, initial_fn=None):
    """Generates the rank's strip (+ halo rows) of Mesh.Rectangle(*rectangle) on device, assembles it with the ordinary
    single-GPU machinery and prepares the partitioned operators.
    velocity_fn / dirichlet_fn / neumann_fn / initial_fn: callables (points (n_loc, 2) CUDA tensor) -> CUDA tensor
    ((n_loc, 2) velocity or None; (n_loc, 1) flags with NaN = unconstrained; (n_loc, 1) initial field)."""
    lib = _lib.load()
    (x0, y0), (x1, y1), nx, ny = rectangle
    if nx + 1 != part.nxn or ny + 1 != part.nyn:
        raise ValueError("partition does not match the rectangle")
    pts = D.empty((part.n_loc, 2)); el = D.empty(((part.nrows_loc - 1) * nx, 4), torch.int32)
    _lib.check(lib.kb_mesh_rectangle_strip(float(x0), float(y0), float(x1), float(y1), nx, ny, part.j_first,
                                           part.nrows_loc, D.ptr(pts), D.ptr(el), D.stream_ptr()))
    mesh = Mesh(element_type="quad")
    mesh._set_device(pts, el, 2)
    mesh.structured_shape = (part.nxn, part.nrows_loc)
    vel = velocity_fn(pts) if velocity_fn is not None else None
    form = formulation_cls(mesh, velocity=vel) if vel is not None else formulation_cls(mesh)
    fs = FEMSolver(analysis_type="transient" if transient else "steady", verbose=False)
    bc = BoundaryCondition()
    flags = dirichlet_fn(pts)
    bc.SetDirichletCriteria(lambda: flags)
    if neumann_fn is not None:
        nfl = neumann_fn(pts)
        bc.SetNeumannCriteria(lambda: nfl)
    sysm = LocalSystem(part)
    sysm.mesh, sysm.form, sysm.bc, sysm.fs, sysm.material = mesh, form, bc, fs, material
    sysm.transient, sysm.initial_fn = transient, initial_fn
    sysm.velocity, sysm.flags = vel, flags
    sysm.symmetric = bool(form.is_symmetric)
    sysm.rowblk = None
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
    K, Flux, M = Assemble(fs, form.function_spaces, form, mesh, material, bc)
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
        if sysm.initial_fn is not None:
            T0 = sysm.initial_fn(mesh.device_points()).reshape(-1).to(torch.float64).clone()
            T0 += bc.UpdateFixDoFs(bc.applied_dirichlet_device, n, 1).reshape(-1)
            sysm.T0 = T0
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
    """IPC-mapped memory of one rank (control block + the two SpMV input vectors p and r) and the mappings of every
    other rank's region: what the peer-memory kernels (csrc/peer.cuh) write into.  Needs one process per GPU with NVLink
    peer access (a B200 HGX node); collectives are only used to exchange the 64-byte IPC handles."""
    CTRL_BYTES = 1024

    def __init__(self, part, dist):
        import ctypes
        lib = _lib.load()
        self.part, self.lib = part, lib
        P, rank = part.world, part.rank
        if P > 8:
            raise ValueError("the peer-memory path supports up to 8 ranks")
        self.stride = _align256(part.n_loc * 8)
        nbytes = self.CTRL_BYTES + 2 * self.stride
        base = ctypes.c_void_p()
        handle = ctypes.create_string_buffer(64)
        _lib.check(lib.kb_peer_alloc(nbytes, ctypes.byref(base), handle))
        self.base = [None] * P
        self.base[rank] = base.value
        info = [None] * P
        dist.all_gather_object(info, (bytes(handle.raw), part.n_loc, part.own_off, part.n_own))
        for r in range(P):
            if r != rank:
                ptr = ctypes.c_void_p()
                _lib.check(lib.kb_peer_open(info[r][0], ctypes.byref(ptr)))
                self.base[r] = ptr.value
        self.ctx = _lib.kb_peer_ctx()
        self.ctx.nranks, self.ctx.rank = P, rank
        for r in range(P):
            self.ctx.ctrl[r] = self.base[r]
        self.vec = [torch.as_tensor(_RawCuda(self.base[rank] + self.CTRL_BYTES + k * self.stride, part.n_loc), device="cuda")
                    for k in range(2)]

        def slot(r, k, offset_elems):
            stride_r = _align256(info[r][1] * 8)
            return self.base[r] + self.CTRL_BYTES + k * stride_r + 8 * offset_elems

        # halo_above slot of rank-1 / halo_below slot of rank+1, for vector k (0 = p, 1 = r)
        self.dn_dst = [slot(rank - 1, k, info[rank - 1][2] + info[rank - 1][3]) if rank > 0 else None for k in range(2)]
        self.up_dst = [slot(rank + 1, k, 0) if rank < P - 1 else None for k in range(2)]
        self.ev = 0          # dot-product events and halo sequence numbers run on across solves (flags are monotonic)
        self.hseq = [0, 0]
        dist.barrier()

    def error(self):
        e = _lib.I64(0)
        _lib.check(self.lib.kb_peer_error(self.ctx, e, D.stream_ptr()))
        return int(e.value)


def dist_bicgstab_peer(sysm, comm, region, Ahat, b, x, rtol=1e-10, maxiter=20000, check_every=50):
    """The partitioned BiCGSTAB iteration with no NCCL call inside: partial dot products and halo rows travel through
    NVLink peer memory, written by the kernels themselves (kb_peer_*).  One part per process.  NCCL is used for the
    setup residual and for the ||r|| all-reduce when convergence is looked at (every check_every iterations)."""
    part, lib, st = sysm.part, sysm.lib, D.stream_ptr
    n, nxn, off = part.n_own, part.nxn, part.own_off
    z = lambda: D.zeros(part.n_loc)
    p, r = region.vec[0], region.vec[1]
    p.zero_(); r.zero_()
    rhat, v, t = z(), z(), z()
    dr = D.zeros(2)
    rho = [D.zeros(1), D.zeros(1)]; alpha = D.zeros(1)
    cur = 0
    own = lambda vec: vec.data_ptr() + 8 * off
    ip = sysm.indptr.data_ptr() + 4 * off
    # r = b - A x ; rhat = p = r ; first halo of p through NCCL
    comm.halo_exchange([part], [x])
    sysm.spmv(Ahat, x, v)
    _lib.check(lib.kb_residual(n, own(b), own(v), own(r), D.ptr(dr), D.ptr(sysm.scratch), st()))
    comm.allreduce([dr])
    rr0, bb = float(dr[0]), float(dr[1])
    tol2 = rtol * rtol * bb
    if not (rr0 > tol2):
        return {"iterations": 0, "relative_residual": (rr0 / bb) ** 0.5 if bb > 0 else 0.0, "status": 0}
    rhat.copy_(r); p.copy_(r); rho[0].copy_(dr[0:1])
    comm.halo_exchange([part], [p])
    torch.cuda.synchronize(); comm.barrier()            # every rank's setup is complete before peers start writing
    it, rr, status, first = 0, rr0, 0, True
    keep = _BestIterate([x], rr0)
    A = (D.ptr(sysm.indices), D.ptr(Ahat), D.ptr(sysm.rowblk), D.ptr(sysm.cols16))
    while it < maxiter:
        for _ in range(min(check_every, maxiter - it)):
            region.ev += 1                               # v = A p ; {<rhat,v>}
            _lib.check(lib.kb_peer_spmv(n, sysm.nnz_own, ip, A[0], A[3], sysm.pat_ptr(off), D.ptr(sysm.pat_tab), A[1], A[2],
                                        sysm.xptr(p, off), own(v), own(rhat), None, region.ctx,
                                        region.ev, -1 if first else 0, region.hseq[0], D.ptr(sysm.scratch), st()))
            first = False
            region.hseq[1] += 1                          # alpha ; s = r - alpha v ; boundary rows of s -> neighbours
            _lib.check(lib.kb_peer_bicgstab_s(n, nxn, region.ctx, region.ev, D.ptr(rho[cur]), own(v), own(r), D.ptr(alpha),
                                              region.dn_dst[1], region.up_dst[1], region.hseq[1], D.ptr(sysm.scratch), st()))
            region.ev += 1                               # t = A s ; {<s,t>,<t,t>,<rhat,t>,<rhat,s>}
            _lib.check(lib.kb_peer_spmv(n, sysm.nnz_own, ip, A[0], A[3], sysm.pat_ptr(off), D.ptr(sysm.pat_tab), A[1], A[2],
                                        sysm.xptr(r, off), own(t), own(r), own(rhat), region.ctx,
                                        region.ev, 1, region.hseq[1], D.ptr(sysm.scratch), st()))
            region.hseq[0] += 1                          # omega, rho, beta ; x, r, p ; boundary rows of p -> neighbours
            _lib.check(lib.kb_peer_bicgstab_update_p(n, nxn, region.ctx, region.ev, D.ptr(rho[cur]), D.ptr(rho[cur ^ 1]),
                                                     D.ptr(alpha), own(t), own(v), own(x), own(r), own(p), region.dn_dst[0],
                                                     region.up_dst[0], region.hseq[0], D.ptr(dr), D.ptr(sysm.scratch), st()))
            cur ^= 1
            it += 1
        comm.allreduce([dr])
        rr = float(dr[0])
        if region.error():
            raise _lib.KiltroB200Error("peer-memory exchange timed out (a rank fell out of the iteration)")
        if not np.isfinite(rr) or rr > 1.0e10 * rr0:           # non-finite, or diverged 1e5-fold (same bound as krylov.cu)
            status = _lib.KB_EBREAKDOWN
            rr = keep.restore([x])
            break
        keep.offer([x], rr)
        if rr <= tol2:
            break
    else:
        status = _lib.KB_ENOTCONV
    if status == 0 and rr > tol2:
        status = _lib.KB_ENOTCONV
    torch.cuda.synchronize(); comm.barrier()
    return {"iterations": it, "relative_residual": (rr / bb) ** 0.5 if bb > 0 else 0.0, "status": status}


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


def solve_steady(systems, comm, rtol=1e-10, maxiter=20000, check_every=50, method="auto", exchange="nccl",
                 max_restarts=16, cycle_iterations=None):
    """Partitioned steady solve.  Returns (list of owned-solution tensors (n_own,), info).
    exchange: "nccl" (halo rows and dot products through torch.distributed) or "peer" (BiCGSTAB only, one part per process:
    the kernels exchange them through NVLink peer memory, no collective inside the iteration).
    Like the single-GPU driver (csrc/krylov.cu), the iterate is judged by its TRUE residual and the Krylov loop is
    restarted from it (<= max_restarts times, maxiter iterations in total) when the recurrence broke down / diverged
    -- from the best checked iterate, see _BestIterate -- or "converged" on an iterate whose true residual does not
    confirm it.  cycle_iterations caps one cycle (restarted BiCGSTAB/CG; tests use it to exercise the restart path)."""
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
    peer = exchange == "peer" and not sym and len(systems) == 1 and isinstance(comm, DistComm) and comm.world > 1
    if peer and getattr(systems[0], "peer_region", None) is None:
        systems[0].peer_region = PeerRegion(systems[0].part, comm.dist)
    info = {"iterations": 0, "status": _lib.KB_ENOTCONV, "restarts": 0, "attainable_accuracy": False}
    tol2_factor = rtol * rtol
    prev_true = float("inf")
    rr_true, bb_true = float("nan"), 0.0
    for cycle in range(max_restarts + 1):
        left = maxiter - info["iterations"]
        if left <= 0:
            break
        budget = left if cycle_iterations is None else min(left, cycle_iterations)
        if peer:
            s0 = systems[0]
            c = dist_bicgstab_peer(s0, comm, s0.peer_region, Ahat[0], b[0], x[0], rtol, budget, check_every)
        else:
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
    info["exchange"] = "peer" if peer else ("nccl" if isinstance(comm, DistComm) else "local")
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
    sols = []
    for s, sc, xh in zip(systems, scales, x):
        out = D.empty(s.part.n_loc)
        _lib.check(s.lib.kb_vec_mul(s.part.n_loc, D.ptr(xh), D.ptr(sc), D.ptr(out), D.stream_ptr()))   # x = D^-1 xhat (or D^-1/2)
        sols.append(out[s.part.own()].clone())
    return sols, info


def run_transient(systems, comm, dt, nsteps, theta=0.0, rtol=1e-10, maxiter=2000, check_every=4):
    """Partitioned theta-method loop (same scheme as kb_transient_run).  Returns (list of owned T^nsteps, info)."""
    parts = [s.part for s in systems]
    Kff = [_masked(s, 1.0, 0.0, 0) for s in systems]
    A = [_masked(s, theta * dt, 1.0, 0) for s in systems]                 # M + theta dt Kff  (rows of fixed DOFs: M only)
    sym = theta == 0.0 or all(s.symmetric for s in systems)
    mode = 1 if sym else 0
    scales, Ahat = _jacobi_scale(systems, comm, A, mode)
    del A
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
