"""Host side of the fused assembly (csrc/assemble_fused.cu): builds/caches the tile plan and launches the kernel."""
import torch

from .. import _lib, device as D


class TilePlan(object):
    def __init__(self):
        self.ntiles = 0
        self.tile_rows = self.tile_row_ptr = self.tile_conn = self.tile_elem_ptr = None
        self.pair_rec = self.pair_ptr = None
        self.info = {}
        self.fits = False
        # template form (csrc/tiletmpl.cu): usable when the tiles of the mesh repeat
        self.templated = False
        self.tile_class = self.trun_ptr = self.trun_row = self.templates = self.scratch = None


def grid_shape(sp, mesh):
    """(nxn, nyn) when the mesh is numbered like Mesh.Rectangle (checked on device once and cached with the pattern:
    mesh.structured_shape is only the hint that says which shape to check), else None -- csrc/assemble_grid.cu."""
    if getattr(sp, "grid", None) is None:
        sp.grid = False
        shape = mesh.structured_shape if (mesh.structured_shape is not None and not mesh.host_is_authoritative()) else None
        el = mesh.device_elements()
        quads = sp.nvar == 1 and el.dim() == 2 and int(el.shape[1]) == 4 and int(el.shape[0]) > 0
        if shape is None and quads:
            # user-provided arrays: the first element of a Mesh.Rectangle numbering is [0, 1, nxn + 1, nxn]
            first = [int(v) for v in el[0].tolist()]
            nxn, nn = first[3], int(mesh.nnodes)
            if first[0] == 0 and first[1] == 1 and nxn >= 2 and first[2] == nxn + 1 and nn % nxn == 0 and nn // nxn >= 2:
                shape = (nxn, nn // nxn)
        if shape is not None and quads and min(shape) >= 2:
            flag = D.empty(1, torch.int32)
            _lib.check(_lib.load().kb_grid_verify(D.ptr(el), int(el.shape[0]), int(shape[0]), int(shape[1]), D.ptr(flag),
                                                  D.stream_ptr()))
            if int(flag.item()) == 0:
                sp.grid = (int(shape[0]), int(shape[1]))
                sp.grid_scratch = D.empty(32, torch.uint8)
    return sp.grid or None


def line_numbered(sp, mesh):
    """True when a line-2 mesh is numbered like Mesh.Line (element i = [i, i+1]); checked on device once, cached with the
    pattern -- csrc/assemble_line.cu."""
    if getattr(sp, "line", None) is None:
        sp.line = False
        el = mesh.device_elements()
        if sp.nvar == 1 and el.dim() == 2 and int(el.shape[1]) == 2 and int(el.shape[0]) > 0 and int(mesh.nnodes) >= 2:
            flag = D.empty(1, torch.int32)
            _lib.check(_lib.load().kb_line_verify(D.ptr(el), int(el.shape[0]), int(mesh.nnodes), D.ptr(flag), D.stream_ptr()))
            sp.line = int(flag.item()) == 0
    return sp.line


def assemble_line(fem_solver, sp, formulation, mesh, material, K, M, Flux):
    fem_solver.assembly_used = "fused_line"
    pts, vel = mesh.device_points(), formulation.device_velocity()
    _lib.check(_lib.load().kb_assemble_line(D.ptr(pts), D.ptr(vel), material.as_struct(), int(formulation.mode),
                                            int(bool(formulation.petrov)), int(mesh.nnodes), D.ptr(sp.indptr), D.ptr(K),
                                            D.ptr(M), D.ptr(Flux), D.stream_ptr()))


def assemble_grid(fem_solver, sp, formulation, mesh, material, K, M, Flux):
    nxn, nyn = sp.grid
    fem_solver.assembly_used = "fused_grid"
    pts, vel = mesh.device_points(), formulation.device_velocity()    # (kept alive: uploads of host arrays are temporaries)
    _lib.check(_lib.load().kb_assemble_grid(D.ptr(pts), D.ptr(vel),
                                            material.as_struct(), int(formulation.mode), int(bool(formulation.petrov)),
                                            nxn, nyn, D.ptr(sp.indptr), D.ptr(K), D.ptr(M), D.ptr(Flux),
                                            D.ptr(sp.grid_scratch), D.stream_ptr()))


def build_tile_plan(sp, mesh):
    """Symbolic phase, part 2 (once per mesh): tiles of rows + their incident elements + per-(row, element) records."""
    lib = _lib.load()
    el = mesh.device_elements()
    nelem, npe = int(el.shape[0]), int(el.shape[1])
    nnode = int(mesh.nnodes)
    plan = TilePlan()
    if nelem == 0 or sp.nvar != 1:
        return plan
    shape = mesh.structured_shape if (mesh.structured_shape is not None and not mesh.host_is_authoritative()) else None
    nxn, nyn = (shape if shape is not None else (0, 0))
    sizes = (_lib.I64 * 4)()
    _lib.check(lib.kb_tile_plan_sizes(nnode, nelem, npe, nxn, nyn, sizes))
    ntiles = int(sizes[0])
    tile_rows = D.empty(nnode, torch.int32)
    tile_row_ptr = D.empty(ntiles + 1, torch.int32)
    tile_elems = D.empty(int(sizes[1]), torch.int32)
    tile_conn = D.empty(int(sizes[1]) * npe, torch.int32)
    tile_elem_ptr = D.empty(ntiles + 1, torch.int32)
    pair_rec = D.empty(int(sizes[2]), torch.int32)          # uint32 payload, int32 storage
    pair_ptr = D.empty(nnode + 1, torch.int32)
    info = (_lib.I64 * 8)()
    _lib.check(lib.kb_tile_plan_build(D.ptr(el), nelem, npe, nnode, D.ptr(sp.indptr), D.ptr(sp.indices),
                                      D.ptr(sp.seg_ptr), D.ptr(sp.gather_src), D.ptr(sp.diag_pos), nxn, nyn,
                                      D.ptr(tile_rows), D.ptr(tile_row_ptr), D.ptr(tile_elems), D.ptr(tile_conn), D.ptr(tile_elem_ptr),
                                      D.ptr(pair_rec), D.ptr(pair_ptr), info, D.stream_ptr()))
    plan.ntiles = ntiles
    plan.tile_rows, plan.tile_row_ptr, plan.tile_elem_ptr, plan.pair_ptr = tile_rows, tile_row_ptr, tile_elem_ptr, pair_ptr
    plan.tile_conn = tile_conn[:int(info[1]) * npe].clone()
    plan.pair_rec = pair_rec[:int(info[2])].clone()
    plan.info = {"ntiles": int(info[0]), "tile_elements": int(info[1]), "pairs": int(info[2]),
                 "max_elements_per_tile": int(info[3]), "max_rows_per_tile": int(info[4]),
                 "max_entries_per_tile": int(info[5]), "structured": bool(info[7]),
                 "rim_factor": float(info[1]) / max(nelem, 1)}
    plan.fits = bool(info[6])
    plan.scratch = D.empty(plan.ntiles + 32, torch.uint8)
    if plan.fits and npe == 4:
        _build_templates(lib, plan, sp, nnode, npe)
    return plan


def _build_templates(lib, plan, sp, nnode, npe):
    """Symbolic phase, part 3: classify the tiles by local structure; one template per class (tiletmpl.cu)."""
    sizes = (_lib.I64 * 3)()
    _lib.check(lib.kb_tile_tmpl_sizes(sizes))
    tile_class = D.empty(plan.ntiles, torch.uint8)
    trun_ptr = D.empty(plan.ntiles + 1, torch.int32)
    trun_row = D.empty(nnode, torch.int32)
    templates = D.empty(int(sizes[0]) * int(sizes[1]), torch.uint8)
    info = (_lib.I64 * 3)()
    _lib.check(lib.kb_tile_tmpl_build(npe, nnode, plan.ntiles, D.ptr(sp.indptr), D.ptr(plan.tile_rows),
                                      D.ptr(plan.tile_row_ptr), D.ptr(plan.tile_elem_ptr), D.ptr(plan.pair_rec),
                                      D.ptr(plan.pair_ptr), D.ptr(tile_class), D.ptr(trun_ptr), D.ptr(trun_row),
                                      D.ptr(templates), info, D.stream_ptr()))
    plan.info["templates"] = int(info[1])
    plan.info["template_runs"] = int(info[2])
    plan.templated = bool(info[0])
    if plan.templated:
        plan.tile_class, plan.trun_ptr = tile_class, trun_ptr
        plan.trun_row = trun_row[:max(int(info[2]), 1)].clone()
        plan.templates = templates[:int(sizes[0]) * int(info[1])].clone()


def assemble_fused(fem_solver, sp, formulation, mesh, material, K, M, Flux):
    lib = _lib.load()
    plan = sp.tile_plan
    pts = mesh.device_points()
    el = mesh.device_elements()
    vel = formulation.device_velocity()
    mat = material.as_struct()
    kind = getattr(fem_solver, "assembly", "auto")
    if plan.templated and formulation.ndim == 2 and kind != "fused_generic":
        fem_solver.assembly_used = "fused_template"
        _lib.check(lib.kb_assemble_fused_tmpl(D.ptr(pts), int(mesh.nnodes), D.ptr(vel), mat, int(formulation.mode),
                                              int(bool(formulation.petrov)), D.ptr(sp.indptr), plan.ntiles,
                                              D.ptr(plan.tile_conn), D.ptr(plan.tile_elem_ptr), D.ptr(plan.tile_class),
                                              D.ptr(plan.trun_ptr), D.ptr(plan.trun_row), D.ptr(plan.templates),
                                              D.ptr(K), D.ptr(M), D.ptr(Flux), D.ptr(plan.scratch), D.stream_ptr()))
        return
    _lib.check(lib.kb_assemble_fused(formulation.ndim, D.ptr(pts), D.ptr(el), int(el.shape[0]), int(mesh.nnodes),
                                     D.ptr(vel), mat, int(formulation.mode), int(bool(formulation.petrov)),
                                     D.ptr(sp.indptr), D.ptr(sp.indices), plan.ntiles, D.ptr(plan.tile_rows),
                                     D.ptr(plan.tile_row_ptr), D.ptr(plan.tile_conn), D.ptr(plan.tile_elem_ptr),
                                     D.ptr(plan.pair_rec), D.ptr(plan.pair_ptr), D.ptr(K), D.ptr(M), D.ptr(Flux),
                                     D.ptr(plan.scratch), D.stream_ptr()))
