"""ctypes binding of libkiltro_b200.so (include/kiltro_b200.h).  There is NO CPU fallback: if the library is missing
or a call is made without a CUDA device the caller gets an exception, never a silently different code path."""
import ctypes
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libkiltro_b200.so")

c_i32p = ctypes.c_void_p      # all device/host buffers are passed as raw addresses
c_vp = ctypes.c_void_p
I64 = ctypes.c_int64
I32 = ctypes.c_int
F64 = ctypes.c_double


class kb_material(ctypes.Structure):
    _fields_ = [("k", F64), ("rho", F64), ("c_v", F64), ("area", F64), ("q", F64)]


class kb_peer_ctx(ctypes.Structure):
    _fields_ = [("nranks", ctypes.c_int32), ("rank", ctypes.c_int32), ("ctrl", ctypes.c_void_p * 8)]


class kb_peer_part(ctypes.Structure):
    _fields_ = [("ctx", kb_peer_ctx), ("n_loc", I64), ("n_own", I64), ("own_off", I64), ("nxn", I64),
                ("d_vec", ctypes.c_void_p * 3), ("d_dn_dst", ctypes.c_void_p * 3), ("d_up_dst", ctypes.c_void_p * 3),
                ("ev", ctypes.c_uint64), ("hseq", ctypes.c_uint64 * 3)]


class kb_peer_op(ctypes.Structure):
    _fields_ = [("nnz_own", I64), ("d_indptr_own", c_vp), ("d_indices", c_vp), ("d_cols16", c_vp), ("d_row_pat_own", c_vp),
                ("d_pat_tab", c_vp), ("d_vals", c_vp), ("d_rowblk", c_vp)]


class kb_solve_info(ctypes.Structure):
    _fields_ = [("iterations", ctypes.c_int32), ("status", ctypes.c_int32), ("residual", F64), ("rhs_norm", F64),
                ("kernel_launches", I64), ("attainable_accuracy", ctypes.c_int32), ("restarts", ctypes.c_int32),
                ("failed_step", ctypes.c_int32), ("steps_done", ctypes.c_int32)]


KB_MODE_REFERENCE = 0
KB_MODE_CONSISTENT = 1
KB_OP_ROBIN = 1
KB_OP_CHARGALERKIN = 2
KB_ENOTCONV = -5
KB_EBREAKDOWN = -4
KB_ESTAGNATED = -6
KB_EPEER = -7

# name -> (restype, argtypes); mirrors include/kiltro_b200.h one to one
SIGNATURES = {
    "kb_error_string": (ctypes.c_char_p, [I32]),
    "kb_version": (I32, []),
    "kb_launch_count": (I64, []),
    "kb_launch_count_reset": (None, []),
    "kb_mesh_line": (I32, [F64, F64, I64, c_vp, c_vp, c_vp]),
    "kb_mesh_rectangle": (I32, [F64, F64, F64, F64, I64, I64, c_vp, c_vp, c_vp]),
    "kb_cast_i64_i32": (I32, [c_vp, I64, c_vp, c_vp]),
    "kb_element_matrices": (I32, [I32, c_vp, c_vp, I64, c_vp, ctypes.POINTER(kb_material), I32, I32, c_vp, c_vp, c_vp,
                                  c_vp]),
    "kb_pattern_build": (I32, [c_vp, I64, I32, I64, I32, c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, c_vp,
                               ctypes.POINTER(I64), c_vp]),
    "kb_pattern_grid": (I32, [I64, I64, c_vp, c_vp, c_vp, ctypes.POINTER(I64), c_vp]),
    "kb_assemble_from_elements": (I32, [I32, I64, I64, c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, c_vp]),
    "kb_tile_plan_sizes": (I32, [I64, I64, I32, I64, I64, ctypes.POINTER(I64)]),
    "kb_tile_plan_build": (I32, [c_vp, I64, I32, I64, c_vp, c_vp, c_vp, c_vp, c_vp, I64, I64, c_vp, c_vp, c_vp, c_vp,
                                 c_vp, c_vp, c_vp, ctypes.POINTER(I64), c_vp]),
    "kb_assemble_fused": (I32, [I32, c_vp, c_vp, I64, I64, c_vp, ctypes.POINTER(kb_material), I32, I32, c_vp, c_vp, I64,
                                c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, c_vp]),
    "kb_tile_tmpl_sizes": (I32, [ctypes.POINTER(I64)]),
    "kb_tile_tmpl_build": (I32, [I32, I64, I64, c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, c_vp,
                                 ctypes.POINTER(I64), c_vp]),
    "kb_assemble_fused_tmpl": (I32, [c_vp, I64, c_vp, ctypes.POINTER(kb_material), I32, I32, c_vp, I64, c_vp, c_vp, c_vp,
                                     c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, c_vp]),
    "kb_grid_verify": (I32, [c_vp, I64, I64, I64, c_vp, c_vp]),
    "kb_assemble_grid": (I32, [c_vp, c_vp, ctypes.POINTER(kb_material), I32, I32, I64, I64, c_vp, c_vp, c_vp, c_vp, c_vp,
                               c_vp]),
    "kb_line_verify": (I32, [c_vp, I64, I64, c_vp, c_vp]),
    "kb_assemble_line": (I32, [c_vp, c_vp, ctypes.POINTER(kb_material), I32, I32, I64, c_vp, c_vp, c_vp, c_vp, c_vp]),
    "kb_element_operator": (I32, [I32, I32, c_vp, c_vp, I64, c_vp, F64, F64, c_vp, c_vp, c_vp]),
    "kb_axpby": (I32, [I64, F64, c_vp, F64, c_vp, c_vp, c_vp]),
    "kb_dirichlet_split_workspace_bytes": (ctypes.c_size_t, [I64]),
    "kb_dirichlet_split": (I32, [c_vp, I64, c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, ctypes.c_size_t, ctypes.POINTER(I64), c_vp]),
    "kb_dirichlet_reduce_workspace_bytes": (ctypes.c_size_t, [I64, I64]),
    "kb_dirichlet_reduce_symbolic": (I32, [I64, c_vp, c_vp, c_vp, I64, c_vp, c_vp, c_vp, c_vp, ctypes.c_size_t,
                                           ctypes.POINTER(I64), c_vp]),
    "kb_dirichlet_reduce_numeric": (I32, [I64, c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, F64, c_vp, I64, c_vp, c_vp, c_vp, c_vp,
                                          c_vp, c_vp]),
    "kb_apply_dirichlet_reduce_workspace_bytes": (ctypes.c_size_t, [I64, I64]),
    "kb_neumann_fluxes": (I32, [c_vp, I64, c_vp, c_vp]),
    "kb_apply_dirichlet_reduce": (I32, [I64, c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, F64, c_vp, I64, c_vp, c_vp, c_vp,
                                        c_vp, c_vp, c_vp, ctypes.c_size_t, ctypes.POINTER(I64), c_vp]),
    "kb_total_component_sol": (I32, [I64, c_vp, c_vp, c_vp, c_vp, c_vp]),
    "kb_spmv_num_blocks": (I64, [I64]),
    "kb_spmv_force_simple": (None, [I32]),
    "kb_spmv_force_idx32": (None, [I32]),
    "kb_set_pdl": (None, [I32]),
    "kb_spmv_force_no_patterns": (None, [I32]),
    "kb_profile_spmv_begin": (I32, [I32]),
    "kb_profile_spmv_end": (I32, [ctypes.POINTER(I64), ctypes.POINTER(F64)]),
    "kb_spmv_partition": (I32, [I64, I64, c_vp, c_vp, c_vp]),
    "kb_spmv": (I32, [I64, I64, c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, c_vp]),
    "kb_krylov_workspace_bytes": (ctypes.c_size_t, [I64, I64]),
    "kb_cg_solve": (I32, [I64, I64, c_vp, c_vp, c_vp, c_vp, c_vp, F64, I32, I32, c_vp, ctypes.c_size_t,
                          ctypes.POINTER(kb_solve_info), c_vp]),
    "kb_bicgstab_solve": (I32, [I64, I64, c_vp, c_vp, c_vp, c_vp, c_vp, F64, I32, I32, c_vp, ctypes.c_size_t,
                                ctypes.POINTER(kb_solve_info), c_vp]),
    "kb_time_step_rule": (I32, [I32, c_vp, c_vp, I64, ctypes.POINTER(kb_material), F64, c_vp, c_vp]),
    "kb_transient_matrix": (I32, [I64, I64, c_vp, c_vp, c_vp, c_vp, c_vp, F64, c_vp, c_vp, c_vp]),
    "kb_mask_free": (I32, [I64, c_vp, c_vp, c_vp, c_vp]),
    "kb_transient_update": (I32, [I64, c_vp, c_vp, c_vp, F64, c_vp, c_vp]),
    "kb_transient_workspace_bytes": (ctypes.c_size_t, [I64, I64]),
    "kb_transient_run": (I32, [I64, I64, c_vp, c_vp, c_vp, c_vp, I32, c_vp, c_vp, c_vp, c_vp, F64, I32,
                               ctypes.POINTER(ctypes.c_int32), I32, c_vp, F64, I32, c_vp, ctypes.c_size_t,
                               ctypes.POINTER(kb_solve_info), c_vp]),
    "kb_mesh_rectangle_strip": (I32, [F64, F64, F64, F64, I64, I64, I64, I64, c_vp, c_vp, c_vp]),
    "kb_blocks_scratch_bytes": (ctypes.c_size_t, [I64]),
    "kb_make_cols16": (I32, [I64, c_vp, c_vp, c_vp, c_vp, c_vp]),
    "kb_make_row_patterns": (I32, [I64, c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, c_vp]),
    "kb_spmv_dots": (I32, [I64, I64, c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, c_vp]),
    "kb_spmv_dots4": (I32, [I64, I64, c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, c_vp]),
    "kb_bicgstab_update_p": (I32, [I64, c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, c_vp]),
    "kb_bicgstab_s2": (I32, [I64, c_vp, c_vp, c_vp, c_vp, c_vp, c_vp]),
    "kb_bicgstab_update_p2": (I32, [I64, c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, c_vp]),
    "kb_cg_update2": (I32, [I64, c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, c_vp]),
    "kb_cg_p2": (I32, [I64, c_vp, c_vp, c_vp, c_vp, c_vp, c_vp]),
    "kb_dot2": (I32, [I64, c_vp, c_vp, c_vp, c_vp, c_vp]),
    "kb_residual": (I32, [I64, c_vp, c_vp, c_vp, c_vp, c_vp, c_vp]),
    "kb_axpy_neg": (I32, [I64, c_vp, c_vp, c_vp, c_vp]),
    "kb_bicgstab_update": (I32, [I64, c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, c_vp]),
    "kb_bicgstab_p": (I32, [I64, c_vp, c_vp, c_vp, c_vp, c_vp, c_vp]),
    "kb_cg_update": (I32, [I64, c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, c_vp]),
    "kb_cg_p": (I32, [I64, c_vp, c_vp, c_vp, c_vp]),
    "kb_vec_mul": (I32, [I64, c_vp, c_vp, c_vp, c_vp]),
    "kb_diag_scale": (I32, [I64, c_vp, c_vp, c_vp, I32, c_vp, c_vp]),
    "kb_scale_matrix": (I32, [I64, c_vp, c_vp, c_vp, I32, c_vp, c_vp, c_vp]),
    "kb_masked_matrix": (I32, [I64, I64, c_vp, c_vp, c_vp, c_vp, c_vp, F64, F64, I32, c_vp, c_vp]),
    "kb_peer_alloc": (I32, [ctypes.c_size_t, ctypes.POINTER(ctypes.c_void_p), ctypes.c_char_p]),
    "kb_peer_open": (I32, [ctypes.c_char_p, ctypes.POINTER(ctypes.c_void_p)]),
    "kb_peer_close": (I32, [c_vp]),
    "kb_peer_free": (I32, [c_vp]),
    "kb_peer_error": (I32, [ctypes.POINTER(kb_peer_ctx), ctypes.POINTER(I64), c_vp]),
    "kb_peer_preload": (I32, []),
    "kb_peer_clear_error": (I32, [ctypes.POINTER(kb_peer_ctx), c_vp]),
    "kb_peer_krylov_workspace_bytes": (ctypes.c_size_t, [I64, I64]),
    "kb_peer_cg_solve": (I32, [ctypes.POINTER(kb_peer_part), ctypes.POINTER(kb_peer_op), c_vp, c_vp, c_vp, F64, I32, I32, c_vp,
                               ctypes.c_size_t, ctypes.POINTER(kb_solve_info), c_vp]),
    "kb_peer_bicgstab_solve": (I32, [ctypes.POINTER(kb_peer_part), ctypes.POINTER(kb_peer_op), c_vp, c_vp, F64, I32, I32, c_vp,
                                     ctypes.c_size_t, ctypes.POINTER(kb_solve_info), c_vp]),
    "kb_peer_transient_workspace_bytes": (ctypes.c_size_t, [I64, I64]),
    "kb_peer_transient_run": (I32, [ctypes.POINTER(kb_peer_part), ctypes.POINTER(kb_peer_op), c_vp, I32, c_vp, c_vp, c_vp, c_vp,
                                    F64, I32, F64, I32, c_vp, ctypes.c_size_t, ctypes.POINTER(kb_solve_info), c_vp]),
    "kb_mg_create": (I32, [I64, I64, I64, ctypes.POINTER(ctypes.c_void_p)]),
    "kb_mg_destroy": (None, [c_vp]),
    "kb_mg_num_levels": (I32, [c_vp]),
    "kb_mg_level_shape": (I32, [c_vp, I32, ctypes.POINTER(I64), ctypes.POINTER(I64)]),
    "kb_mg_workspace_bytes": (ctypes.c_size_t, [c_vp]),
    "kb_mg_transfer_table": (I32, [I64, I64, c_vp, c_vp, c_vp]),
    "kb_mg_setup": (I32, [c_vp, c_vp, c_vp, c_vp, ctypes.POINTER(kb_material), I32, I64, I64, c_vp, c_vp, c_vp, I32, F64,
                          c_vp, ctypes.c_size_t, c_vp]),
    "kb_mg_apply": (I32, [c_vp, c_vp, c_vp, c_vp]),
    "kb_mg_set_smoother": (I32, [c_vp, I32, ctypes.POINTER(F64)]),
    "kb_mg_set_cycle": (I32, [c_vp, I32, ctypes.POINTER(ctypes.c_int32)]),
    "kb_mg_set_fused": (None, [I32]),
    "kb_mg_set_tail": (None, [I32]),
    "kb_mg_set_assemble_tiled": (None, [I32]),
    "kb_mg_set_aggressive": (None, [I32]),
    "kb_mg_set_upwind_restriction": (None, [I32]),
    "kb_mg_set_overweight": (None, [F64]),
    "kb_mg_set_rows_per_task": (None, [I32]),
    "kb_mg_set_min_rows": (None, [I32]),
    "kb_mg_set_prefetch_rows": (None, [I32]),
    "kb_mg_set_prefetch_rows_l1": (None, [I32]),
    "kb_mg_bicgstab_workspace_bytes": (ctypes.c_size_t, [I64, I64]),
    "kb_mg_bicgstab_solve": (I32, [c_vp, I64, I64, c_vp, c_vp, c_vp, c_vp, c_vp, F64, I32, I32, c_vp, ctypes.c_size_t,
                                   ctypes.POINTER(kb_solve_info), c_vp]),
}

_LIB = None


class KiltroB200Error(RuntimeError):
    pass


def load():
    """Loads the CUDA library; raises (loudly) when it has not been built."""
    global _LIB
    if _LIB is None:
        if not os.path.isfile(LIB_PATH):
            raise KiltroB200Error(
                "libkiltro_b200.so is missing (%s). Build it with `python -m kiltro_b200.build` or "
                "__graft_entry__.build(); kiltro_b200 has no CPU fallback." % LIB_PATH)
        lib = ctypes.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(lib, name)          # AttributeError here = header/library mismatch: fail loudly
            fn.restype = res
            fn.argtypes = args
        _LIB = lib
    return _LIB


def check(rc):
    if rc != 0:
        msg = load().kb_error_string(int(rc))
        raise KiltroB200Error("libkiltro_b200 call failed (%d): %s" % (rc, msg.decode() if msg else "?"))
