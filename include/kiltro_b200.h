/* kiltro_b200.h -- C ABI of libkiltro_b200.so (B200 / sm_100a implementation of Kiltro's FEM hot path).
 *
 * Conventions:
 *   - every pointer named d_* is a DEVICE pointer owned by the caller (torch tensor storage, cudaMalloc, ...);
 *     h_* is a HOST pointer; nothing is allocated behind the caller's back except where a function says so: the
 *     once-per-mesh symbolic builders (kb_pattern_build, kb_tile_plan_build, kb_tile_tmpl_build) take their sort scratch
 *     with cudaMallocAsync and free it before returning, kb_peer_alloc creates the region it exports, kb_mg_create a host
 *     plan.  Nothing on the per-step path (assembly, boundary conditions, solvers, time loops, multigrid set-up and
 *     cycles) allocates: those functions take caller-owned workspaces sized by their *_workspace_bytes() queries;
 *   - `stream` is a cudaStream_t passed as void*; all work is stream-ordered; functions that return a count or a
 *     convergence record to the host synchronise that stream themselves and say so;
 *   - return value: 0 on success, >0 a cudaError_t, <0 a KB_E* code; kb_error_string() decodes both;
 *   - connectivity / CSR indices are int32 (the reference's native code is int32 too:
 *     kiltro/FiniteElements/ComputeSparsityPattern.pyx:15-41, Assembly.py:26), values are fp64.
 *
 * Each entry point cites the reference interface it replaces as file:line under /root/reference.
 */
#ifndef KILTRO_B200_H
#define KILTRO_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define KB_EINVAL     (-1)   /* bad argument (null pointer, unsupported element type, ...) */
#define KB_ETOOBIG    (-2)   /* a count does not fit the int32 index space */
#define KB_EWORKSPACE (-3)   /* caller-provided workspace too small */
#define KB_EBREAKDOWN (-4)   /* Krylov breakdown (rho, omega or p.Ap vanished / non-finite) */
#define KB_ENOTCONV   (-5)   /* maxiter reached without meeting the tolerance */
#define KB_EPEER      (-7)   /* a peer-memory wait timed out (a rank fell out of the partitioned iteration) */
#define KB_ESTAGNATED (-6)   /* the iteration stalled with ||b - A x|| more than 10x above rtol ||b|| (fp64 floor or lost recurrence) */

/* advection term of the element matrix */
#define KB_MODE_REFERENCE  0 /* scalar broadcast, what Temperature/TemperaturePG compute: VariationalPrinciple.py:152,156,164-173 */
#define KB_MODE_CONSISTENT 1 /* W (x) (u.grad N): MaterialBase.py:42-58 (the examples' HeatAdvectionDiffusion) */

/* element operators of kb_element_operator */
#define KB_OP_ROBIN 1        /* AssemblyRobinForces, FiniteElements/Assembly.py:121-175 */
#define KB_OP_CHARGALERKIN 2 /* GetElementalCharacteristicGalerkin, VariationalPrinciple.py:264-308 */

/* material constants, MaterialBase.py:4-39 */
typedef struct kb_material { double k, rho, c_v, area, q; } kb_material;

/* result record of the Krylov solvers */
typedef struct kb_solve_info {
    int32_t iterations;      /* iterations performed */
    int32_t status;          /* 0: TRUE residual ||b - A x||_2 <= rtol ||b||_2 of the caller's (unscaled) system, or <= 10 rtol
                                ||b||_2 with attainable_accuracy = 1; else KB_ENOTCONV, KB_EBREAKDOWN, KB_ESTAGNATED */
    double  residual;        /* final TRUE ||b - A x||_2, recomputed from the returned x in the caller's system */
    double  rhs_norm;        /* ||b||_2 of the caller's system */
    int64_t kernel_launches; /* kernels launched by the call */
    int32_t attainable_accuracy; /* 1: restarts from the true residual stopped improving it (fp64 rounding floor) */
    int32_t restarts;        /* restart cycles taken (0 = the first cycle converged and was confirmed) */
    int32_t failed_step;     /* kb_transient_run: 1-based time step whose inner solve failed and stopped the loop (0 = none) */
    int32_t steps_done;      /* kb_transient_run: time steps completed */
} kb_solve_info;

const char *kb_error_string(int code);
int kb_version(void);
/* number of kernels this library has launched since load / since the last reset (bench.py's gpu_launches) */
int64_t kb_launch_count(void);
void kb_launch_count_reset(void);

/* ---------------------------------------------------------------------------------------------------------
 * K8  mesh generators                       replaces kiltro/MeshGeneration/Mesh.py:14-30 (Line), :33-58 (Rectangle)
 * points (nnode, ndim) fp64 row-major, bit-identical to numpy.linspace/meshgrid; elements (nelem, npe) int32.
 * ------------------------------------------------------------------------------------------------------- */
int kb_mesh_line(double left, double right, int64_t n, double *d_points, int32_t *d_elements, void *stream);
int kb_mesh_rectangle(double x0, double y0, double x1, double y1, int64_t nx, int64_t ny,
                      double *d_points, int32_t *d_elements, void *stream);
/* node rows [j_first, j_first+nrows) of the global Rectangle mesh as a local mesh (row-block partition, K9) */
int kb_mesh_rectangle_strip(double x0, double y0, double x1, double y1, int64_t nx, int64_t ny, int64_t j_first,
                            int64_t nrows, double *d_points, int32_t *d_elements, void *stream);
/* Mesh.elements is int64 in the reference (Mesh.py:20,44); device connectivity is int32 */
int kb_cast_i64_i32(const int64_t *d_in, int64_t n, int32_t *d_out, void *stream);

/* ---------------------------------------------------------------------------------------------------------
 * K1  batched element matrices              replaces VariationalPrinciple.py:42-69 (GetLocalMass), :117-173
 *     (Temperature.GetLocalConvection), :216-261,:311-342 (TemperaturePG), FunctionSpace.py:5-62 (tables)
 * ndim 1 -> line-2 (npe 2, 2 Gauss points), ndim 2 -> quad-4 (npe 4, 2x2 Gauss points).
 * d_velocity may be NULL (zero velocity).  d_Ke (nelem, npe*npe) row-major = convectionel.ravel();
 * d_Fe (nelem, npe); d_Me (nelem, npe*npe) or NULL (steady: Me is not computed, VariationalPrinciple.py:104-107).
 * ------------------------------------------------------------------------------------------------------- */
int kb_element_matrices(int ndim, const double *d_points, const int32_t *d_elements, int64_t nelem,
                        const double *d_velocity, const kb_material *material, int mode, int petrov,
                        double *d_Ke, double *d_Fe, double *d_Me, void *stream);

/* ---------------------------------------------------------------------------------------------------------
 * K2  symbolic phase                        replaces FiniteElements/ComputeSparsityPattern.pyx:44-112 and
 *     ComputeSparsityPattern.h:22-62 (_ComputeSparsityPattern_), :67-122 (_ComputeDataIndices_)
 * COO keys of every element entry -> device LSD radix sort -> unique.  nent = nelem*(npe*nvar)^2.
 * Outputs (all device, caller-allocated):
 *   d_indptr (nnode*nvar+1), d_indices (capacity nent, first nnz valid; like .pyx:64,91),
 *   d_data_local (nent), d_data_global (nent)          -- the reference's two scatter maps;
 *   d_seg_ptr (nent+1, first nnz+1 valid), d_gather_src (nent) -- row-owner gather map: CSR entry u sums
 *        Ke_flat[d_gather_src[p]] for p in [seg_ptr[u], seg_ptr[u+1]) ; contributors are in ascending element
 *        order, the order of the reference's sequential += (Assembly.py:46,60-61);
 *   d_diag_pos (nnode*nvar) position of the diagonal entry of each row (-1 if absent).
 * *h_nnz receives nnz (the call synchronises `stream`).
 * ------------------------------------------------------------------------------------------------------- */
int kb_pattern_build(const int32_t *d_elements, int64_t nelem, int npe, int64_t nnode, int nvar,
                     int32_t *d_indptr, int32_t *d_indices, int32_t *d_data_local, int32_t *d_data_global,
                     int32_t *d_seg_ptr, int32_t *d_gather_src, int32_t *d_diag_pos, int64_t *h_nnz, void *stream);
/* The same indptr / indices / diag_pos for a quad-4 mesh numbered like Mesh.Rectangle (verified with kb_grid_verify),
 * written down from the shape without forming or sorting the element entries; *h_nnz = (3 nxn - 2)(3 nyn - 2); stream-ordered,
 * no synchronisation.  d_indices must hold that many entries.  The reference's two scatter maps and the gather map still
 * come from kb_pattern_build (the host mirror builds them lazily, on first use). */
int kb_pattern_grid(int64_t nxn, int64_t nyn, int32_t *d_indptr, int32_t *d_indices, int32_t *d_diag_pos, int64_t *h_nnz,
                    void *stream);

/* ---------------------------------------------------------------------------------------------------------
 * K3  numeric assembly                      replaces FiniteElements/Assembly.py:46-72 (element loop, scatter-add)
 * (a) two-step seam: CSR values from materialised element matrices through the gather map (deterministic segmented
 *     reduction, no atomics).  d_Me/d_M may be NULL.  d_Flux (nnode) from d_Fe through the diagonal's segment.
 * ------------------------------------------------------------------------------------------------------- */
int kb_assemble_from_elements(int npe, int64_t nnode, int64_t nnz, const int32_t *d_seg_ptr,
                              const int32_t *d_gather_src, const int32_t *d_diag_pos,
                              const double *d_Ke, const double *d_Me, const double *d_Fe,
                              double *d_K, double *d_M, double *d_Flux, void *stream);

/* (b) fused path: element matrices are computed on chip and never written to HBM.  A "tile plan" built once per mesh
 *     (symbolic; layout in csrc/tileplan.cuh) assigns to each CTA a set of CSR rows and the elements incident to them.
 *     kb_tile_plan_sizes -> h_sizes = {ntiles, capacity of d_tile_elems, capacity of d_pair_rec, structured?}.
 *     nx_nodes/ny_nodes: node grid of a Mesh.Rectangle-numbered mesh (tiling hint; pass 0,0 for arbitrary meshes).
 *     kb_tile_plan_build -> h_info = {ntiles, n tile elements, n pairs, max elements/tile, max rows/tile,
 *     max entries/tile, fits (1 = usable by kb_assemble_fused), structured}; synchronises `stream`.
 *     d_tile_rows (nnode), d_tile_row_ptr / d_tile_elem_ptr (ntiles+1), d_pair_ptr (nnode+1); d_tile_elems (element ids,
 *     build scratch) and d_tile_conn (npe ints per tile element: tile-local copy of the connectivity, what the kernel
 *     reads) have capacity h_sizes[1] and npe*h_sizes[1]. */
int kb_tile_plan_sizes(int64_t nnode, int64_t nelem, int npe, int64_t nx_nodes, int64_t ny_nodes, int64_t *h_sizes);
int kb_tile_plan_build(const int32_t *d_elements, int64_t nelem, int npe, int64_t nnode,
                       const int32_t *d_indptr, const int32_t *d_indices, const int32_t *d_seg_ptr,
                       const int32_t *d_gather_src, const int32_t *d_diag_pos, int64_t nx_nodes, int64_t ny_nodes,
                       int32_t *d_tile_rows, int32_t *d_tile_row_ptr, int32_t *d_tile_elems, int32_t *d_tile_conn,
                       int32_t *d_tile_elem_ptr, uint32_t *d_pair_rec, int32_t *d_pair_ptr, int64_t *h_info,
                       void *stream);
/* d_M may be NULL (steady).  d_velocity may be NULL.  d_scratch: ntiles + 32 bytes (flags of the tiles holding a quad the
 * fast element routine declines -- non-convex / inverted -- which a second, normally empty, launch redoes). */
int kb_assemble_fused(int ndim, const double *d_points, const int32_t *d_elements, int64_t nelem, int64_t nnode,
                      const double *d_velocity, const kb_material *material, int mode, int petrov,
                      const int32_t *d_indptr, const int32_t *d_indices,
                      int64_t ntiles, const int32_t *d_tile_rows, const int32_t *d_tile_row_ptr,
                      const int32_t *d_tile_conn, const int32_t *d_tile_elem_ptr,
                      const uint32_t *d_pair_rec, const int32_t *d_pair_ptr,
                      double *d_K, double *d_M, double *d_Flux, void *d_scratch, void *stream);

/* (c) fused path, template form (quad-4 meshes whose tiles repeat: everything Mesh.Rectangle generates,
 *     MeshGeneration/Mesh.py:33-58, and row-block strips of it).  kb_tile_tmpl_build classifies the tiles of a plan by
 *     their LOCAL structure (layout in csrc/tileplan.cuh) and stores one template per class; kb_assemble_fused_tmpl then
 *     reads, per tile, only the tile-local connectivity and one (row, indptr) pair per run of rows.  Same summation order
 *     and results as kb_assemble_fused.
 *     kb_tile_tmpl_sizes -> h_sizes = {bytes per template, max templates, max runs per tile}.
 *     kb_tile_tmpl_build: d_tile_class (ntiles bytes), d_trun_ptr (ntiles+1), d_trun_row (capacity nnode),
 *     d_templates (max templates x bytes per template); h_info = {usable (1/0), templates, runs}; synchronises.
 *     kb_assemble_fused_tmpl: d_scratch = ntiles + 32 bytes (flags of the tiles holding an element the fast element
 *     routine declines -- non-convex / inverted quads -- which a second, normally empty, launch redoes). */
int kb_tile_tmpl_sizes(int64_t *h_sizes);
int kb_tile_tmpl_build(int npe, int64_t nnode, int64_t ntiles, const int32_t *d_indptr, const int32_t *d_tile_rows,
                       const int32_t *d_tile_row_ptr, const int32_t *d_tile_elem_ptr, const uint32_t *d_pair_rec,
                       const int32_t *d_pair_ptr, uint8_t *d_tile_class, int32_t *d_trun_ptr, int32_t *d_trun_row,
                       void *d_templates, int64_t *h_info, void *stream);
int kb_assemble_fused_tmpl(const double *d_points, int64_t nnode, const double *d_velocity,
                           const kb_material *material, int mode, int petrov, const int32_t *d_indptr,
                           int64_t ntiles, const int32_t *d_tile_conn, const int32_t *d_tile_elem_ptr,
                           const uint8_t *d_tile_class, const int32_t *d_trun_ptr, const int32_t *d_trun_row,
                           const void *d_templates, double *d_K, double *d_M, double *d_Flux, void *d_scratch,
                           void *stream);

/* (d) fused path, grid form (quad-4 meshes numbered like Mesh.Rectangle, MeshGeneration/Mesh.py:33-58: node (i, j) =
 *     j * nxn + i, element (i, j) = [c, c+1, c+nxn+1, c+nxn] with c = j * nxn + i; row-block strips of such a mesh too).
 *     kb_grid_verify checks the connectivity against that numbering (*d_flag = 0: it matches).  kb_assemble_grid then
 *     needs neither the connectivity nor the column indices: a warp sweeps a strip of node columns upwards, element
 *     matrices stay in registers and neighbouring lanes exchange them by shuffles (csrc/assemble_grid.cu).  Same
 *     summation order and results as kb_assemble_fused.  d_M / d_velocity may be NULL.  d_scratch: 32 bytes. */
int kb_grid_verify(const int32_t *d_elements, int64_t nelem, int64_t nxn, int64_t nyn, int32_t *d_flag, void *stream);
int kb_assemble_grid(const double *d_points, const double *d_velocity, const kb_material *material, int mode, int petrov,
                     int64_t nxn, int64_t nyn, const int32_t *d_indptr, double *d_K, double *d_M, double *d_Flux,
                     void *d_scratch, void *stream);

/* (e) fused path, grid form for LINE meshes numbered like Mesh.Line (MeshGeneration/Mesh.py:14-30: element i = [i, i+1]).
 *     kb_line_verify checks the connectivity (*d_flag = 0: it matches); kb_assemble_line evaluates, per node, the element
 *     matrices of its two elements in registers (VariationalPrinciple.py:42-69,117-161) and writes the node's tridiagonal
 *     CSR row; contributions are added in ascending element order (Assembly.py:46,60-61), results equal
 *     kb_assemble_fused's to rounding.  d_M / d_velocity may be NULL. */
int kb_line_verify(const int32_t *d_elements, int64_t nelem, int64_t nnode, int32_t *d_flag, void *stream);
int kb_assemble_line(const double *d_points, const double *d_velocity, const kb_material *material, int mode, int petrov,
                     int64_t nnode, const int32_t *d_indptr, double *d_K, double *d_M, double *d_Flux, void *stream);

/* ---------------------------------------------------------------------------------------------------------
 * K10 element operators of the "next" rows (seam form: element matrices to HBM, then kb_assemble_from_elements)
 *   KB_OP_ROBIN         replaces AssemblyRobinForces (FiniteElements/Assembly.py:121-175): Ke = h sum_g N (x) N dV,
 *                       Fe as shipped (:157,164); p0 = h (boundary_condition.applied_robin_convection), p1 = T_inf
 *                       (ground_robin_convection); d_velocity ignored.
 *   KB_OP_CHARGALERKIN  replaces AssembleCharacteristicGalerkin (Assembly.py:95-117) over
 *                       GetElementalCharacteristicGalerkin (VariationalPrinciple.py:264-308): Ke = sum_g (u.grad N) (x)
 *                       (u.grad N) dV, Fe = p0 sum_g (u.grad N) dV with p0 = material.q * material.area; p1 unused.
 *   d_Ke (nelem x npe^2, row-major), d_Fe (nelem x npe).
 * kb_axpby: out = a x + b y (y may be NULL); folds an operator into K / F on the shared sparsity pattern.
 * ------------------------------------------------------------------------------------------------------- */
int kb_element_operator(int ndim, int op, const double *d_points, const int32_t *d_elements, int64_t nelem,
                        const double *d_velocity, double p0, double p1, double *d_Ke, double *d_Fe, void *stream);
int kb_axpby(int64_t n, double a, const double *d_x, double b, const double *d_y, double *d_out, void *stream);

/* ---------------------------------------------------------------------------------------------------------
 * K4  boundary conditions                   replaces kiltro/BoundaryCondition.py
 * ------------------------------------------------------------------------------------------------------- */
/* :100-122 GetDirichletBoundaryConditions.  flags (ndof) fp64, NaN = free.  d_renum[i] = index of DOF i in the reduced
 * system if free, -1 - (index among the fixed DOFs) otherwise.  d_columns_in / d_columns_out (capacity ndof each, int64
 * like the reference), d_applied (capacity ndof).  d_prev_renum (optional, != d_renum): the renumbering of an earlier
 * call -- h_counts[2] = 1 when the new one equals it (callers keep their cached reduced structure).  d_workspace:
 * kb_dirichlet_split_workspace_bytes(ndof) bytes.  h_counts[0]=n_free, h_counts[1]=n_fixed (synchronises). */
size_t kb_dirichlet_split_workspace_bytes(int64_t ndof);
int kb_dirichlet_split(const double *d_flags, int64_t ndof, const int32_t *d_prev_renum, int32_t *d_renum,
                       int64_t *d_columns_in, int64_t *d_columns_out, double *d_applied, void *d_workspace,
                       size_t workspace_bytes, int64_t *h_counts, void *stream);
/* :125-140 ComputeNeumannFluxes: F[i] = flag[i] if not NaN else 0 */
int kb_neumann_fluxes(const double *d_flags, int64_t ndof, double *d_F, void *stream);
/* :169-200 ApplyDirichletGetReducedMatrices, in two phases.  d_applied (n_fixed): prescribed values in columns_out order
 * (what kb_dirichlet_split wrote, possibly scaled by the caller's load increment).
 *   symbolic (once per sparsity pattern + Dirichlet set; synchronises): row pointers d_Kb_indptr (n_free+1) and column
 *     indices d_Kb_indices (capacity nnz) of K_b = K[in,:][:,in], and one 32-bit survival mask per row of K
 *     (d_row_mask, ndof words).  h_info[0] = nnz(K_b); h_info[1] = 1 when every free row has <= 32 entries, i.e. the masks
 *     can drive the numeric phase (otherwise use the one-shot form below).  d_workspace: *_workspace_bytes(ndof, n_free).
 *   numeric (every step; asynchronous, no allocation, no host round trip): d_Kb_data / d_Mb_data (nnz(K_b)) and, in place,
 *       F[free] -= (sum_{j fixed, !isclose(T_j,0)} K_ij T_j, ascending j) * load_factor ;  d_Fb (n_free) = F[in].
 *     d_Kb_data = NULL: right-hand side only.  Column indices are read only in rows that lose an entry.
 *   one-shot (kb_apply_dirichlet_reduce: any row length, structure recomputed on every call; synchronises for nnz). */
size_t kb_dirichlet_reduce_workspace_bytes(int64_t ndof, int64_t n_free);
int kb_dirichlet_reduce_symbolic(int64_t ndof, const int32_t *d_indptr, const int32_t *d_indices, const int32_t *d_renum,
                                 int64_t n_free, uint32_t *d_row_mask, int32_t *d_Kb_indptr, int32_t *d_Kb_indices,
                                 void *d_workspace, size_t workspace_bytes, int64_t *h_info, void *stream);
int kb_dirichlet_reduce_numeric(int64_t ndof, const int32_t *d_indptr, const int32_t *d_indices, const double *d_K,
                                const double *d_M, const int32_t *d_renum, const double *d_applied, double load_factor,
                                double *d_F, int64_t n_free, const uint32_t *d_row_mask, const int32_t *d_Kb_indptr,
                                double *d_Kb_data, double *d_Mb_data, double *d_Fb, void *stream);
size_t kb_apply_dirichlet_reduce_workspace_bytes(int64_t ndof, int64_t n_free);
int kb_apply_dirichlet_reduce(int64_t ndof, const int32_t *d_indptr, const int32_t *d_indices, const double *d_K,
                              const double *d_M, const int32_t *d_renum, const double *d_applied, double load_factor,
                              double *d_F, int64_t n_free, int32_t *d_Kb_indptr, int32_t *d_Kb_indices,
                              double *d_Kb_data, double *d_Mb_data, double *d_Fb, void *d_workspace,
                              size_t workspace_bytes, int64_t *h_nnz_b, void *stream);
/* :203-241 TotalComponentSol / UpdateFixDoFs / UpdateFreeDoFs: T[i] = sol[renum[i]] if free (or 0 when d_sol NULL),
 * applied[-1-renum[i]] if fixed (or 0 when d_applied NULL) */
int kb_total_component_sol(int64_t ndof, const int32_t *d_renum, const double *d_sol, const double *d_applied,
                           double *d_T, void *stream);

/* ---------------------------------------------------------------------------------------------------------
 * K5  CSR SpMV y = A x                      (building block of K6; replaces nothing 1:1 -- the reference's only
 *     mat-vec is the dense np.dot in FEMSolver.py:281)
 * d_rowblk (nblk+1): row partition with ~equal nnz per CTA from kb_spmv_partition (nblk = kb_spmv_num_blocks).
 * d_dot_x / d_dot_out: optional fused dot products: out[0] = <w, y>, out[1] = <y, y> where w = d_dot_w (may be NULL).
 * ------------------------------------------------------------------------------------------------------- */
int64_t kb_spmv_num_blocks(int64_t nnz);
/* live timing of the SpMV launches inside kb_cg_solve / kb_bicgstab_solve: between begin and end, one SpMV launch of every
 * `every`-th iteration is bracketed by CUDA events on the launching stream (bench.py's roofline leg). */
void kb_spmv_force_simple(int on);   /* development switch: 1 = use the one-shot (non-TMA) SpMV kernel */
void kb_spmv_force_idx32(int on);    /* development switch: 1 = keep 32-bit column indices inside the solvers */
void kb_spmv_force_no_patterns(int on); /* development switch: 1 = no row-pattern form (csrc/spmv_pat.cuh) in the solvers */
void kb_set_pdl(int on);             /* development switch: 0 = no programmatic dependent launch in the Krylov loops (default 1) */
int kb_profile_spmv_begin(int every);
int kb_profile_spmv_end(int64_t *h_samples, double *h_total_ms);
int kb_spmv_partition(int64_t nrows, int64_t nnz, const int32_t *d_indptr, int32_t *d_rowblk, void *stream);
int kb_spmv(int64_t nrows, int64_t nnz, const int32_t *d_indptr, const int32_t *d_indices, const double *d_vals,
            const int32_t *d_rowblk, const double *d_x, double *d_y, void *stream);

/* ---------------------------------------------------------------------------------------------------------
 * K6  Krylov solvers                        replace LinearSolver.Solve -> scipy spsolve, FEMSolver.py:350-364
 * Jacobi preconditioning is applied as an explicit diagonal scaling of a private copy of the values
 * (CG: D^-1/2 A D^-1/2, BiCGSTAB: A D^-1), which is algebraically the Jacobi-preconditioned method.
 * workspace: kb_krylov_workspace_bytes(nrows, nnz) bytes of device memory.  x: initial guess in, solution out.
 * The calls synchronise `stream` every `check_every` iterations to test convergence ||r|| <= rtol*||b||.
 * ------------------------------------------------------------------------------------------------------- */
size_t kb_krylov_workspace_bytes(int64_t nrows, int64_t nnz);
int kb_cg_solve(int64_t nrows, int64_t nnz, const int32_t *d_indptr, const int32_t *d_indices, const double *d_vals,
                const double *d_b, double *d_x, double rtol, int maxiter, int check_every,
                void *d_workspace, size_t workspace_bytes, kb_solve_info *h_info, void *stream);
int kb_bicgstab_solve(int64_t nrows, int64_t nnz, const int32_t *d_indptr, const int32_t *d_indices,
                      const double *d_vals, const double *d_b, double *d_x, double rtol, int maxiter, int check_every,
                      void *d_workspace, size_t workspace_bytes, kb_solve_info *h_info, void *stream);

/* ---------------------------------------------------------------------------------------------------------
 * K7  transient loop                        replaces FEMSolver.TransientSolver, FEMSolver.py:253-297 (+ :138-145)
 * Every step solves, on the FULL system with the reference's (M^-1)[in,in] convention,
 *     (M + theta dt Kff) y = Kff T^n + Fm ;   T^{n+1} = T^n - dt y at free DOFs, prescribed values at fixed DOFs
 * (Kff: rows and columns of fixed DOFs zeroed; Fm: F zeroed at fixed DOFs).  theta = 0 is the reference's intended
 * forward Euler; theta > 0 is new capability.
 * ------------------------------------------------------------------------------------------------------- */
/* dt = factor * min_e rho c_v |X1-X0|^2 / (2k)   (FEMSolver.py:138-144); result written to *d_dt (device) */
int kb_time_step_rule(int ndim, const double *d_points, const int32_t *d_elements, int64_t nelem,
                      const kb_material *material, double factor, double *d_dt, void *stream);
/* Kff and A = M + theta_dt*Kff on the shared pattern */
int kb_transient_matrix(int64_t ndof, int64_t nnz, const int32_t *d_indptr, const int32_t *d_indices,
                        const double *d_K, const double *d_M, const int32_t *d_renum, double theta_dt,
                        double *d_Kff, double *d_A, void *stream);
/* out = in at free DOFs, 0 at fixed DOFs */
int kb_mask_free(int64_t ndof, const int32_t *d_renum, const double *d_in, double *d_out, void *stream);
/* one update: T = prescribed value at fixed DOFs, T - dt*y at free DOFs (in place) */
int kb_transient_update(int64_t ndof, const int32_t *d_renum, const double *d_applied, const double *d_y, double dt,
                        double *d_T, void *stream);
/* The whole time loop on device.  d_T: T^0 in, T^nsteps out.  h_snap_steps: ascending step indices (0 = initial state)
 * whose state is copied to d_snapshots (nsnap x ndof).  a_symmetric selects CG (1) or BiCGSTAB (0) for A.  h_info:
 * iterations = total Krylov iterations, residual = worst relative residual over the steps.  Synchronises `stream`. */
size_t kb_transient_workspace_bytes(int64_t ndof, int64_t nnz);
int kb_transient_run(int64_t ndof, int64_t nnz, const int32_t *d_indptr, const int32_t *d_indices, const double *d_Kff,
                     const double *d_A, int a_symmetric, const double *d_Fm, const int32_t *d_renum,
                     const double *d_applied, double *d_T, double dt, int nsteps, const int32_t *h_snap_steps, int nsnap,
                     double *d_snapshots, double rtol, int maxiter, void *d_workspace, size_t workspace_bytes,
                     kb_solve_info *h_info, void *stream);

/* ---------------------------------------------------------------------------------------------------------
 * K9  building blocks of the row-block partitioned (multi-GPU) solves           (new capability)
 * The fused kernels of K6 exposed one by one, scalars as DEVICE pointers, local partial dots returned in d_out2 so the
 * caller can all-reduce them (NCCL) between the dot and its use.  d_scratch: kb_blocks_scratch_bytes(nnz) bytes, zeroed
 * once by the caller.  Row-range views of a local CSR are expressed by offsetting d_indptr / d_y / vector pointers.
 * ------------------------------------------------------------------------------------------------------- */
size_t kb_blocks_scratch_bytes(int64_t nnz);
/* d_cols16 (optional, from kb_make_cols16): 16-bit column offsets col - row; when given, the kernel streams them instead
 * of d_indices (10 instead of 12 bytes per non-zero) and d_x must point at the x entry of the FIRST ROW of the view. */
int kb_make_cols16(int64_t nrows, const int32_t *d_indptr, const int32_t *d_indices, int16_t *d_cols16, int32_t *d_flag,
                   void *stream);                                         /* *d_flag = 1 when an offset does not fit */
/* d_row_pat / d_pat_tab (optional, from kb_make_row_patterns; take precedence over d_cols16): row-pattern form of
 * csrc/spmv_pat.cuh -- one byte per row names its tuple of col - row offsets, the <= 256 tuples live in shared memory, no
 * column index is streamed (8 bytes per non-zero).  Same convention for d_x as with d_cols16; the view's first row must
 * be a multiple of 16 (bulk-copy alignment of the pattern bytes).
 * kb_make_row_patterns: d_row_pat (nrows rounded up to 16, + 16 bytes), d_pat_keys (256 x 8 bytes of scratch), d_pat_tab
 * (256 x 16 int16); *d_flag = 1 when the matrix has more than 256 patterns, a row longer than 16 or an offset beyond 16 bits. */
int kb_make_row_patterns(int64_t nrows, const int32_t *d_indptr, const int32_t *d_indices, uint8_t *d_row_pat,
                         void *d_pat_keys, int16_t *d_pat_tab, int32_t *d_flag, void *stream);
int kb_spmv_dots(int64_t nrows, int64_t nnz, const int32_t *d_indptr, const int32_t *d_indices, const int16_t *d_cols16,
                 const uint8_t *d_row_pat, const int16_t *d_pat_tab,
                 const double *d_vals, const int32_t *d_rowblk, const double *d_x, double *d_y, const double *d_w,
                 double *d_out2, void *d_scratch, void *stream);          /* y = A x ; out = {<w,y>, <y,y>} */
int kb_spmv_dots4(int64_t nrows, int64_t nnz, const int32_t *d_indptr, const int32_t *d_indices, const int16_t *d_cols16,
                  const uint8_t *d_row_pat, const int16_t *d_pat_tab,
                  const double *d_vals, const int32_t *d_rowblk, const double *d_x, double *d_y, const double *d_w,
                  const double *d_w2, double *d_out4, void *d_scratch, void *stream);   /* {<w,y>,<y,y>,<w2,y>,<w2,w>} */
/* merged BiCGSTAB update: x += alpha p + omega s ; r = s - omega t ; p = r + beta (p - omega v) ; out2[0] = <r,r> */
int kb_bicgstab_update_p(int64_t n, const double *d_alpha, const double *d_omega, const double *d_beta,
                         const double *d_t, const double *d_v, double *d_x, double *d_r, double *d_p, double *d_out2,
                         void *d_scratch, void *stream);
/* scalar-fused forms (no scalar kernels between a collective and the vector work): every thread derives alpha / omega /
 * beta from the all-reduced dots; rho is double-buffered by the caller (d_rho_in != d_rho_out). */
int kb_bicgstab_s2(int64_t n, const double *d_rho, const double *d_dots, const double *d_v, double *d_r,
                   double *d_alpha_out, void *stream);                    /* alpha = rho/dots[0] ; r -= alpha v */
int kb_bicgstab_update_p2(int64_t n, const double *d_rho_in, double *d_rho_out, const double *d_alpha,
                          const double *d_dots4, const double *d_t, const double *d_v, double *d_x, double *d_r,
                          double *d_p, double *d_out2, void *d_scratch, void *stream);
int kb_cg_update2(int64_t n, const double *d_rho, const double *d_dots, const double *d_p, const double *d_q, double *d_x,
                  double *d_r, double *d_out2, void *d_scratch, void *stream);   /* alpha = rho/dots[0] */
int kb_cg_p2(int64_t n, const double *d_rho_in, double *d_rho_out, const double *d_dots, const double *d_r, double *d_p,
             void *stream);                                               /* beta = dots[0]/rho_in ; rho_out = dots[0] */
int kb_dot2(int64_t n, const double *d_a, const double *d_b, double *d_out2, void *d_scratch, void *stream); /* {<a,b>,<b,b>} */
int kb_residual(int64_t n, const double *d_b, const double *d_q, double *d_r, double *d_out2, void *d_scratch,
                void *stream);                                            /* r = b - q ; out = {<r,r>, <b,b>} */
int kb_axpy_neg(int64_t n, const double *d_alpha, const double *d_v, double *d_r, void *stream);      /* r -= alpha v */
int kb_bicgstab_update(int64_t n, const double *d_alpha, const double *d_omega, const double *d_p, const double *d_t,
                       const double *d_rhat, double *d_x, double *d_r, double *d_out2, void *d_scratch, void *stream);
int kb_bicgstab_p(int64_t n, const double *d_beta, const double *d_omega, const double *d_r, const double *d_v,
                  double *d_p, void *stream);                             /* p = r + beta (p - omega v) */
int kb_cg_update(int64_t n, const double *d_alpha, const double *d_p, const double *d_q, double *d_x, double *d_r,
                 double *d_out2, void *d_scratch, void *stream);          /* x += alpha p ; r -= alpha q ; out[0] = <r,r> */
int kb_cg_p(int64_t n, const double *d_beta, const double *d_r, double *d_p, void *stream);           /* p = r + beta p */
int kb_vec_mul(int64_t n, const double *d_a, const double *d_b, double *d_out, void *stream);         /* out = a .* b */
int kb_diag_scale(int64_t nrows, const int32_t *d_indptr, const int32_t *d_indices, const double *d_vals, int mode,
                  double *d_scale, void *stream);                         /* mode 0: 1/diag ; 1: 1/sqrt(diag) */
int kb_scale_matrix(int64_t nrows, const int32_t *d_indptr, const int32_t *d_indices, const double *d_vals, int mode,
                    const double *d_scale, double *d_out, void *stream);  /* mode 0: A D ; 1: D A D (D = diag(scale)) */
/* out = c_k*Kff + c_m*M (M may be NULL) (+ 1 on the diagonal of fixed rows when identity_fixed) */
int kb_masked_matrix(int64_t ndof, int64_t nnz, const int32_t *d_indptr, const int32_t *d_indices, const double *d_K,
                     const double *d_M, const int32_t *d_renum, double c_k, double c_m, int identity_fixed,
                     double *d_out, void *stream);

/* ---------------------------------------------------------------------------------------------------------
 * K9b peer-memory exchange (NVLink / NVSwitch): the partitioned BiCGSTAB iteration without NCCL calls
 * A rank allocates one IPC-exportable region (control block of 128 words + the SpMV input vectors), all ranks exchange the
 * 64-byte handles (any transport) and map each other's regions.  Kernels publish partial dot products into every rank's
 * control block and write their boundary mesh rows straight into the neighbours' halo slots; consumers wait on sequence
 * flags (csrc/peer.cuh).  Every wait times out into an error word (kb_peer_error) instead of hanging.
 * ------------------------------------------------------------------------------------------------------- */
typedef struct kb_peer_ctx { int32_t nranks, rank; void *ctrl[8]; } kb_peer_ctx;   /* control blocks as mapped locally */
int kb_peer_alloc(size_t bytes, void **d_ptr, unsigned char *h_handle64);          /* cudaMalloc + zero + IPC handle */
int kb_peer_open(const unsigned char *h_handle64, void **d_ptr);
int kb_peer_close(void *d_ptr);
int kb_peer_free(void *d_ptr);
int kb_peer_error(const kb_peer_ctx *ctx, int64_t *h_err, void *stream);           /* 0 ok, 1/2: a wait timed out */
int kb_peer_clear_error(const kb_peer_ctx *ctx, void *stream);
/* Loads every kernel of the partitioned loops (CUDA loads kernels lazily, and a load may wait for running kernels: ranks
 * that share one process / device must not meet that while a peer's kernel waits for them).  Synchronises the device. */
int kb_peer_preload(void);

/* One rank's share of a partitioned system: the mapped control blocks, the rank's exchanged vectors inside its region
 * (full local vectors: [halo row below | owned rows | halo row above]; 0 = p, 1 = r -- the SpMV inputs of the Krylov loops --
 * 2 = T of the time loop), the neighbours' halo slots of each, and the running event / halo sequence numbers (the calls
 * advance them; every rank makes the same calls, so they stay equal across ranks). */
typedef struct kb_peer_part {
    kb_peer_ctx ctx;
    int64_t n_loc, n_own, own_off, nxn;   /* local / owned DOFs, first owned entry, DOFs per mesh row (= halo row length) */
    void *d_vec[3];
    void *d_dn_dst[3];                    /* halo-ABOVE slot of rank-1 for each vector (NULL on rank 0) */
    void *d_up_dst[3];                    /* halo-BELOW slot of rank+1 (NULL on the last rank) */
    uint64_t ev;
    uint64_t hseq[3];
} kb_peer_part;
/* Owned-row view of the rank's local CSR (columns in local numbering).  d_cols16 / d_row_pat_own + d_pat_tab optional
 * (kb_make_cols16 / kb_make_row_patterns); d_indptr_own = indptr + own_off, d_row_pat_own = row_pat + own_off. */
typedef struct kb_peer_op {
    int64_t nnz_own;
    const int32_t *d_indptr_own, *d_indices;
    const int16_t *d_cols16;
    const uint8_t *d_row_pat_own;
    const int16_t *d_pat_tab;
    const double  *d_vals;                /* the operator the loop iterates on (already Jacobi-scaled) */
    const int32_t *d_rowblk;              /* kb_spmv_partition of the owned-row view */
} kb_peer_op;
/* Partitioned Jacobi-Krylov solves of the scaled system  Ahat xhat = bhat  (extends LinearSolver.Solve, FEMSolver.py:350-364,
 * to a mesh split into strips): every halo row and every dot product travels through peer memory inside the kernels; the
 * host reads its own device's state every check_every iterations; restarts from the true residual and verdict as
 * kb_cg_solve / kb_bicgstab_solve.  d_b_own / d_x_own: owned parts (x: initial guess in, solution out).  d_scale_own (CG,
 * optional): D^-1/2, so that the verdict is taken on the unscaled residual.  Returns KB_EPEER when a wait timed out. */
size_t kb_peer_krylov_workspace_bytes(int64_t n_own, int64_t nnz_own);
int kb_peer_cg_solve(kb_peer_part *part, const kb_peer_op *op, const double *d_b_own, double *d_x_own,
                     const double *d_scale_own, double rtol, int maxiter, int check_every, void *d_workspace,
                     size_t workspace_bytes, kb_solve_info *h_info, void *stream);
int kb_peer_bicgstab_solve(kb_peer_part *part, const kb_peer_op *op, const double *d_b_own, double *d_x_own, double rtol,
                           int maxiter, int check_every, void *d_workspace, size_t workspace_bytes, kb_solve_info *h_info,
                           void *stream);
/* Partitioned theta-method loop (extends FEMSolver.TransientSolver, FEMSolver.py:253-297): nsteps steps of
 *   (M + theta dt Kff) y = Kff T^n + Fm ;  T^{n+1} = T^n - dt y / prescribed values,
 * T in part->d_vec[2] (owned part; halo rows are exchanged by the kernels).  op->d_vals: the Jacobi-scaled inner operator,
 * d_scale_own its scaling vector (D^-1/2 when a_symmetric, else D^-1), d_Kff_vals: Kff (unscaled) on the same pattern. */
size_t kb_peer_transient_workspace_bytes(int64_t n_own, int64_t nnz_own);
int kb_peer_transient_run(kb_peer_part *part, const kb_peer_op *op, const double *d_Kff_vals, int a_symmetric,
                          const double *d_scale_own, const double *d_Fm_own, const int32_t *d_renum_own,
                          const double *d_applied, double dt, int nsteps, double rtol, int maxiter, void *d_workspace,
                          size_t workspace_bytes, kb_solve_info *h_info, void *stream);

/* ---------------------------------------------------------------------------------------------------------
 * Opt-in solver beside the prescribed Jacobi method: BiCGSTAB right-preconditioned by ONE geometric-multigrid V-cycle,
 * for quad-4 meshes numbered like Mesh.Rectangle (MeshGeneration/Mesh.py:33-58).  Replaces the same call as
 * kb_bicgstab_solve: LinearSolver.Solve -> scipy.sparse.linalg.spsolve, FEMSolver.py:350-364.  csrc/multigrid.cu,
 * csrc/krylov_mg.cuh; host mirror kiltro_b200/multigrid.py (LinearSolver(preconditioner=GridMultigrid(...))).
 *
 * kb_mg is a HOST-side plan (level shapes, 1-D transfer tables, device pointers into the caller's workspace); it owns no
 * device memory.  Levels: nxn x nyn nodes, then about half the elements per direction per level down to <= coarsest_nodes
 * (<= 1023) nodes.  kb_mg_setup builds the operators: level 0 from the reduced CSR matrix K_b of
 * kb_apply_dirichlet_reduce (every entry must lie in the 9-point neighbourhood of its row: else KB_EINVAL), the coarser
 * ones rediscretised with the path's element routine and the Petrov-Galerkin weights of TemperaturePG
 * (VariationalPrinciple.py:216-261,311-342; d_velocity NULL = pure diffusion; with a velocity `mode` must be
 * KB_MODE_CONSISTENT); coordinates / velocities / Dirichlet flags of a coarse node are those of the nearest finer node.
 * d_renum: the renumbering of kb_dirichlet_split (reduced index, or < 0 for a prescribed DOF).  The workspace
 * (kb_mg_workspace_bytes) must stay untouched between kb_mg_setup and the last apply / solve; passing the same pointer to
 * a later kb_mg_setup (new values, same mesh) skips the one-off initialisation.  kb_mg_setup synchronises the stream
 * (it reads back an error flag); kb_mg_apply (one V-cycle z = M^-1 v on reduced fp64 vectors; the cycle itself runs in
 * fp32) is stream-ordered without synchronisation.  kb_mg_transfer_table is host-only arithmetic: fine node i interpolates
 * (1 - w[i]) * coarse[j[i]] + w[i] * coarse[j[i] + 1]; start[c], c = 0..ncoarse: first fine i with j[i] >= c. */
typedef struct kb_mg kb_mg;
int kb_mg_create(int64_t nxn, int64_t nyn, int64_t coarsest_nodes, kb_mg **out);
void kb_mg_destroy(kb_mg *mg);
int kb_mg_num_levels(const kb_mg *mg);
int kb_mg_level_shape(const kb_mg *mg, int level, int64_t *nxn, int64_t *nyn);
size_t kb_mg_workspace_bytes(const kb_mg *mg);
int kb_mg_transfer_table(int64_t nfine, int64_t ncoarse, int32_t *h_j, float *h_w, int32_t *h_start);
int kb_mg_setup(kb_mg *mg, const double *d_points, const double *d_velocity, const int32_t *d_renum,
                const kb_material *material, int mode, int64_t nfree, int64_t nnz, const int32_t *d_indptr,
                const int32_t *d_indices, const double *d_vals, int nu, double omega, void *d_workspace,
                size_t workspace_bytes, void *stream);
int kb_mg_apply(kb_mg *mg, const double *d_v, double *d_z, void *stream);
/* per-sweep damping factors h_omega[0..nu) (nu <= 4) instead of one omega; pass omega <= 0 to kb_mg_setup to keep them */
int kb_mg_set_smoother(kb_mg *mg, int nu, const double *h_omega);
/* cycle schedule: h_reps[l] (1..8) = coarse-grid corrections per visit of level l -- the coarser levels are cycled that many
 * times, each from the corrected iterate; levels >= n keep 1 (all 1 = V-cycle) */
int kb_mg_set_cycle(kb_mg *mg, int n, const int32_t *h_reps);
/* development switches: fused legs of the cycle (one kernel per leg and level instead of one per sweep; bit-identical
 * results) on / off, and the node rows per warp task of the fused kernels */
void kb_mg_set_fused(int on);
void kb_mg_set_assemble_tiled(int on);     /* coarse operators: every element evaluated once per CTA tile (bit-identical; default on) */
void kb_mg_set_tail(int on);               /* the coarsest levels as one shared-memory kernel (bit-identical; default on) */
void kb_mg_set_aggressive(int nodes);      /* plans made afterwards coarsen levels with <= nodes per direction by 4 (default 0: off) */
void kb_mg_set_upwind_restriction(int on);   /* 1: residuals are restricted with the coarse Petrov-Galerkin test functions (default 0; set it before kb_mg_setup, which then computes the weights) */
void kb_mg_set_overweight(double eta);        /* residual over-weighting factor of the restriction (default 1) */
void kb_mg_set_rows_per_task(int rows);
void kb_mg_set_min_rows(int rows);        /* shortest task (rows) the per-level choice may pick */
void kb_mg_set_prefetch_rows(int rows);   /* L2 prefetch distance of the fused kernels (0 = off) */
void kb_mg_set_prefetch_rows_l1(int rows);   /* L1 prefetch distance (0, 1 or 2 rows) */
/* x: initial guess in, solution out; verdict, restarts and kb_solve_info as kb_bicgstab_solve (TRUE residual of the
 * system as handed over).  Synchronises the stream every check_every iterations.  The row partition and row-pattern
 * dictionary of the matrix are kept in d_workspace and reused while the SAME d_indptr / d_indices / d_workspace pointers
 * (and sizes) are passed again: the caller must not change the pattern arrays in place, nor let their storage be recycled
 * for another pattern, between such calls (values may change freely). */
size_t kb_mg_bicgstab_workspace_bytes(int64_t nrows, int64_t nnz);
int kb_mg_bicgstab_solve(kb_mg *mg, int64_t nrows, int64_t nnz, const int32_t *d_indptr, const int32_t *d_indices,
                         const double *d_vals, const double *d_b, double *d_x, double rtol, int maxiter, int check_every,
                         void *d_workspace, size_t workspace_bytes, kb_solve_info *h_info, void *stream);

#ifdef __cplusplus
}
#endif
#endif /* KILTRO_B200_H */
