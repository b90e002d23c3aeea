/*
 * nsb.h -- C ABI of the B200-native Navier-Stokes timestep engine.
 *
 * Drop-in boundary for ONE hot path of lelecaruso/NavierStokes_Project_NM4PDE: the
 * per-timestep cell-loop assembly of the P2-P1 velocity-pressure block system and its
 * block-preconditioned GMRES solve.  The reference has no FFI; its seam is the three
 * protected methods of `NavierStokes` whose only caller is `NavierStokes::solve()`
 * (Navier-Stokes/src/NavierStokes2D.cpp:731-734).  Every entry point below names the
 * reference interface it replaces (paths relative to /root/reference/Navier-Stokes).
 *
 * Conventions: plain pointers and sizes, no C++/torch types; every function returns 0 on
 * success and a negative code on error (`nsb_last_error` gives the message; no exception
 * crosses the boundary); a handle is not thread-safe; one handle per GPU / rank; the caller
 * owns host arrays, the library owns device memory.  Indices are int32, values FP64.
 *
 * Numbering contract (the reference's, single rank): DoFs are in
 * DoFRenumbering::component_wise order (src/NavierStokes2D.cpp:67-69): velocity block first,
 * components interleaved per P2 node (dof = dim*node + c), then pressure (n_u + vertex).
 * `cell_dofs` uses the FESystem local order: per vertex [u_0..u_{dim-1}, p], then per edge
 * [u_0..u_{dim-1}], edges (0,1),(1,2),(2,0),(0,3),(1,3),(2,3).
 *
 * There is NO CPU fallback: every compute entry point fails with NSB_ERR_CUDA when no
 * sm_100 device is usable.
 */
#ifndef NSB_H
#define NSB_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif
#if defined(__GNUC__)
#pragma GCC visibility push(default)
#endif

#define NSB_OK 0
#define NSB_ERR_ARG (-1)
#define NSB_ERR_CUDA (-2)
#define NSB_ERR_STATE (-3)
#define NSB_ERR_NOCONV (-4) /* SolverControl::NoConvergence */
#define NSB_ERR_NCCL (-5)
#define NSB_ERR_IO (-6)

typedef struct nsb_handle_s *nsb_handle;

/* blocks of the 2x2 system (TrilinosWrappers::BlockSparseMatrix, include/NavierStokes2D.hpp:228-237) */
#define NSB_BLK_F 0  /* system_matrix.block(0,0)            n_u x n_u */
#define NSB_BLK_BT 1 /* system_matrix.block(0,1) = -B^T     n_u x n_p */
#define NSB_BLK_B 2  /* system_matrix.block(1,0) = +B       n_p x n_u */
#define NSB_BLK_MP 3 /* pressure_mass.block(1,1)            n_p x n_p */
#define NSB_BLK_S 4  /* B diag(-1/D) B_t built by mmult (include/Preconditioners.hpp:248,358) */
/* which matrix of the class */
#define NSB_MAT_SYSTEM 0
#define NSB_MAT_MASS 1
#define NSB_MAT_STIFFNESS 2
#define NSB_MAT_CONVECTION 3

/* the three copies of `class NavierStokes` */
#define NSB_VARIANT_2D 0   /* src/NavierStokes2D.cpp  */
#define NSB_VARIANT_3D 1   /* src/NavierStokes3D.cpp  (no Temam term after step 1, :456) */
#define NSB_VARIANT_CONV 2 /* src/Convergence3D.cpp   (convection twice in step 1, :277,:284) */

/* `preconditioner_type` of solve_time_step (src/NavierStokes2D.cpp:547-619) */
#define NSB_PREC_YOSIDA 0
#define NSB_PREC_SIMPLE 1
#define NSB_PREC_AYOSIDA 2
#define NSB_PREC_ASIMPLE 3

typedef struct nsb_params {
  double nu;             /* include/NavierStokes2D.hpp:159 (1e-3), Convergence3D.hpp:326 (1e-2) */
  double deltat;         /* ctor argument */
  int32_t variant;       /* NSB_VARIANT_* */
  int32_t precond_type;  /* NSB_PREC_*; reference: 2D 3, 3D 0, CONV 0 */
  int32_t gmres_tmp;     /* SolverGMRES max_n_tmp_vectors, deal.II default 30 */
  int32_t outer_maxit;   /* 100000, src/NavierStokes2D.cpp:534 */
  double outer_tol;      /* 1e-4 ABSOLUTE, src/NavierStokes2D.cpp:535 */
  int32_t inner_maxit;   /* 10000 (aSIMPLE/SIMPLE) or 100000 (Yosida), Preconditioners.hpp:259,368 */
  double inner_rtol;     /* 1e-2, Preconditioners.hpp:260 */
  double alpha_simple;   /* 0.5, Preconditioners.hpp:207 */
  double alpha_asimple;  /* 1.0, Preconditioners.hpp:328 */
  int32_t dirichlet_mode; /* 0: keep nonzero diagonal, rhs = g*a_ii (deal.II Trilinos path); 1: replace by dbar */
  int32_t assembly_kernel; /* 0: tensor-contracted (default), 1: quadrature-loop kernel */
  int32_t sptrsv_kernel;   /* 0: default for this build, 1: level-scheduled launches, 2: chunked persistent */
  int32_t ilu_ordering;    /* 0: natural local row order (Ifpack in the reference; replay mode),
                              1: greedy multicolour ordering of the ILU(0) factors (throughput mode;
                              a different but equally valid ILU(0), like a different mpirun -n P),
                              2: block multicolour (blocks of 32 rows solved sequentially by a warp),
                              3: subdomain ordering (compact parts whose interior rows are solved by one CTA
                                 out of shared memory, separator rows last; csrc/kernels_sd.cu).
                              Must be set before nsb_finalize_setup. */
  int32_t orthogonalisation; /* 0: modified Gram-Schmidt exactly as SolverGMRES (replay mode),
                              1: batched classical Gram-Schmidt (throughput mode): all coefficients of one
                              Arnoldi step from ONE fused multi-dot kernel and one all-reduce; inner solves
                              re-orthogonalise on deal.II's loss test, the outer solve always does two passes */
  int32_t ilu_ordering_schur; /* ordering of the Schur-complement factors: -1 (nsb_default_params): same as
                              ilu_ordering; 0..3 as above.  The pressure matrix has one right-hand side and short
                              colours, so its sweeps are latency-bound: at 19.9 M DoF the point multicolour sweeps
                              (1) cost 0.41 ms per apply against 0.83 ms for the block sweeps (2) at the same CG
                              iteration count, while F_s prefers 2 (fewer inner iterations).  Before nsb_finalize_setup. */
  int32_t reserved[5];
} nsb_params;

/* Fill *p with the reference's literals for the given variant. */
int nsb_default_params(nsb_params *p, int variant);

/* ---- lifetime ------------------------------------------------------------------------- */
/* replaces: NavierStokes ctor/dtor (include/NavierStokes2D.hpp:84-103).  nccl_unique_id is the
 * 128-byte ncclUniqueId made by nsb_get_unique_id on rank 0 (ignored when nranks == 1). */
int nsb_create(nsb_handle *h, int dim, int device_id, int nranks, int rank, const void *nccl_unique_id);
int nsb_destroy(nsb_handle h);
const char *nsb_last_error(nsb_handle h); /* h may be NULL: error of the last failed nsb_create */
int nsb_get_unique_id(void *out128);
int nsb_device_count(void); /* number of usable CUDA devices, 0 if none (never fails) */

/* ---- setup() -------------------------------------------------------------------------- */
/* replaces: dof_handler / mesh / fe for the hot loop (src/NavierStokes2D.cpp:58-93).
 * vertex_coords[n_cells][dim+1][dim], cell_dofs[n_cells][dofs_per_cell] (local indices on
 * this rank: owned DoFs first in each block, then ghosts).  n_u / n_p count local DoFs
 * (owned + ghost); n_u_owned / n_p_owned the owned ones (== n_u / n_p on one rank).
 * Cells are all cells that touch an owned DoF (owned cells + the ghost layer). */
int nsb_set_mesh(nsb_handle h, int32_t n_cells, const double *vertex_coords, const int32_t *cell_dofs,
                 int32_t n_u, int32_t n_p, int32_t n_u_owned, int32_t n_p_owned);
/* replaces: quadrature = QGaussSimplex<dim>(fe->degree+1) (src/NavierStokes2D.cpp:45) */
int nsb_set_quadrature(nsb_handle h, int32_t n_q, const double *xi, const double *w);
int nsb_set_params(nsb_handle h, const nsb_params *p);
/* replaces: TrilinosWrappers::BlockSparsityPattern + make_sparsity_pattern
 * (src/NavierStokes2D.cpp:109-149).  The library derives the same pattern from cell_dofs;
 * this call builds it (and all static analysis: scatter maps, Schur pattern, ILU schedules). */
int nsb_finalize_setup(nsb_handle h);
/* Optional: hand in the reference's own CSR pattern of a block (reference layout, single
 * rank); it is compared with the derived one and NSB_ERR_ARG is returned on any mismatch. */
int nsb_check_pattern(nsb_handle h, int blk, const int32_t *rowptr, const int32_t *colind);
/* Reference-layout CSR pattern of a block (for the parity harness). */
int nsb_get_pattern_size(nsb_handle h, int blk, int32_t *n_rows, int64_t *nnz);
int nsb_get_pattern(nsb_handle h, int blk, int32_t *rowptr, int32_t *colind);
/* Multi-rank ghost exchange plan (replaces Epetra_Import of the ghosted vectors,
 * src/NavierStokes2D.cpp:637,709).  For each neighbour k: send_idx lists local owned
 * P2-node / pressure indices to send, recv counts say how many ghosts arrive (ghosts are
 * stored in arrival order: neighbour by neighbour). */
int nsb_set_halo(nsb_handle h, int32_t n_neighbours, const int32_t *neighbour_ranks,
                 const int32_t *send_node_ptr, const int32_t *send_node_idx, const int32_t *recv_node_cnt,
                 const int32_t *send_p_ptr, const int32_t *send_p_idx, const int32_t *recv_p_cnt);

/* Peer-memory transport over NVLink (replaces MPI point-to-point / MPI_Allreduce of the Epetra layer
 * with direct stores into the peers' HBM, see csrc/halo.cu).  After nsb_set_halo every rank calls
 * nsb_p2p_export (fills a 64-byte cudaIpcMemHandle of its mailbox), the caller all-gathers the
 * handles in rank order (MPI_Allgather / torch.distributed) and hands them to nsb_p2p_attach.
 * Without these two calls, or when peer mapping fails (negative return), the handle keeps using NCCL.
 * nsb_stat(h, "p2p") tells which transport is active. */
int nsb_p2p_export(nsb_handle h, void *handle64);
int nsb_p2p_attach(nsb_handle h, const void *handles /* nranks * 64 bytes */);

/* ---- boundary data -------------------------------------------------------------------- */
/* replaces: VectorTools::interpolate_boundary_values (src/NavierStokes2D.cpp:328-353): rows are
 * velocity DoF indices (local, owned); all components of a node must be listed (the
 * reference only uses full velocity masks).  Values may change every step. */
int nsb_set_dirichlet(nsb_handle h, int32_t n_rows, const int32_t *rows);
int nsb_set_dirichlet_values(nsb_handle h, const double *values);
/* replaces: the Neumann face loop of Convergence3D.cpp:309-330 / 503-527: the caller
 * integrates the face term on the host and hands in the dense contribution to
 * system_rhs.block(0) (length n_u, owned part used); NULL clears it. */
int nsb_set_neumann_rhs(nsb_handle h, const double *rhs_u);

/* ---- drag / lift ---------------------------------------------------------------------- */
/* replaces: the face loop + MPI sum of NavierStokes::compute_forces (src/NavierStokes2D.cpp:752-859:
 * QGauss<1>(3), forces = (nu grad u - p I)(-n) JxW; src/NavierStokes3D.cpp:744-840: QGaussSimplex<2>(3),
 * tangential formula with t = (n_y, -n_x, 0)) on the device, so that solve() no longer copies the whole
 * solution to the host every step.  nsb_set_force_faces (once, after nsb_set_mesh): the obstacle faces
 * (boundary id 3) of LOCALLY OWNED cells -- face k lies in local cell face_cell[k] opposite to its local
 * vertex face_local[k] (the pairs nsh_dofs_boundary_faces returns) -- and the face rule on the unit face
 * (xi[n_q][dim-1], weights summing to 1 in 2D, 1/2 in 3D).  nsb_compute_forces: out[0] = drag,
 * out[1] = lift of the current solution, the raw integrals summed over all ranks (collective: every rank
 * calls it; the c_d / c_l scaling by the mean velocity stays with the caller). */
int nsb_set_force_faces(nsb_handle h, int32_t n_faces, const int32_t *face_cell, const int32_t *face_local,
                        int32_t n_q, const double *xi, const double *w);
int nsb_compute_forces(nsb_handle h, double rho, double *out);

/* ---- state ---------------------------------------------------------------------------- */
/* replaces: VectorTools::interpolate(u_0) -> solution_owned; solution = solution_owned
 * (src/NavierStokes2D.cpp:708-709).  x has n_u + n_p entries [u block | p block] (local). */
int nsb_set_solution(nsb_handle h, const double *x);
int nsb_get_solution(nsb_handle h, double *x); /* multi-rank: ghost entries are refreshed first (collective) */
/* replaces: Utilities::MPI::sum / MPI_Reduce of a few doubles (src/NavierStokes3D.cpp:830-831,
 * Convergence3D.cpp:785-790): vals[n] (n <= 64) summed over all ranks in place, on the engine's transport
 * (collective: every rank calls; a no-op on one rank) */
int nsb_allreduce_sum(nsb_handle h, double *vals, int32_t n);

/* ---- the hot path --------------------------------------------------------------------- */
/* replaces: NavierStokes::assemble(time)            src/NavierStokes2D.cpp:164-357 */
int nsb_assemble_first(nsb_handle h, double time);
/* replaces: NavierStokes::assemble_time_step(time)  src/NavierStokes2D.cpp:361-527 */
int nsb_assemble_step(nsb_handle h, double time);
/* replaces: NavierStokes::solve_time_step           src/NavierStokes2D.cpp:530-639
 * outer_iters = solver_control.last_step(); t_prec / t_solve = time_prec / time_solve. */
int nsb_solve_step(nsb_handle h, int32_t *outer_iters, double *t_prec, double *t_solve);
/* One whole time step with host buffers (the e2e call of bench.py): uploads the Dirichlet
 * values, runs assemble_first/assemble_step + solve_step, downloads the solution. */
int nsb_step_host(nsb_handle h, int first, double time, const double *dirichlet_values, double *solution_out,
                  int32_t *outer_iters);

/* ---- parity harness / operator-level entry points ------------------------------------- */
/* Reference-layout values of a block of one of the class's matrices (pattern of nsb_get_pattern). */
int nsb_get_matrix_values(nsb_handle h, int mat, int blk, double *vals);
int nsb_get_rhs(nsb_handle h, double *rhs);
/* y = system_matrix * x  (BlockSparseMatrix::vmult), host vectors of n_u + n_p */
int nsb_op_system_vmult(nsb_handle h, const double *x, double *y);
/* y = block * x for NSB_BLK_F / BT / B / S */
int nsb_op_block_vmult(nsb_handle h, int blk, const double *x, double *y);
/* Preconditioner::initialize (include/Preconditioners.hpp:223-252, 335-363) */
int nsb_op_precond_init(nsb_handle h);
/* y = ILU(0)^{-1} x for which = 0 (F) or 1 (S)   (TrilinosWrappers::PreconditionILU::vmult) */
int nsb_op_ilu_apply(nsb_handle h, int which, const double *x, double *y);
/* dst = P^{-1} src; dst_in (may be NULL = zeros) is the incoming content of dst, which
 * aSIMPLE uses as initial guess (include/Preconditioners.hpp:273) */
int nsb_op_precond_vmult(nsb_handle h, const double *src, const double *dst_in, double *dst);
/* Row ordering of the ILU(0) factors (which = 0: F_s over P2 nodes, 1: S over pressure DoFs):
 * factor row k is matrix row order[k]; identity for ilu_ordering = 0. */
int nsb_get_ilu_order(nsb_handle h, int which, int32_t *order);
/* Schur complement values on the pattern of NSB_BLK_S */
int nsb_get_schur_values(nsb_handle h, double *vals);

/* ---- measurement ---------------------------------------------------------------------- */
/* Statistics of the last solve_step (inner iteration counts etc.); unknown names give -1. */
double nsb_stat(nsb_handle h, const char *name);
/* Times `iters` launches of one kernel with CUDA events on the launching stream; returns the
 * average ms per launch and the algorithmic bytes one launch moves ("kernels").
 * which: "spmv_system", "spmv_F", "spmv_S", "assemble_step", "ilu_F", "ilu_S", "dot", "axpy",
 * "add_and_dot" */
int nsb_bench_kernel(nsb_handle h, const char *which, int iters, int flush_l2, double *ms_per_launch,
                     double *bytes_per_launch);
/* CUDA-event stopwatch on the handle's launching stream: start (which = 0) / stop (which = 1);
 * nsb_timer_elapsed_ms synchronises on the stop event. */
int nsb_timer_mark(nsb_handle h, int which);
int nsb_timer_elapsed_ms(nsb_handle h, double *ms);
/* number of this library's kernel launches since the last call with reset != 0 */
int64_t nsb_launch_count(nsb_handle h, int reset);

/* CPU-only self check of the subdomain ILU ordering (ilu_ordering = 3) and its packed storage: builds the
 * multi-level ordering of the given (structurally symmetric, diagonal included) graph with parts of
 * <= leaf_levels[l] rows on level l (three entries, 0 ends the list; no further level below min_active
 * active rows), packs the factors as nsb_finalize_setup does, runs the triangular solves through a host emulation
 * of the device kernels with synthetic values and bs right-hand sides, and returns the largest difference
 * to plain substitution relative to max |y| (> 1e29: a structural invariant is violated).  stats[8]:
 * parts, rows in parts, final separator colours, largest (rows + ring), most colours of a part, slot
 * efficiency in per mille, levels, rows of level 1.  order_out[n] (may be NULL): factor row -> row.
 * Test infrastructure. */
int nsb_debug_sd_check(int32_t n, const int32_t *rowptr, const int32_t *colind, const double *xyz, int32_t gdim,
                       const int32_t *leaf_levels, int32_t min_active, int32_t bs, double *rel_err, int32_t *stats,
                       int32_t *order_out);

/* CPU-only self check of the block multicolour ILU storage (ilu_ordering = 2): the ordering of an n x n pattern (blocks
 * of one colour independent), then both factors packed as for the device and walked by a host emulation of the sweep
 * kernel (four passes of eight rows over the entries leaving a block, staged-list indices below xcap, sequential
 * elimination inside the block) with a pseudo-random factor and bs right-hand sides; *rel_err = largest difference to
 * plain forward / backward substitution (>= 1e30: structural violation).  stats[4]: blocks, colours, max outside rows
 * and max in-block entries per block.  order_out[n] (may be NULL): factor row -> row.  Test infrastructure. */
int nsb_debug_bsell_check(int32_t n, const int32_t *rowptr, const int32_t *colind, int32_t bs, int32_t xcap, double *rel_err,
                          int32_t *stats, int32_t *order_out);

/* CPU-only self check of the point multicolour ILU storage (ilu_ordering = 1) and of the SELL-32 SpMV format: greedy
 * colouring of an n x n pattern (colours are independent sets), the strictly lower / upper factor parts in SELL-32 with
 * the colours as row ranges (`lanes` = 1 | 4 lanes per row, rows length-sorted inside windows of `window` rows), and the
 * whole pattern in SELL-32; a host emulation of the sweep / SpMV kernel (one warp per slice, partial sums of the lanes of
 * a row combined, slices of a colour in reverse order) with a pseudo-random factor and bs right-hand sides; *rel_err =
 * largest difference to plain substitution and to a CSR product (>= 1e30: structural violation).  stats[3]: colours,
 * slices of L, padded slots per stored entry of L in per mille.  Test infrastructure. */
int nsb_debug_sell_check(int32_t n, const int32_t *rowptr, const int32_t *colind, int32_t bs, int32_t lanes, int32_t window,
                         double *rel_err, int32_t *stats, int32_t *order_out);

/* CPU-only fingerprint of the HOST side of nsb_set_mesh + nsb_finalize_setup (sparsity patterns, scatter map, SpMV
 * formats, ILU orderings and their packed storage): the same code runs on a handle without device state and every
 * array setup would upload is hashed (FNV-1a, upload order) into hashes[cap]; *n_hashes = number of uploads.  Pins the
 * device data structures on a machine without a GPU (tests/test_setup_fingerprint.py).  Test infrastructure: it
 * computes nothing else, and the product entry points still fail without a device. */
int nsb_debug_setup_fingerprint(int32_t dim, int32_t n_cells, const double *vertex_coords, const int32_t *cell_dofs,
                                int32_t n_u, int32_t n_p, int32_t n_u_owned, int32_t n_p_owned, int32_t ilu_ordering,
                                int32_t ilu_ordering_schur, uint64_t *hashes, int32_t cap, int32_t *n_hashes);

/* ---- host prerequisites (cold path; replaces deal.II GridIn / DoFHandler in setup()) ---- */
typedef struct nsh_mesh_s *nsh_mesh;
typedef struct nsh_dofs_s *nsh_dofs;

/* Generators with the reference's geometry and boundary ids (mesh/Cylinder2D.geo:40-43,
 * mesh/Cylinder3D.geo:126-129, mesh/mesh-cube.geo:16-21).  `s` scales the resolution. */
nsh_mesh nsh_mesh_cylinder2d(int s);
nsh_mesh nsh_mesh_cylinder3d(int s, int nz);
nsh_mesh nsh_mesh_cube(int n);
nsh_mesh nsh_mesh_box(int dim, int nx, int ny, int nz, const double *lo, const double *hi);
/* replaces: GridIn::read_msh (src/NavierStokes2D.cpp:10-14); Gmsh v2 / v4.1 ASCII */
nsh_mesh nsh_mesh_read_msh(const char *path);
int nsh_mesh_write_msh(nsh_mesh m, const char *path);
void nsh_mesh_free(nsh_mesh m);
int nsh_mesh_dim(nsh_mesh m);
int32_t nsh_mesh_n_vertices(nsh_mesh m);
int32_t nsh_mesh_n_cells(nsh_mesh m);
int32_t nsh_mesh_n_bfaces(nsh_mesh m);
const double *nsh_mesh_vertices(nsh_mesh m); /* [n_vertices][dim] */
const int32_t *nsh_mesh_cells(nsh_mesh m);   /* [n_cells][dim+1] */
const int32_t *nsh_mesh_bfaces(nsh_mesh m);  /* [n_bfaces][dim] vertex ids */
const int32_t *nsh_mesh_bface_ids(nsh_mesh m);
const int32_t *nsh_mesh_bface_cells(nsh_mesh m); /* [n_bfaces] owning cell */
/* reorder cells (and thereby the DoF numbering): 0 = generator order, 1 = blocked/coloured */
int nsh_mesh_reorder_cells(nsh_mesh m, int mode, int block);

/* replaces: dof_handler.distribute_dofs + DoFRenumbering::component_wise (src/NavierStokes2D.cpp:62-69) */
nsh_dofs nsh_dofs_create(nsh_mesh m);
void nsh_dofs_free(nsh_dofs d);
int32_t nsh_dofs_n_nodes(nsh_dofs d); /* P2 nodes: n_u = dim * n_nodes */
int32_t nsh_dofs_n_p(nsh_dofs d);
int32_t nsh_dofs_per_cell(nsh_dofs d);
const int32_t *nsh_dofs_cell_dofs(nsh_dofs d);   /* [n_cells][dofs_per_cell], reference layout */
const double *nsh_dofs_node_xyz(nsh_dofs d);     /* [n_nodes][dim] support points of velocity DoFs */
const double *nsh_dofs_p_xyz(nsh_dofs d);        /* [n_p][dim] */
const double *nsh_dofs_cell_coords(nsh_dofs d);  /* [n_cells][dim+1][dim] */
/* P2 nodes on boundary faces whose id is in ids[], in first-visit order; returns the count
 * (out may be NULL to query it). */
int32_t nsh_dofs_boundary_nodes(nsh_dofs d, nsh_mesh m, const int32_t *ids, int32_t n_ids, int32_t *out);
/* Face-quadrature data on boundary faces with the given id (for the Neumann term and the
 * drag/lift integrals): returns n_faces; arrays may be NULL to query. */
int32_t nsh_dofs_boundary_faces(nsh_dofs d, nsh_mesh m, int32_t id, int32_t *face_cell, int32_t *face_local);
/* replaces: the face loop of NavierStokes::compute_forces over the faces with `boundary_id`
 * (src/NavierStokes2D.cpp:752-859: QGauss<1>(3), (nu grad u - p I) n; src/NavierStokes3D.cpp:744-840:
 * QGaussSimplex<2>(3), tangential formula); out[0] = drag, out[1] = lift (unscaled integrals) */
int nsh_boundary_forces(nsh_mesh m, nsh_dofs d, const double *solution, int32_t boundary_id, double nu, double rho,
                        double *out);
/* replaces: VectorTools::point_value (src/NavierStokes2D.cpp:875-889): velocity components and
 * pressure of the solution vector [u | p] at point x, out[dim + 1]; NSB_ERR_ARG when no cell holds x */
int nsh_dofs_point_value(nsh_dofs d, const double *solution, const double *x, double *out);
/* the cell VectorTools::point_value evaluates in: first cell holding x (barycentric coordinates lam[dim + 1]
 * filled), -1 when none does; lets a rank of a multi-rank run evaluate only the points of the cells it owns */
int32_t nsh_dofs_find_cell(nsh_dofs d, const double *x, double *lam);
/* minimal stand-in for DataOut::write_vtu (src/NavierStokes2D.cpp:642-675): ASCII .vtu, linear cells,
 * point data "velocity" and "pressure" of the solution vector [u | p] */
int nsh_write_vtu(nsh_mesh m, nsh_dofs d, const double *solution, const char *path);
/* Partition cells into nparts (recursive coordinate bisection); part[n_cells]. */
int nsh_partition_cells(nsh_mesh m, int nparts, int32_t *part);

/* replaces: GridTools::partition_triangulation + parallel::fullydistributed::Triangulation + the Epetra maps
 * (src/NavierStokes2D.cpp:16-19, 71-87) for `rank` of `nranks`: the local (owned + two-layer ghost) cells in
 * local numbering (owned DoFs first, ghosts grouped by owner rank and sorted by global id) and the ghost
 * exchange plan of nsb_set_halo -- worked out from the replicated mesh without any communication.  The arrays
 * are what nsb_set_mesh / nsb_set_halo take; *_gid map local -> global, g2l_* global -> local (-1: not local). */
typedef struct nsh_local_s *nsh_local;
nsh_local nsh_local_create(nsh_mesh m, nsh_dofs d, int32_t nranks, int32_t rank);
void nsh_local_free(nsh_local l);
int32_t nsh_local_n_cells(nsh_local l);
int32_t nsh_local_n_nodes(nsh_local l);       /* local P2 nodes, owned + ghost */
int32_t nsh_local_n_p(nsh_local l);
int32_t nsh_local_n_nodes_owned(nsh_local l);
int32_t nsh_local_n_p_owned(nsh_local l);
const int32_t *nsh_local_cells(nsh_local l);      /* [n_cells] global cell ids, ascending */
const int32_t *nsh_local_cell_part(nsh_local l);  /* [global cells] owner rank of every cell */
const int32_t *nsh_local_cell_dofs(nsh_local l);  /* [n_cells][dofs_per_cell], local numbering */
const double *nsh_local_cell_coords(nsh_local l); /* [n_cells][dim+1][dim] */
const int32_t *nsh_local_node_gid(nsh_local l);
const int32_t *nsh_local_p_gid(nsh_local l);
const int32_t *nsh_local_g2l_node(nsh_local l);   /* [global nodes] */
const int32_t *nsh_local_g2l_cell(nsh_local l);   /* [global cells] */
/* the seven arrays of nsb_set_halo; returns the number of neighbours */
int32_t nsh_local_halo(nsh_local l, const int32_t **nb_ranks, const int32_t **send_node_ptr, const int32_t **send_node_idx,
                       const int32_t **recv_node_cnt, const int32_t **send_p_ptr, const int32_t **send_p_idx,
                       const int32_t **recv_p_cnt);

#if defined(__GNUC__)
#pragma GCC visibility pop
#endif
#ifdef __cplusplus
}
#endif
#endif /* NSB_H */
