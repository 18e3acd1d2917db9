/* b200mf.h -- C ABI of the B200-native matrix-free operator engine.
 *
 * This is the drop-in boundary for deal.II's matrix-free hot path.  Every entry point
 * names the reference interface it replaces (paths relative to the deal.II tree,
 * 9.9.0-pre).  Conventions:
 *   - every function returns B200MF_OK (0) or a negative error code;
 *     b200mf_last_error() returns a thread-local message (replaces Assert/AssertThrow,
 *     include/deal.II/base/exceptions.h, which cannot cross a C ABI);
 *   - "device pointer" = plain CUDA device address in the caller's current device;
 *     vectors are caller-owned arrays of `number` laid out as
 *     LinearAlgebra::distributed::Vector stores them: [locally owned | ghosts]
 *     (include/deal.II/lac/la_parallel_vector.h, base/partitioner.h:199);
 *   - `stream` is a cudaStream_t passed as void* (0 = legacy default stream), work is
 *     enqueued asynchronously unless stated otherwise;
 *   - local dof order inside a cell is lexicographic, x fastest
 *     (matrix_free/portable_matrix_free.templates.h:296-298).
 * There is no CPU fallback: without a CUDA device every compute entry point fails with
 * B200MF_ERR_CUDA.
 * Threading: a b200mf_setup owns per-setup scratch (solver work vectors, reduction slots, the
 * staging buffers of the *_host entry points, the ticket/flag arrays of the bulk brick path), so
 * the solver, *_host and vmult entry points of ONE setup must not run concurrently from several
 * host threads or streams (Portable::MatrixFree::cell_loop is not re-entrant on one object either:
 * it re-zeroes the ghost section of src).  Different setups are independent.
 */
#ifndef B200MF_H
#define B200MF_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define B200MF_VERSION 100 /* 0.1.0 */

enum {
  B200MF_OK = 0,
  B200MF_ERR_INVALID = -1, /* bad argument (AssertThrow in the reference)        */
  B200MF_ERR_CUDA = -2,    /* CUDA runtime/driver failure, or no device           */
  B200MF_ERR_UNSUPPORTED = -3,
  B200MF_ERR_NOCONVERGENCE = -4, /* SolverControl::NoConvergence, lac/solver_control.h */
  B200MF_ERR_COMM = -5
};

enum { B200MF_F64 = 0, B200MF_F32 = 1 };

/* How the cell geometry is handed over.  The engine classifies every cell as
 * cartesian / affine / general like internal::MatrixFreeFunctions::MappingInfo
 * (matrix_free/mapping_info.templates.h:428-573) and stores one Jacobian per
 * distinct affine cell and merged metric terms per quadrature point otherwise. */
enum {
  B200MF_GEOMETRY_Q1_VERTICES = 0, /* cell_vertices[n_cells][2^dim][dim], MappingQ1      */
  B200MF_GEOMETRY_JACOBIANS = 1    /* the arrays Portable::MatrixFree stores:
                                      inv_jacobian[n_cells][n_q][dim][dim] and
                                      JxW[n_cells][n_q]
                                      (portable_matrix_free.templates.h:325-338)         */
};

enum { B200MF_CELLS_CARTESIAN = 0, B200MF_CELLS_AFFINE = 1, B200MF_CELLS_GENERAL = 2 };

/* Bit 31 of a local_to_global entry marks a dof that CPU MatrixFree would have dropped
 * from the cell's index list (constrained: read as 0, never written;
 * matrix_free/fe_evaluation.h:3059-3171).  Portable::MatrixFree semantics = bit never set. */
#define B200MF_L2G_CONSTRAINED 0x80000000u

typedef struct b200mf_setup b200mf_setup;
typedef struct b200mf_mesh b200mf_mesh;
typedef struct b200mf_comm b200mf_comm;

/* ------------------------------------------------------------------------------------
 * Setup.  Replaces Portable::MatrixFree<dim,Number>::reinit(mapping, dof_handler,
 * constraints, Quadrature<1>, AdditionalData)  (matrix_free/portable_matrix_free.h:480-527,
 * impl portable_matrix_free.templates.h:1021-1433).  All pointers are HOST pointers, the
 * engine copies what it needs to the device and owns it.
 * ---------------------------------------------------------------------------------- */
typedef struct {
  int dim;           /* 2 or 3                                                          */
  int degree;        /* FE_Q degree p, 1..8                                             */
  int n_q_points_1d; /* QGauss points per direction, in [p+1, 12].  p+1 runs the tuned (collocation)
                        kernels; more points (over-integration) run a plain per-cell kernel (all cell
                        kinds, hanging nodes included)                                                */
  int number;        /* B200MF_F64 / B200MF_F32: arithmetic AND vector element type     */
  uint64_t n_cells;  /* locally owned cells                                             */
  uint64_t n_owned_dofs; /* Partitioner::locally_owned_size()                           */
  uint64_t n_ghost_dofs; /* Partitioner::n_ghost_indices()                              */

  /* [n_cells][(p+1)^dim] process-local indices, lexicographic, hanging-node entries
   * already redirected to the coarse neighbour as HangingNodes::setup_constraints does
   * (matrix_free/hanging_nodes_internal.h:852-880).                                    */
  const uint32_t *local_to_global;
  /* [n_cells] ConstraintKinds bit masks (hanging_nodes_internal.h:40-60) or NULL.       */
  const uint16_t *constraint_mask;

  int geometry; /* B200MF_GEOMETRY_*                                                     */
  const double *cell_vertices; /* Q1_VERTICES                                           */
  const double *inv_jacobian;  /* JACOBIANS                                             */
  const double *JxW;           /* JACOBIANS                                             */

  /* 1D shape data exactly as internal::MatrixFreeFunctions::ShapeInfo holds it
   * (matrix_free/shape_info.templates.h:895-898, 973-984), row-major [i*n_q + q].
   * NULL => the engine builds FE_Q(p) (Gauss-Lobatto nodes) x QGauss(p+1) itself.       */
  const double *shape_values;
  const double *shape_gradients_collocation;
  const double *quadrature_weights;           /* [n_q]                                  */
  const double *subface_interpolation_matrix; /* [n][n] constraint_weights, or NULL     */

  /* local indices (< n_owned_dofs) of constrained dofs, the list
   * copy_constrained_values() walks (portable_matrix_free.templates.h:1366-1425).       */
  const uint32_t *constrained_dofs;
  uint64_t n_constrained_dofs;

  /* Cells [0, n_cells_interior) touch no ghost dof and may run while the ghost exchange
   * is in flight (the colour-0/2 vs colour-1 split of
   * portable_matrix_free.templates.h:1090-1133).  Without ghost dofs the value is ignored; with ghost
   * dofs 0 means "no cell may run before the ghost values have arrived".                          */
  uint64_t n_cells_interior;
} b200mf_setup_desc;

int b200mf_setup_create(const b200mf_setup_desc *desc, b200mf_setup **out);
int b200mf_setup_destroy(b200mf_setup *s);

/* Introspection (PMF::memory_consumption, get_padding_length & friends). */
typedef struct {
  int dim, degree, n_q_points_1d, number;
  uint64_t n_cells, n_owned_dofs, n_ghost_dofs, n_constrained_dofs;
  int cell_kind;                  /* B200MF_CELLS_* of the stored geometry               */
  uint64_t n_distinct_geometries; /* affine compression table size                       */
  uint64_t device_bytes;          /* total device memory owned by the setup              */
  uint64_t geometry_bytes, index_bytes;
  /* cells served by the brick kernel: n_bricks aligned windows of cells_per_brick consecutive
   * cells that form a block of a Morton-ordered Cartesian mesh (index lists compressed to
   * one entry per lattice node, cf. DoFInfo::IndexStorageVariants, matrix_free/dof_info.h) */
  uint64_t n_bricks, cells_per_brick;
} b200mf_setup_info;
int b200mf_setup_get_info(const b200mf_setup *s, b200mf_setup_info *info);

/* Test hook: the device code path of Portable::internal::resolve_hanging_nodes
 * (matrix_free/portable_hanging_nodes_internal.h:414-459) applied to the (degree+1)^dim values of
 * ONE cell (HOST array, overwritten) for a ConstraintKinds mask; transpose != 0 applies the
 * transposed interpolation.  This is what tests/matrix_free/hanging_node_kernels_01.cc exercises in
 * the reference; the parity test feeds its golden vectors through this entry point.            */
int b200mf_debug_resolve_hanging_nodes(int dim, int degree, int number, uint16_t constraint_mask,
                                       int transpose, double *values_host);

/* HOST-only probe of the brick detection the setup runs on the index lists (no device needed):
 * how many aligned windows of cells_per_brick consecutive cells fit together as a block, and how
 * many lattice nodes of those bricks are "complete" (touched by no cell outside their brick, so
 * vmult stores them without atomics).  Geometry is not looked at.                             */
int b200mf_brick_probe(const b200mf_setup_desc *desc, uint64_t *n_bricks, uint64_t *cells_per_brick,
                       uint64_t *n_complete_dofs);

/* HOST-only probe of the bulk brick tables the setup derives from the index lists (no device
 * needed; dealii_b200/csrc/bulk_setup.cpp): the write protocol of vmult on brick meshes is "the first
 * toucher of a dof stores it, later touchers add after the first toucher's flag"; this reports how
 * the lattice nodes of all bricks split into own-range nodes (moved by bulk copies), first-toucher
 * nodes outside an own range (scalar stores), later-toucher nodes (RED), how many distinct relative
 * index patterns describe the bricks, and how many dofs have to be zeroed before a vmult because
 * no brick stores them.  usable = 0 when the setup would fall back to the per-node index maps.  */
typedef struct {
  uint64_t n_bricks, n_patterns, n_own, n_first_scalar, n_later, n_zero, n_general_cells,
      n_boundary_bricks;
  int usable;
  /* b200mf_setup_get_bulk_info only: which path vmult uses, and the per-vmult times the setup
   * measured for both (0 when the choice was forced with B200MF_BULK=0/1)                      */
  int enabled;
  double tuned_ms_index_map, tuned_ms_bulk;
  /* the coloured launches (bricks of one colour share no dof: no atomics, first toucher stores: no
   * memset; bit-reproducible), the third candidate of the setup-time measurement                */
  double tuned_ms_coloured;
  int n_colours, n_coloured_launches;
  uint64_t n_zero_coloured;
  int path; /* 0 = index maps + memset + atomics, 1 = coloured launches, 2 = bulk tables           */
  /* 1: every brick's lattice is an affine image of the numbering (index = base + x + sy y + sz z, e.g.
   * after DoFRenumbering::lexicographic, dofs/dof_renumbering.h:1327-1342): path 0 computes the indices
   * and streams no index map                                                                       */
  int strided;
} b200mf_bulk_info;
int b200mf_bulk_probe(const b200mf_setup_desc *desc, b200mf_bulk_info *info);
/* The setup times both brick paths (index maps + memset + atomics vs bulk tables +
 * first-toucher-stores) on scratch vectors and keeps the faster one ("atomics or a colouring chosen
 * by measurement").  This switch overrides the choice (tests, A/B runs); returns whether bulk
 * tables exist.                                                                                  */
int b200mf_setup_enable_bulk(b200mf_setup *s, int enable);
/* path: 0 = index maps + memset + atomics, 1 = coloured launches, 2 = bulk tables; returns the path
 * now active or -1 if the setup does not have it.                                                */
int b200mf_setup_select_brick_path(b200mf_setup *s, int path);
/* A/B switch for the computed-index (strided) bricks; returns whether the setup has them.            */
int b200mf_setup_enable_strided(b200mf_setup *s, int enable);
int b200mf_setup_get_bulk_info(const b200mf_setup *s, b200mf_bulk_info *info);

/* Quadrature point coordinates, the input of PMF::evaluate_coefficients functors
 * (portable_matrix_free.h:585, get_quadrature_point :417): writes
 * out[(cell*n_q_total + q)*dim + d] to a HOST array. */
int b200mf_get_quadrature_points(const b200mf_setup *s, double *out_host);

/* ------------------------------------------------------------------------------------
 * Operator.  The user functor of Portable::MatrixFree::cell_loop cannot cross a C ABI;
 * the engine implements the functor family of the reference's tests, tutorials and
 * MatrixFreeOperators:   (c_grad grad u, grad v) + (c_mass u, v)
 *   Laplace  : tests/performance/timing_matrix_free_kokkos.cc:129-152,
 *              MatrixFreeOperators::LaplaceOperator (matrix_free/operators.h:892)
 *   Helmholtz: examples/step-64/step-64.cc:120-219
 * Coefficients are device arrays [n_cells * n_q^dim] indexed like
 * PMF::Data::local_q_point_id (portable_matrix_free.h:400-415) or NULL.
 * ---------------------------------------------------------------------------------- */
typedef struct {
  const void *grad_coefficient; /* NULL => 1                                            */
  const void *mass_coefficient; /* NULL => no mass term unless mass_constant != 0       */
  double grad_constant;         /* multiplies the gradient term (1 for Laplace)         */
  double mass_constant;         /* constant mass coefficient (0 => none)                */
} b200mf_operator;

/* Portable::MatrixFree::cell_loop(func, src, dst): dst += sum_cells P^T A_cell P src
 * (portable_matrix_free.h:562; serial_cell_loop portable_matrix_free.templates.h:1440). */
int b200mf_cell_loop(const b200mf_setup *s, const b200mf_operator *op, void *dst,
                     const void *src, void *stream);
/* The same over the local cells [cell_begin, cell_end): the pieces distributed_cell_loop
 * interleaves with the ghost exchange (portable_matrix_free.templates.h:1602-1656).      */
int b200mf_cell_loop_range(const b200mf_setup *s, const b200mf_operator *op, void *dst,
                           const void *src, uint64_t cell_begin, uint64_t cell_end,
                           void *stream);
/* Operator::vmult of the reference's users: dst = 0; cell_loop; copy_constrained_values
 * (examples/step-64/step-64.cc:313-325; MatrixFreeOperators::Base::vmult
 * matrix_free/operators.h:1487 for setups built with B200MF_L2G_CONSTRAINED bits).      */
int b200mf_vmult(const b200mf_setup *s, const b200mf_operator *op, void *dst,
                 const void *src, void *stream);
/* PMF::copy_constrained_values / set_constrained_values
 * (portable_matrix_free.templates.h:711-751, 780-820).                                  */
int b200mf_copy_constrained_values(const b200mf_setup *s, void *dst, const void *src,
                                   void *stream);
int b200mf_set_constrained_values(const b200mf_setup *s, void *dst, double value,
                                  void *stream);
/* MatrixFreeTools::compute_diagonal for Portable::MatrixFree (matrix_free/tools.h:1392-1569):
 * diag_i = A_ii, constrained entries = 1.                                               */
int b200mf_compute_diagonal(const b200mf_setup *s, const b200mf_operator *op, void *diag,
                            void *stream);
/* Same call with HOST vectors (copies inside): the end-to-end path of bench.py.          */
int b200mf_vmult_host(const b200mf_setup *s, const b200mf_operator *op, void *dst_host,
                      const void *src_host);

/* n_vectors independent vmults dst_host[k] = A src_host[k] on HOST vectors (n_owned elements each,
 * page-locked for the copies to overlap), pipelined over two device slots: the upload of vector
 * k+1 and the download of vector k-1 use both directions of the host link while vector k is
 * multiplied.  Returns after the last result has arrived.                                      */
int b200mf_vmult_host_batch(const b200mf_setup *s, const b200mf_operator *op, int n_vectors,
                            void *const *dst_host, const void *const *src_host);

/* ------------------------------------------------------------------------------------
 * Vector kernels: LinearAlgebra::distributed::Vector<Number, MemorySpace::Default>
 * BLAS-1 (lac/vector_operations_internal.h:2140-2660).  n counts elements of `number`.
 * Reductions return through a HOST pointer after synchronising `stream`.
 * ---------------------------------------------------------------------------------- */
int b200mf_vec_set(int number, void *x, double value, uint64_t n, void *stream);
int b200mf_vec_axpy(int number, void *y, double a, const void *x, uint64_t n, void *stream);
int b200mf_vec_sadd(int number, void *y, double s, double a, const void *x, uint64_t n,
                    void *stream); /* y = s*y + a*x */
int b200mf_vec_scale_by(int number, void *y, const void *d, const void *x, uint64_t n,
                        void *stream); /* y = d .* x  (DiagonalMatrix::vmult)            */
int b200mf_vec_dot(int number, const void *x, const void *y, uint64_t n, double *result_host,
                   void *stream);
/* *result_device += x . y without a host round trip (stays on `stream`)                        */
int b200mf_vec_dot_device(int number, const void *x, const void *y, uint64_t n, double *result_device,
                          void *stream);
/* norm_sqr / l2_norm / l1_norm / linfty_norm (vector_operations_internal.h:2480-2560)           */
int b200mf_vec_norm_sqr(int number, const void *x, uint64_t n, double *result_host, void *stream);
int b200mf_vec_norm_2(int number, const void *x, uint64_t n, double *result_host, void *stream);
int b200mf_vec_norm_1(int number, const void *x, uint64_t n, double *result_host, void *stream);
int b200mf_vec_norm_linfty(int number, const void *x, uint64_t n, double *result_host, void *stream);
/* y = a x (+ b v if v != NULL): Vector::equ (Vectorization_equ_au / equ_aubv :2223-2260)         */
int b200mf_vec_equ(int number, void *y, double a, const void *x, double b, const void *v, uint64_t n,
                   void *stream);
/* y = s y + a x + b w: Vector::sadd(s, a, V, b, W) (Vectorization_sadd_xavbw :2312)              */
int b200mf_vec_sadd_xavbw(int number, void *y, double s, double a, const void *x, double b,
                          const void *w, uint64_t n, void *stream);
/* y *= a (d == NULL) or y = y .* d: Vector::operator*= / Vector::scale (:2188, :2338)            */
int b200mf_vec_scale(int number, void *y, double a, const void *d, uint64_t n, void *stream);
/* y += a x; returns y . w: Vector::add_and_dot (AddAndDot :2590), the residual update of SolverCG */
int b200mf_vec_add_and_dot(int number, void *y, double a, const void *x, const void *w, uint64_t n,
                           double *result_host, void *stream);

/* ------------------------------------------------------------------------------------
 * Solver (one process: b200mf_dist_cg_solve / b200mf_mg_dist_cg_solve are the multi-rank entry points).
 * SolverCG<LA::d::Vector>::solve (lac/solver_cg.h:1391) with
 * PreconditionIdentity, DiagonalMatrix (Jacobi, lac/diagonal_matrix.h:435) or
 * PreconditionChebyshev over Jacobi (lac/precondition.h:3928-4121), vector updates and
 * dot products fused around the operator kernel.
 * ---------------------------------------------------------------------------------- */
enum { B200MF_PRECOND_NONE = 0, B200MF_PRECOND_JACOBI = 1, B200MF_PRECOND_CHEBYSHEV = 2 };

typedef struct {
  int preconditioner;
  const void *inverse_diagonal; /* device, [n_owned]; JACOBI / CHEBYSHEV                  */
  /* PreconditionChebyshev::AdditionalData (lac/precondition.h:2121-2175)                */
  int chebyshev_degree;
  double smoothing_range;
  int eig_cg_n_iterations;
  double safety_factor; /* 0 => 1.2                                                      */
  double tolerance;     /* absolute, on ||r||_2 (SolverControl)                          */
  int max_iterations;
  uint64_t first_owned_global_index; /* for the Chebyshev initial guess (global i % 11)   */
  /* Jacobi / identity CG: read the residual norm back (one blocking 24-byte copy) only every
   * check_every iterations; 0 or 1 = every iteration like SolverControl (lac/solver_control.h). With
   * k > 1 the solve may run up to k - 1 iterations past the tolerance.                           */
  int check_every;
  /* PreconditionChebyshev::AdditionalData::max_eigenvalue: used when eig_cg_n_iterations == 0 (no Lanczos
   * estimate): largest eigenvalue = max_eigenvalue (0 => 1), smallest = max_eigenvalue / smoothing_range
   * (lac/precondition.h:2563-2568)                                                                       */
  double max_eigenvalue;
  /* The Lanczos start vector of the eigenvalue estimate: 0 = its entries at this setup's constrained dofs are
   * zeroed, which is what the reference does with AdditionalData::constraints = the problem's constraints
   * (precondition.h:2494-2500, constraints.set_zero); 1 = they are kept, the reference's default (empty
   * AdditionalData::constraints, as in step-37's level smoothers).                                        */
  int eig_keep_constrained_entries;
} b200mf_solver_desc;

typedef struct {
  int iterations;
  double residual;          /* last ||r||_2                                               */
  double initial_residual;
  double chebyshev_max_eigenvalue, chebyshev_min_eigenvalue;
  uint64_t operator_applications;
} b200mf_solver_result;

int b200mf_cg_solve(const b200mf_setup *s, const b200mf_operator *op,
                    const b200mf_solver_desc *solver, void *x, const void *b,
                    b200mf_solver_result *result, void *stream);
/* Building blocks of the same fused CG for callers that interleave their own communication
 * (multi-GPU: all-reduce of the scalar slots between the kernels, see dealii_b200/distributed.py
 * and INTEGRATION.md).  `scratch`: caller-owned, zero-initialised device array of 24 doubles,
 * slot(k) = scratch + 8*(k%3) = [p.Ap, r.r, r.z] of iteration k
 * (the partial sums of lac/solver_cg.h:871-893).
 *   init : r = b - Ax (Ax may be NULL), p = D^-1 r;   slot(1)[1,2] += r.r, r.z
 *   post : alpha = slot(it)[2]/slot(it)[0]; r -= alpha v;  slot(it+1)[1,2] += r.r, r.D^-1 r
 *   pre  : x += alpha p; p = beta p + D^-1 r (beta = slot(it+1)[2]/slot(it)[2]); slot(it+2) = 0
 *   final: x += alpha p
 * d = NULL => PreconditionIdentity.                                                      */
int b200mf_cg_init(int number, void *r, void *p, const void *b, const void *Ax, const void *d,
                   uint64_t n, double *scratch, void *stream);
int b200mf_cg_post(int number, void *r, const void *v, const void *d, uint64_t n, double *scratch,
                   int it, void *stream);
int b200mf_cg_pre(int number, void *x, void *p, const void *r, const void *d, uint64_t n,
                  double *scratch, int it, void *stream);
int b200mf_cg_final(int number, void *x, const void *p, uint64_t n, const double *scratch, int it,
                    void *stream);
/* cell loop over a cell range that also accumulates src.(A src) of these cells into the
 * device double *dot_accum (the p.Ap of CG); copy_constrained_values adding src_c^2.      */
int b200mf_cell_loop_range_dot(const b200mf_setup *s, const b200mf_operator *op, void *dst,
                               const void *src, uint64_t cell_begin, uint64_t cell_end,
                               double *dot_accum, void *stream);
int b200mf_copy_constrained_values_dot(const b200mf_setup *s, void *dst, const void *src,
                                       double *dot_accum, void *stream);
/* One piece of a vmult over the local cells [cell_begin, cell_end): b200mf_cell_loop_range_dot
 * for callers that zeroed dst (b200mf_vmult_prepare) before the first piece and let nothing else write it
 * until the last piece is enqueued (distributed_cell_loop's interior / boundary pieces,
 * portable_matrix_free.templates.h:1602-1656, after the dst = 0 of the operator's vmult).  With
 * that guarantee the engine stores (instead of adds) the dofs only one brick of cells touches.
 * dot_accum may be NULL.                                                                     */
/* The "dst = 0" that precedes those pieces: zeroes at least every entry of dst (owned and ghost)
 * that the pieces accumulate into -- all of dst, or only the dofs on brick surfaces when the
 * brick kernel serves the operator and stores everything else.                                */
int b200mf_vmult_prepare(const b200mf_setup *s, const b200mf_operator *op, void *dst, void *stream);
int b200mf_vmult_range(const b200mf_setup *s, const b200mf_operator *op, void *dst,
                       const void *src, uint64_t cell_begin, uint64_t cell_end,
                       double *dot_accum, void *stream);

/* Ghost exchange kernels: Utilities::MPI::Partitioner::export_to_ghosted_array_start packs
 * buf[i] = vec[import_indices[i]] (base/partitioner.templates.h:119-137);
 * import_from_ghosted_array_finish adds vec[import_indices[i]] += buf[i] (:604-671).
 * The transport between ranks (ncclSend/ncclRecv) is the caller's.                        */
int b200mf_ghost_pack(int number, void *buf, const void *vec, const uint32_t *import_indices,
                      uint64_t n, void *stream);
int b200mf_ghost_unpack_add(int number, void *vec, const void *buf,
                            const uint32_t *import_indices, uint64_t n, void *stream);

/* ------------------------------------------------------------------------------------
 * Multi-GPU: communicator, Utilities::MPI::Partitioner, ghost exchange, distributed cell loop and
 * distributed CG behind the C ABI (one process per GPU; NCCL over NVLink).
 * ---------------------------------------------------------------------------------- */
#define B200MF_UNIQUE_ID_BYTES 128
typedef struct b200mf_partitioner b200mf_partitioner;
/* rank 0 calls b200mf_comm_get_unique_id and hands the 128 bytes to the other ranks by any means
 * (MPI_Bcast in a deal.II host); every rank then calls b200mf_comm_create on ITS current device.  */
int b200mf_comm_get_unique_id(void *id_out);
int b200mf_comm_create(const void *unique_id, int n_ranks, int rank, b200mf_comm **out);
int b200mf_comm_destroy(b200mf_comm *c);
/* Utilities::MPI::sum of `count` device doubles, in place, on `stream`                            */
int b200mf_comm_allreduce_sum(b200mf_comm *c, double *device_values, int count, void *stream);

/* Utilities::MPI::Partitioner(locally_owned, ghost_indices, communicator)
 * (source/base/partitioner.cc:185-330): rank_offsets[r] .. rank_offsets[r+1] is rank r's owned
 * global range, ghost_global the sorted ghost indices of THIS rank (HOST arrays).  The ghost lists
 * are exchanged over the communicator to build import_targets / import_indices.                  */
int b200mf_partitioner_create(b200mf_comm *c, const uint64_t *rank_offsets, const uint64_t *ghost_global,
                              uint64_t n_ghost, int number, b200mf_partitioner **out);
/* The same index algebra without a communicator (and without a device): the caller supplies which
 * ranks ghost which of this rank's dofs (global indices, concatenated in the order of
 * import_ranks).  Objects built this way cannot exchange.                                        */
int b200mf_partitioner_create_host(int n_ranks, int rank, const uint64_t *rank_offsets,
                                   const uint64_t *ghost_global, uint64_t n_ghost, int number,
                                   const int *import_ranks, const uint64_t *import_counts,
                                   int n_import_ranks, const uint64_t *import_global,
                                   b200mf_partitioner **out);
int b200mf_partitioner_destroy(b200mf_partitioner *p);
typedef struct {
  uint64_t n_owned, n_ghost, n_import;
  int n_ghost_targets, n_import_targets;
  const int *ghost_target_ranks;        /* Partitioner::ghost_targets(): (rank, count) ascending   */
  const uint64_t *ghost_target_counts;
  const int *import_target_ranks;       /* Partitioner::import_targets()                           */
  const uint64_t *import_target_counts;
  const uint32_t *import_indices;       /* [n_import] local owned indices, per import target       */
} b200mf_partitioner_info;
int b200mf_partitioner_get_info(const b200mf_partitioner *p, b200mf_partitioner_info *info);
/* LA::d::Vector::update_ghost_values / compress(VectorOperation::add) / zero_out_ghost_values
 * (lac/la_parallel_vector.templates.h:1026-1332) on `stream`; compress also zeroes the ghosts.    */
int b200mf_update_ghost_values(const b200mf_partitioner *p, void *vec, void *stream);
int b200mf_compress_add(const b200mf_partitioner *p, void *vec, void *stream);
int b200mf_zero_out_ghost_values(const b200mf_partitioner *p, void *vec, void *stream);
/* Operator::vmult on distributed vectors = Portable::MatrixFree::distributed_cell_loop
 * (portable_matrix_free.templates.h:1567-1690) inside vmult: dst = 0; update_ghost_values(src)
 * overlapped with the first half of the interior cells; the cells touching ghosts; compress(dst)
 * overlapped with the rest of the interior cells; ghosts of src and dst zeroed;
 * copy_constrained_values.  src is logically const (its ghost section is written and re-zeroed). */
int b200mf_dist_vmult(const b200mf_setup *s, const b200mf_partitioner *p, const b200mf_operator *op,
                      void *dst, void *src, void *stream);
/* b200mf_vmult_host_batch on distributed vectors: every rank streams n_vectors HOST vectors (owned
 * part, page-locked) through the ghost-exchanging vmult, copies of neighbouring vectors overlapped. */
int b200mf_dist_vmult_host_batch(const b200mf_setup *s, const b200mf_partitioner *p,
                                 const b200mf_operator *op, int n_vectors, void *const *dst_host,
                                 const void *const *src_host);
int b200mf_dist_compute_diagonal(const b200mf_setup *s, const b200mf_partitioner *p,
                                 const b200mf_operator *op, void *diag, void *stream);
/* SolverCG (PreconditionIdentity / Jacobi) on distributed vectors: b200mf_cg_solve with the partial
 * sums of every iteration all-reduced (lac/solver_cg.h:871-893, Utilities::MPI::sum).             */
int b200mf_dist_cg_solve(const b200mf_setup *s, const b200mf_partitioner *p, const b200mf_operator *op,
                         const b200mf_solver_desc *solver, void *x, const void *b,
                         b200mf_solver_result *result, void *stream);

/* x and b are HOST arrays of n_owned_dofs elements (copies inside).                      */
int b200mf_cg_solve_host(const b200mf_setup *s, const b200mf_operator *op,
                         const b200mf_solver_desc *solver, void *x_host, const void *b_host,
                         b200mf_solver_result *result);

/* ------------------------------------------------------------------------------------
 * Geometric multigrid: the V-cycle preconditioner of step-37 with the
 * engine's operator on every level.  Replaces, on the device,
 *   Multigrid::level_v_step                       multigrid/multigrid.templates.h:112-171
 *   PreconditionMG::vmult                         multigrid/multigrid.h
 *   MGTransferMatrixFree::prolongate / restrict_and_add   multigrid/mg_transfer_matrix_free.h
 *   mg::SmootherRelaxation<PreconditionChebyshev> (degree 5, range 15, 10 Lanczos iterations) and
 *   MGCoarseGridApplySmoother (Chebyshev in solver mode on level 0)   examples/step-37/step-37.cc:956-988
 * Levels are serial, globally refined meshes of one number type (float levels under a double
 * CG as in step-37 are supported: b200mf_mg_vcycle / b200mf_mg_cg_solve convert at the top).
 * ---------------------------------------------------------------------------------- */
typedef struct b200mf_mg b200mf_mg;

typedef struct {
  int32_t n_levels;                    /* coarsest level first                                         */
  const b200mf_setup *const *levels;   /* n_levels setups; level l+1 is level l refined once           */
  const b200mf_operator *operators;    /* n_levels operators (coefficient arrays belong to their level) */
  /* NULL: the children of cell c of level l are the cells (c << dim) + k of level l+1, k = x + 2 y + 4 z
   * (b200mf_mesh_create in Morton order, deal.II's refine_global); else n_levels - 1 HOST arrays
   * [n_cells(l)][2^dim] with the cell indices of the children on level l+1                            */
  const uint32_t *const *child_cells;
  int32_t smoother_degree;             /* PreconditionChebyshev::AdditionalData::degree (step-37: 5)    */
  double smoothing_range;              /* ... smoothing_range (15)                                     */
  int32_t eig_cg_n_iterations;         /* ... eig_cg_n_iterations (10)                                 */
  double coarse_tolerance;             /* level 0: smoothing_range < 1 = relative tolerance of the
                                          Chebyshev solver, its degree from the error estimate (1e-3)  */
  double safety_factor;                /* on the largest Lanczos eigenvalue; 0 => 1.2                  */
  /* NULL (one rank) or n_levels partitioners of the level number type: level l lives on the partitioned
   * mesh of its setup (vectors = n_owned + n_ghost entries).  Every rank's cells of level l+1 must be the
   * children of its own cells of level l (b200mf_mesh_create_partitioned with the same coarse cells and
   * rank count on every level does that); child_cells are then given in local cell indices, from
   * b200mf_partition_view::cell_morton_position.  All ranks call every b200mf_mg_* function together.   */
  const b200mf_partitioner *const *partitioners;
} b200mf_mg_desc;

typedef struct {
  double eig_min, eig_max;             /* estimates of the Jacobi-preconditioned level operator        */
  int32_t degree;                      /* polynomial degree used on this level                         */
  int32_t eig_cg_iterations;
  uint64_t n_dofs;
  const void *inverse_diagonal;        /* DEVICE, level number type                                    */
} b200mf_mg_level_info;

/* Computes the level diagonals, transfer weights and eigenvalue estimates (the lazy
 * estimate_eigenvalues of the reference's smoothers happens here).  The setups must outlive it. */
int b200mf_mg_create(const b200mf_mg_desc *desc, b200mf_mg **out, void *stream);
/* HOST: the 1D embedding matrix of FE_Q(degree) the transfer kernels apply per direction,
 * P[X * (degree+1) + i] = l_i(x_X) for the 2*degree+1 nodes x_X of the two children in the parent's unit cell
 * (out: (2*degree+1)*(degree+1) doubles).  Needs no device.                                             */
int b200mf_mg_prolongation_matrix_1d(int degree, double *out);
void b200mf_mg_destroy(b200mf_mg *mg);
int b200mf_mg_get_level_info(const b200mf_mg *mg, int level, b200mf_mg_level_info *info);
/* MGTransferMatrixFree::prolongate(to_level, dst, src): dst (level to_level) = P src (level to_level-1);
 * restrict_and_add(from_level, dst, src): dst (level from_level-1) += P^T src.  Level number type. */
int b200mf_mg_prolongate(const b200mf_mg *mg, int to_level, void *dst, const void *src, void *stream);
int b200mf_mg_restrict_and_add(const b200mf_mg *mg, int from_level, void *dst, const void *src,
                               void *stream);
/* PreconditionMG::vmult: dst = one V-cycle applied to src, vectors of type `number` on the finest level. */
int b200mf_mg_vcycle(b200mf_mg *mg, int number, void *dst, const void *src, void *stream);
/* SolverCG on `system` (the finest-level operator, possibly of another number type than the levels)
 * preconditioned by the V-cycle; tolerance is absolute (SolverControl).                               */
int b200mf_mg_cg_solve(b200mf_mg *mg, const b200mf_setup *system, const b200mf_operator *op,
                       double tolerance, int max_iterations, void *x, const void *b,
                       b200mf_solver_result *result, void *stream);
/* the same on a partitioned mesh: the system operator with its partitioner (of the system's number type);
 * dot products are all-reduced over the communicator, the V-cycle exchanges ghosts level by level      */
int b200mf_mg_dist_cg_solve(b200mf_mg *mg, const b200mf_setup *system,
                            const b200mf_partitioner *system_partitioner, const b200mf_operator *op,
                            double tolerance, int max_iterations, void *x, const void *b,
                            b200mf_solver_result *result, void *stream);

/* ------------------------------------------------------------------------------------
 * Synthetic mesh + DoF generator (HOST).  Stands in for GridGenerator::hyper_cube +
 * Triangulation::refine_global / subdivided_hyper_cube + DoFHandler::distribute_dofs
 * (source/dofs/dof_handler_policy.cc:1676-1719) for the >=100 M-DoF runs the reference's
 * host setup cannot hold; numbering is bit-identical to deal.II's (tests/test_mesh.py).
 * ---------------------------------------------------------------------------------- */
enum { B200MF_MESH_MORTON = 0, B200MF_MESH_LEXICOGRAPHIC = 1 };
enum { B200MF_DEFORM_NONE = 0, B200MF_DEFORM_SINE = 1 };

typedef struct {
  int dim, degree;
  int cells_per_direction; /* power of two for MORTON                                    */
  int cell_order;          /* B200MF_MESH_*                                              */
  double left, right;
  int deformation;      /* B200MF_DEFORM_SINE: x += a * prod_d sin(pi (x_d-left)/(right-left)) e */
  double deformation_amplitude;
  int dirichlet_boundary;   /* 1 => collect boundary dofs as constrained                 */
  int mark_constrained_l2g; /* 1 => set B200MF_L2G_CONSTRAINED bits (CPU-MF semantics)   */
  /* 0: DoFHandler::distribute_dofs numbering (first touch); 1: that numbering followed by
   * DoFRenumbering::lexicographic (dofs/dof_renumbering.h:1327-1342; source/dofs/dof_renumbering.cc:
   * 2636-2684: support points sorted by z, then y, then x) -- b200mf_mesh_create only           */
  int dof_numbering;
} b200mf_mesh_desc;

typedef struct {
  uint64_t n_cells, n_dofs, n_boundary_dofs;
  int dofs_per_cell, vertices_per_cell, dim;
  const uint32_t *local_to_global; /* [n_cells][dofs_per_cell], lexicographic           */
  const double *cell_vertices;     /* [n_cells][2^dim][dim]                             */
  const uint32_t *boundary_dofs;   /* sorted                                            */
} b200mf_mesh_view;

int b200mf_mesh_create(const b200mf_mesh_desc *desc, b200mf_mesh **out);

/* The same mesh partitioned like parallel::distributed::Triangulation (p4est): the active
 * cells in Morton order are cut into n_ranks equal contiguous chunks; DoFs on partition
 * interfaces belong to the lowest rank touching them and every rank numbers its owned DoFs
 * by first touch over its own cells, shifted by the DoF counts of the lower ranks
 * (source/dofs/dof_handler_policy.cc:3644-3760).  The domain is coarse[0] x coarse[1] x
 * coarse[2] unit cubes (subdivided_hyper_rectangle, lexicographic) each refined
 * log2(cells_per_direction) times; the chunks must be boxes (n_ranks = number of coarse cells,
 * or one coarse cell and n_ranks a power of two).  Every rank builds ITS part only, with no
 * communication; local indices follow LinearAlgebra::distributed::Vector /
 * Utilities::MPI::Partitioner: [owned | ghosts sorted by global index].                 */
enum { B200MF_GHOSTS_RELEVANT = 0, /* all dofs of ghost cells: Portable::MatrixFree's set
                                      (portable_matrix_free.templates.h:1326-1343)         */
       B200MF_GHOSTS_TOUCHED = 1   /* only dofs the own cells touch: CPU MatrixFree's tight
                                      set (source/matrix_free/dof_info.cc:145-262)         */ };
typedef struct {
  b200mf_mesh_desc mesh; /* cell_order must be B200MF_MESH_MORTON                          */
  int coarse[3];         /* coarse cells per direction (0 => 1)                            */
  int n_ranks, rank;
  int ghost_mode;        /* B200MF_GHOSTS_*                                                */
  int want_lattice_ids;  /* 1 => also return the global lattice id of every local dof (tests) */
} b200mf_partition_desc;

typedef struct {
  uint64_t n_global_dofs, n_global_cells;
  uint64_t first_owned_global; /* start of this rank's contiguous global range            */
  uint64_t n_owned, n_ghost;
  uint64_t n_cells_interior;   /* local cells [0, n_cells_interior) touch no ghost dof      */
  const uint64_t *rank_offsets;  /* [n_ranks + 1] global range starts of all ranks          */
  const uint64_t *ghost_global;  /* [n_ghost] sorted global indices of the ghost dofs        */
  const uint64_t *lattice_ids;   /* [n_owned + n_ghost] or NULL                             */
  /* [n_cells] position of every local cell in this rank's Morton chunk (local cells are ordered
   * [touching no ghost | rest]); the children of the cell at position c are at positions
   * (c << dim) + k of the mesh refined once more -- the multigrid transfer's child tables        */
  const uint64_t *cell_morton_position;
} b200mf_partition_view;

int b200mf_mesh_create_partitioned(const b200mf_partition_desc *desc, b200mf_mesh **out);

/* The partitioned mesh with one level of adaptive refinement and hanging nodes (BASELINE
 * configs[3]): n_ranks = number of coarse cubes, one cube per rank; every cell whose centre is
 * closer than ball_radius (in cube edges) to the centre of its cube is refined once more, as
 * tests/matrix_free/matrix_vector_03.cc refines.  Active cell order, DoF numbering
 * (source/dofs/dof_handler_policy.cc:1676-1719: first touch, level by level), ConstraintKinds masks
 * and the redirected index lists of Portable::MatrixFree (matrix_free/hanging_nodes_internal.h:
 * 40-60, 520-880) are the reference's (bit-identical to deal.II on one rank, tests/test_adaptive_mesh.py);
 * the constrained-dof list of the mesh view holds Dirichlet and hanging-node dofs.  ghost_mode must
 * be B200MF_GHOSTS_TOUCHED.  brick_friendly_order = 1 emits the cells in an order that keeps whole
 * unmasked b^3 blocks together (coarse blocks, refined blocks, other unmasked cells, masked cells,
 * then the cells touching ghost dofs); active_cell_index maps back to the reference's order.   */
typedef struct {
  b200mf_partition_desc part;
  double ball_radius;
  int brick_friendly_order;
} b200mf_adaptive_desc;
typedef struct {
  const uint16_t *constraint_mask;   /* [n_cells]                                            */
  const uint32_t *active_cell_index; /* [n_cells] position in the reference's active cell order */
  uint64_t n_hanging_dofs, n_masked_cells;
  const double *dof_coords;          /* [n_owned + n_ghost][3] support points, if want_lattice_ids */
} b200mf_adaptive_view;
int b200mf_mesh_create_adaptive(const b200mf_adaptive_desc *desc, b200mf_mesh **out);
int b200mf_mesh_adaptive_view_get(const b200mf_mesh *m, b200mf_adaptive_view *view);
int b200mf_mesh_partition_view_get(const b200mf_mesh *m, b200mf_partition_view *view);
int b200mf_mesh_view_get(const b200mf_mesh *m, b200mf_mesh_view *view);
int b200mf_mesh_destroy(b200mf_mesh *m);
/* Convenience: setup straight from a generated mesh. */
int b200mf_setup_create_from_mesh(const b200mf_mesh *m, int number, b200mf_setup **out);

/* ------------------------------------------------------------------------------------ */
const char *b200mf_last_error(void);
int b200mf_version(void);
/* Number of engine kernels launched by this process so far (bench.py "gpu_launches"). */
uint64_t b200mf_kernel_launch_count(void);

#ifdef __cplusplus
}
#endif
#endif /* B200MF_H */
