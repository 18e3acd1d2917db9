// Internal declarations shared by the translation units of libb200mf.so.
#pragma once
#include <cuda_runtime.h>

#include <atomic>
#include <cstdarg>
#include <cstdint>
#include <cstdio>
#include <string>
#include <vector>

#include "../../include/b200mf.h"

namespace b200mf {

void set_error(const char *fmt, ...);
extern std::atomic<uint64_t> g_launch_count;
inline void count_launch(uint64_t n = 1) { g_launch_count.fetch_add(n, std::memory_order_relaxed); }

#define B200MF_CUDA_CHECK(expr)                                                          \
  do {                                                                                   \
    cudaError_t err__ = (expr);                                                          \
    if (err__ != cudaSuccess) {                                                          \
      ::b200mf::set_error("%s:%d: %s failed: %s", __FILE__, __LINE__, #expr,             \
                          cudaGetErrorString(err__));                                    \
      return B200MF_ERR_CUDA;                                                            \
    }                                                                                    \
  } while (0)

#define B200MF_REQUIRE(cond, ...)                                                        \
  do {                                                                                   \
    if (!(cond)) {                                                                       \
      ::b200mf::set_error(__VA_ARGS__);                                                  \
      return B200MF_ERR_INVALID;                                                         \
    }                                                                                    \
  } while (0)

constexpr int kMaxN = 9; // degree <= 8
// flag bits of a brick map entry (bit 31 = B200MF_L2G_CONSTRAINED)
constexpr uint32_t B200MF_BRICK_COMPLETE = 0x40000000u; // no cell outside the brick touches the dof
constexpr uint32_t B200MF_BRICK_FIRST = 0x20000000u;    // coloured launches: this brick stores the dof
constexpr uint32_t B200MF_BRICK_INDEX = 0x1fffffffu;
constexpr int kL2gPadCells = 16; // >= cells per warp of every plane-kernel configuration
// cells per direction of a brick: b^3 consecutive cells of a Morton-ordered mesh form a block;
// chosen by measurement (Q1: b = 16 (L = 17) beats 8 by 20 % in FP64; Q3/Q4: b = 4 beats 2 by 20 %; Q5: b = 2 (L = 11, many small CTAs) beats
// b = 4 (L = 21, one 148 KB CTA per SM) by 15 %)
constexpr int brick_edge(int degree) { return degree == 1 ? 16 : degree == 2 ? 8 : degree <= 4 ? 4 : 2; }

// Even-odd packed 1D matrix for out[q] = sum_i M[i][q] in[i] with
// M[n-1-i][n-1-q] = +/- M[i][q]  (cf. shape_info.templates.h:1153-1180 convert_to_eo).
//   E[i*hq + q] = (M[i][q] + M[i][n-1-q]) / 2,  i < h, q < hq
//   O[i*h  + q] = (M[i][q] - M[i][n-1-q]) / 2,  i < h, q < h
//   mid[q]      = M[(n-1)/2][q],                q < hq   (n odd only)
// with h = n/2, hq = (n+1)/2.
template <typename Number, int n>
struct EoMatrix {
  static constexpr int h = n / 2, hq = (n + 1) / 2;
  Number E[h * hq > 0 ? h * hq : 1];
  Number O[h * h > 0 ? h * h : 1];
  Number mid[hq];
};

template <typename Number, int n>
struct ShapeData {
  EoMatrix<Number, n> S;  // dofs -> quadrature points (symmetric)
  EoMatrix<Number, n> St; // quadrature points -> dofs
  EoMatrix<Number, n> D;  // collocation derivative (skew-symmetric)
  EoMatrix<Number, n> Dt;
  EoMatrix<Number, n> DtW; // Dt with the quadrature weight of the input point folded in
  Number w[n];            // 1D quadrature weights
  Number w2[n * n];       // w[a] * w[b] at [a * n + b]
};

// 1D matrices of the brick kernel (brick_kernel.cuh): reference mass matrix and the three
// metric-scaled stiffness matrices, packed with sym = +1
template <typename Number, int n>
struct BrickMatrices {
  EoMatrix<Number, n> M, Kx, Ky, Kz;
};

// Layout of the merged metric of general cells: rows of n quadrature points along x hold
// their NS = dim(dim+1)/2 tensor entries next to each other,
//   3D: [cell][z][y][s][x]      2D: [cell][y][s][x]
// so that a thread reads the 6n (3n) numbers of one row with wide, fully used loads.
template <int dim>
__host__ __device__ inline unsigned long long metric_offset(int n, unsigned long long cell, int s,
                                                            int q) {
  constexpr int NS = dim * (dim + 1) / 2;
  const int x = q % n, rest = q / n; // rest = y (2D) or y + n z (3D)
  return ((cell * (unsigned long long)(dim == 2 ? n : n * n) + rest) * NS + s) * n + x;
}

// Host-side description of the operator passed to the kernels.
template <typename Number>
struct OperatorArgs {
  const Number *grad_coef;
  const Number *mass_coef;
  Number grad_const;
  Number mass_const;
  int has_mass;
};

struct BulkStatsData {
  uint64_t n_bricks = 0, n_patterns = 0, n_own = 0, n_first_scalar = 0, n_later = 0, n_zero = 0,
           n_general_cells = 0, n_boundary_bricks = 0;
};

struct Setup {
  int dim = 0, degree = 0, n = 0, number = 0;
  // quadrature points per direction; > n = over-integration: every cell runs overint.cu, the collocation
  // kernels (bricks, plane, generic) and their tables are not used
  int n_q_1d = 0;
  void *d_overint_tables = nullptr; // S[n][Q] | Dq[Q][Q] | w[Q]  (Number)
  uint64_t n_cells = 0, n_owned = 0, n_ghost = 0, n_constrained = 0, n_cells_interior = 0;
  int dofs_per_cell = 0;
  int cell_kind = B200MF_CELLS_GENERAL;
  uint64_t n_geom = 0;

  // device arrays
  uint32_t *d_l2g = nullptr;
  uint16_t *d_mask = nullptr;
  uint32_t *d_geom_id = nullptr;   // per cell index into d_geom_table (null if table has 1 entry)
  void *d_geom_table = nullptr;    // cartesian: [n_geom][4]; affine: [n_geom][7]  (Number)
  void *d_metric = nullptr;        // general: [n_cells][6 or 3][n_q_total] merged w*det*J^-1 J^-T
  void *d_jxw = nullptr;           // general: [n_cells][n_q_total]
  uint32_t *d_constrained = nullptr;
  void *d_weights = nullptr;       // subface interpolation matrix [n*n] (Number)
  void *d_diag_tables = nullptr;   // [SS | GG | SG] of the sum-factorised diagonal, each [n*n] (Number)
  double *d_qpoints = nullptr;     // optional cache
  // host copies needed later
  std::vector<double> shape_values, shape_grad_colloc, q_weights, q_points_1d, subface;
  std::vector<double> h_vertices;  // kept only when small (quadrature point queries)
  bool has_vertices = false;
  bool any_mask = false;
  uint64_t device_bytes = 0, geometry_bytes = 0, index_bytes = 0;

  // bricks (brick_kernel.cuh / brick_setup.cpp): one index per lattice node of every aligned
  // window of brick_b^3 consecutive cells that forms a block; runs of consecutive bricks
  struct BrickRun { uint64_t cell_begin, cell_end, first_brick; uint32_t geom; };
  uint32_t *d_brick_map = nullptr;
  uint64_t n_bricks = 0;
  // every brick's lattice is an affine image of the numbering (index = base + x + sy y + sz z):
  // {base, sy, sz, face flags} per brick, the kernels compute indices instead of reading the map
  uint4 *d_brick_strided = nullptr;
  bool strided_enabled = true;
  // dofs vmult has to zero before its cell loop (all that no brick stores); valid if have_zero_list
  uint32_t *d_zero_list = nullptr;
  uint64_t n_zero_list = 0;
  bool have_zero_list = false;
  int brick_b = 0;
  std::vector<BrickRun> brick_runs;
  double geom0[4] = {0, 0, 0, 0}; // cartesian metric diagonal + det of the single-geometry mesh
  std::vector<double> h_geom_table;   // cartesian: [n_geom][4] metric diagonal + det (host copy for the bricks)
  std::vector<uint32_t> h_geom_id;    // per cell, kept while n_geom > 1 (brick detection)
  // cells that may carry a hanging-node mask all lie in [masked_begin, masked_end): the kernels
  // without mask handling (plane kernel, sum-factorised diagonal) serve the cells outside
  uint64_t masked_begin = 0, masked_end = 0;

  // coloured brick launches (brick_setup.cpp: build_colouring): no atomics, no memset of dst
  struct Colouring {
    bool ready = false, enabled = true;
    int n_colours = 0;
    uint64_t half = 0; // first half of the interior cells (piece 0 of the distributed schedule)
    struct Launch { int piece; uint32_t geom; uint64_t offset, count; };
    struct CellRange { int piece; uint64_t begin, end; };
    std::vector<Launch> launches;   // in launch order: piece, cell shape, colour
    std::vector<CellRange> general; // cells no brick covers, per piece
    uint32_t *d_list = nullptr, *d_zero = nullptr;
    uint64_t n_zero = 0;
    float tuned_ms = 0.f;
  };
  mutable Colouring colouring;

  // bulk brick path (bulk_kernel.cuh / bulk_setup.cpp): pattern tables + per-brick descriptors in
  // execution order, first-toucher-stores write protocol
  struct Bulk {
    bool ready = false, enabled = true;
    float tuned_ms[2] = {0.f, 0.f}; // setup-time measurement: index-map path, bulk path
    BulkStatsData stats;
    uint32_t n_exec = 0, n_patterns = 0;
    int L = 0, TP = 0;
    uint64_t exec_boundary_begin = 0, exec_boundary_end = 0; // bricks touching ghost dofs
    uint32_t *d_desc = nullptr, *d_other = nullptr, *d_phdr = nullptr;
    uint16_t *d_own_pos = nullptr, *d_other_pos = nullptr;
    uint32_t *d_flags = nullptr, *d_ticket = nullptr, *d_zero = nullptr;
    uint64_t n_zero = 0;
    std::vector<std::pair<uint64_t, uint64_t>> general_ranges; // cells no brick covers
    uint32_t epoch = 0; // launch counter: flags hold the epoch of the launch that set them
    // distributed vmult (comm.cu): boundary bricks wait for *sync_ghost_ready == epoch and count
    // themselves into *sync_boundary_done; null = no synchronisation with a ghost exchange
    uint32_t *sync_ghost_ready = nullptr, *sync_boundary_done = nullptr;
  };
  mutable Bulk bulk;

  // scratch for reductions / solver
  double *d_scratch = nullptr;
  double *h_pinned = nullptr;
  void *d_work[6] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
  uint64_t work_elems = 0;
  // staging buffers for the *_host entry points
  void *d_stage[2] = {nullptr, nullptr};
  // pipeline of b200mf_vmult_host_batch: two input and two output staging vectors, three streams
  void *d_pipe_in[2] = {nullptr, nullptr}, *d_pipe_out[2] = {nullptr, nullptr};
  cudaStream_t pipe_stream[3] = {nullptr, nullptr, nullptr}; // h2d, compute, d2h
  cudaEvent_t pipe_event[3][2] = {{nullptr, nullptr}, {nullptr, nullptr}, {nullptr, nullptr}};
};

size_t number_size(int number);

// dst = 0; cell loop; copy_constrained_values -- optionally accumulating src.dst into
// *dot_accum (device double) on the way
int vmult_impl(const Setup &s, const b200mf_operator &op, void *dst, const void *src,
               cudaStream_t stream, double *dot_accum);
int copy_constrained_impl(const Setup &s, void *dst, const void *src, cudaStream_t st,
                          double *dot_accum);
// the "dst = 0" of vmult for a cell loop that runs in dst-was-zeroed mode
int vmult_prepare_impl(const Setup &s, const b200mf_operator &op, void *dst, cudaStream_t st);
int set_constrained_impl(const Setup &s, void *dst, double value, cudaStream_t st);

// comm.cu: the operator / reductions of a level that may be partitioned over ranks (p == nullptr: one rank)
} // namespace b200mf
struct b200mf_partitioner;
namespace b200mf {
int level_vmult(const Setup &s, const b200mf_partitioner *p, const b200mf_operator &op, void *dst, void *src,
                cudaStream_t st, double *dot_accum);
int level_allreduce(const b200mf_partitioner *p, double *device_values, int count, cudaStream_t st);
uint64_t level_first_owned(const b200mf_partitioner *p);
bool level_is_distributed(const b200mf_partitioner *p);

// overint.cu
int launch_overint(const Setup &s, const b200mf_operator &op, void *dst, const void *src, uint64_t cell_begin,
                   uint64_t cell_end, cudaStream_t st, double *dot_accum, bool diagonal);

// kernels_dispatch.cu
int launch_cell_loop(const Setup &s, const b200mf_operator &op, void *dst, const void *src,
                     uint64_t cell_begin, uint64_t cell_end, cudaStream_t stream,
                     double *dot_accum, bool dst_zeroed = false);
int launch_compute_diagonal(const Setup &s, const b200mf_operator &op, void *diag,
                            cudaStream_t stream);

// bulk_setup.cpp
using BulkStats = BulkStatsData;
int build_bulk(const b200mf_setup_desc &d, Setup &s, const std::vector<uint32_t> &maps, uint64_t nb,
               bool upload, BulkStats *stats);
void free_bulk(Setup &s);
int launch_bulk(const Setup &s, const b200mf_operator &op, void *dst, const void *src,
                cudaStream_t stream, double *dot_accum);

// brick_setup.cpp
int build_bricks(const b200mf_setup_desc &d, Setup &s, bool upload = true,
                 uint64_t *n_complete_out = nullptr, BulkStats *bulk_stats = nullptr);
// kernels: the bricks [brick_begin, brick_begin + n_bricks) of the setup
int launch_bricks(const Setup &s, const b200mf_operator &op, void *dst, const void *src,
                  uint64_t brick_begin, uint64_t n_bricks, cudaStream_t stream, double *dot_accum,
                  bool overwrite, uint32_t geom = 0, const uint32_t *list = nullptr);
// the coloured launches and the per-cell kernels of one piece of the schedule (-1: all pieces)
int launch_coloured(const Setup &s, const b200mf_operator &op, void *dst, const void *src, int piece,
                    cudaStream_t stream, double *dot_accum);
int coloured_prepare(const Setup &s, void *dst, cudaStream_t stream);
bool coloured_enabled(const Setup &s, const b200mf_operator &op);

// shape.cpp
template <typename Number, int n>
void fill_brick_matrices(const Setup &s, const b200mf_operator &op, BrickMatrices<Number, n> &out,
                         uint32_t geom = 0);
void build_fe_q_shape_data(int degree, std::vector<double> &shape_values,
                           std::vector<double> &shape_grad_colloc, std::vector<double> &q_weights,
                           std::vector<double> &q_points, std::vector<double> &subface);

void build_prolongation_1d(int degree, std::vector<double> &P);
// FE_Q(degree) with QGauss(Q), Q > degree + 1: S[i * Q + q] = l_i(x_q) (Gauss-Lobatto Lagrange basis at the
// Gauss points), Dq[a * Q + b] = derivative of the Lagrange polynomial of Gauss point a at Gauss point b
void build_overint_shape_data(int degree, int Q, std::vector<double> &S, std::vector<double> &Dq,
                              std::vector<double> &weights, std::vector<double> &points);

template <typename Number, int n>
void fill_shape_data(const Setup &s, ShapeData<Number, n> &out);

} // namespace b200mf

struct b200mf_setup {
  b200mf::Setup impl;
};
