// C ABI entry points: setup (Portable::MatrixFree::reinit), operator application
// (cell_loop / vmult / copy_constrained_values / compute_diagonal).
#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <map>
#include <unordered_map>

#include "internal.h"
#include "vector_ops.cuh"

namespace b200mf {

static thread_local std::string g_error;
std::atomic<uint64_t> g_launch_count{0};

void set_error(const char *fmt, ...) {
  char buf[1024];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof(buf), fmt, ap);
  va_end(ap);
  g_error = buf;
}

size_t number_size(int number) { return number == B200MF_F64 ? 8 : 4; }

#define DECL_N(N)                                                                             \
  int launch_cells_n##N(const Setup &, const b200mf_operator &, void *, const void *, uint64_t, \
                        uint64_t, cudaStream_t, bool, double *, bool);
DECL_N(2) DECL_N(3) DECL_N(4) DECL_N(5) DECL_N(6) DECL_N(7) DECL_N(8) DECL_N(9)
#undef DECL_N
#define DECL_N(N)                                                                              \
  int launch_bricks_n##N(const Setup &, const b200mf_operator &, void *, const void *, uint64_t, \
                         uint64_t, cudaStream_t, double *, bool, uint32_t, const uint32_t *);
DECL_N(2) DECL_N(3) DECL_N(4) DECL_N(5) DECL_N(6) DECL_N(7) DECL_N(8) DECL_N(9)
#undef DECL_N

#define DECL_N(N)                                                                           \
  int launch_bulk_n##N(const Setup &, const b200mf_operator &, void *, const void *, cudaStream_t, double *);
DECL_N(2) DECL_N(3) DECL_N(4) DECL_N(5) DECL_N(6) DECL_N(7) DECL_N(8) DECL_N(9)
#undef DECL_N

int launch_bulk(const Setup &s, const b200mf_operator &op, void *dst, const void *src, cudaStream_t st,
                double *dot) {
  switch (s.n) {
    case 2: return launch_bulk_n2(s, op, dst, src, st, dot);
    case 3: return launch_bulk_n3(s, op, dst, src, st, dot);
    case 4: return launch_bulk_n4(s, op, dst, src, st, dot);
    case 5: return launch_bulk_n5(s, op, dst, src, st, dot);
    case 6: return launch_bulk_n6(s, op, dst, src, st, dot);
    case 7: return launch_bulk_n7(s, op, dst, src, st, dot);
    case 8: return launch_bulk_n8(s, op, dst, src, st, dot);
    case 9: return launch_bulk_n9(s, op, dst, src, st, dot);
  }
  set_error("unsupported degree %d", s.degree);
  return B200MF_ERR_UNSUPPORTED;
}

#define DECL_N(N) int debug_resolve_n##N(int, int, unsigned, int, double *);
DECL_N(2) DECL_N(3) DECL_N(4) DECL_N(5) DECL_N(6) DECL_N(7) DECL_N(8) DECL_N(9)
#undef DECL_N

int launch_bricks(const Setup &s, const b200mf_operator &op, void *dst, const void *src,
                  uint64_t bb, uint64_t nb, cudaStream_t st, double *dot, bool ow, uint32_t geom,
                  const uint32_t *list) {
  switch (s.n) {
    case 2: return launch_bricks_n2(s, op, dst, src, bb, nb, st, dot, ow, geom, list);
    case 3: return launch_bricks_n3(s, op, dst, src, bb, nb, st, dot, ow, geom, list);
    case 4: return launch_bricks_n4(s, op, dst, src, bb, nb, st, dot, ow, geom, list);
    case 5: return launch_bricks_n5(s, op, dst, src, bb, nb, st, dot, ow, geom, list);
    case 6: return launch_bricks_n6(s, op, dst, src, bb, nb, st, dot, ow, geom, list);
    case 7: return launch_bricks_n7(s, op, dst, src, bb, nb, st, dot, ow, geom, list);
    case 8: return launch_bricks_n8(s, op, dst, src, bb, nb, st, dot, ow, geom, list);
    case 9: return launch_bricks_n9(s, op, dst, src, bb, nb, st, dot, ow, geom, list);
  }
  set_error("unsupported degree %d", s.degree);
  return B200MF_ERR_UNSUPPORTED;
}

static int launch_cells_part(const Setup &s, const b200mf_operator &op, void *dst, const void *src,
                             uint64_t b, uint64_t e, cudaStream_t st, bool diag, double *dot, bool masked) {
  if (e <= b) return B200MF_OK;
  switch (s.n) {
    case 2: return launch_cells_n2(s, op, dst, src, b, e, st, diag, dot, masked);
    case 3: return launch_cells_n3(s, op, dst, src, b, e, st, diag, dot, masked);
    case 4: return launch_cells_n4(s, op, dst, src, b, e, st, diag, dot, masked);
    case 5: return launch_cells_n5(s, op, dst, src, b, e, st, diag, dot, masked);
    case 6: return launch_cells_n6(s, op, dst, src, b, e, st, diag, dot, masked);
    case 7: return launch_cells_n7(s, op, dst, src, b, e, st, diag, dot, masked);
    case 8: return launch_cells_n8(s, op, dst, src, b, e, st, diag, dot, masked);
    case 9: return launch_cells_n9(s, op, dst, src, b, e, st, diag, dot, masked);
  }
  set_error("unsupported degree %d", s.degree);
  return B200MF_ERR_UNSUPPORTED;
}

// kernel selection per cell range: only the cells inside [masked_begin, masked_end) may carry a
// hanging-node mask and need the kernel that resolves them
static int launch_cells(const Setup &s, const b200mf_operator &op, void *dst, const void *src,
                        uint64_t b, uint64_t e, cudaStream_t st, bool diag, double *dot) {
  if (!s.any_mask) return launch_cells_part(s, op, dst, src, b, e, st, diag, dot, false);
  const uint64_t mb = std::min(std::max(b, s.masked_begin), e), me = std::max(std::min(e, s.masked_end), mb);
  int rc = launch_cells_part(s, op, dst, src, b, mb, st, diag, dot, false);
  if (rc != B200MF_OK) return rc;
  rc = launch_cells_part(s, op, dst, src, mb, me, st, diag, dot, true);
  if (rc != B200MF_OK) return rc;
  return launch_cells_part(s, op, dst, src, me, e, st, diag, dot, false);
}

static bool bricks_enabled(const Setup &s, const b200mf_operator &op) {
  static const bool bricks_off = std::getenv("B200MF_KERNEL") != nullptr &&
                                 std::string(std::getenv("B200MF_KERNEL")) != "brick";
  return s.n_bricks != 0 && !bricks_off && op.grad_coefficient == nullptr && op.mass_coefficient == nullptr;
}

static bool selective_zero(const Setup &s, const b200mf_operator &op) {
  static const bool full_memset = std::getenv("B200MF_FULL_MEMSET") != nullptr;
  return bricks_enabled(s, op) && s.have_zero_list && !full_memset;
}

int launch_cell_loop(const Setup &s, const b200mf_operator &op, void *dst, const void *src,
                     uint64_t cell_begin, uint64_t cell_end, cudaStream_t stream,
                     double *dot_accum, bool dst_zeroed) {
  // Cartesian cells + constant coefficients: whole bricks of cells go to the brick kernel,
  // whatever is left of the range to the per-cell kernels (B200MF_KERNEL=v1|plane: A/B runs)
  if (s.n_q_1d > s.n) return launch_overint(s, op, dst, src, cell_begin, cell_end, stream, dot_accum, false);
  if (!bricks_enabled(s, op))
    return launch_cells(s, op, dst, src, cell_begin, cell_end, stream, false, dot_accum);
  const uint64_t W = (uint64_t)s.brick_b * s.brick_b * s.brick_b;
  uint64_t pos = cell_begin;
  for (const Setup::BrickRun &run : s.brick_runs) {
    if (run.cell_end <= pos) continue;
    if (run.cell_begin >= cell_end) break;
    uint64_t rb = std::max(run.cell_begin, pos);
    rb = run.cell_begin + (rb - run.cell_begin + W - 1) / W * W;
    uint64_t re = std::min(run.cell_end, cell_end);
    re = run.cell_begin + (re - run.cell_begin) / W * W;
    if (dst_zeroed && selective_zero(s, op) &&
        (rb != std::max(run.cell_begin, pos) || re != std::min(run.cell_end, cell_end))) {
      // the dofs a brick stores were not zeroed by b200mf_vmult_prepare: its cells must not be
      // split over pieces (they would be added by the per-cell kernels)
      set_error("vmult piece [%llu, %llu) cuts a brick of %llu cells: align the pieces of a vmult to "
                "b200mf_setup_info.cells_per_brick", (unsigned long long)cell_begin,
                (unsigned long long)cell_end, (unsigned long long)W);
      return B200MF_ERR_INVALID;
    }
    if (re <= rb) continue;
    if (pos < rb) {
      int rc = launch_cells(s, op, dst, src, pos, rb, stream, false, dot_accum);
      if (rc != B200MF_OK) return rc;
    }
    int rc = launch_bricks(s, op, dst, src, run.first_brick + (rb - run.cell_begin) / W, (re - rb) / W,
                           stream, dot_accum, dst_zeroed, run.geom);
    if (rc != B200MF_OK) return rc;
    pos = re;
  }
  if (pos < cell_end) return launch_cells(s, op, dst, src, pos, cell_end, stream, false, dot_accum);
  return B200MF_OK;
}

int launch_compute_diagonal(const Setup &s, const b200mf_operator &op, void *diag,
                            cudaStream_t stream) {
  if (s.n_q_1d > s.n) return launch_overint(s, op, diag, diag, 0, s.n_cells, stream, nullptr, true);
  return launch_cells(s, op, diag, diag, 0, s.n_cells, stream, true, nullptr);
}

// ---------------------------------------------------------------------------------------
// device-side geometry setup (replaces the host FEValues loop of
// portable_matrix_free.templates.h:267-346)
// ---------------------------------------------------------------------------------------
template <int dim>
__device__ inline void q1_jacobian(const double *v /*[2^dim][dim]*/, const double *xi,
                                   double J[dim][dim], double *x /*[dim] or null*/) {
  for (int d = 0; d < dim; ++d) {
    for (int e = 0; e < dim; ++e) J[d][e] = 0.0;
    if (x) x[d] = 0.0;
  }
  for (int k = 0; k < (1 << dim); ++k) {
    double N = 1.0, dN[dim];
    for (int e = 0; e < dim; ++e) dN[e] = 1.0;
    for (int d = 0; d < dim; ++d) {
      const int b = (k >> d) & 1;
      const double f = b ? xi[d] : 1.0 - xi[d];
      const double df = b ? 1.0 : -1.0;
      N *= f;
      for (int e = 0; e < dim; ++e) dN[e] *= (e == d) ? df : f;
    }
    for (int d = 0; d < dim; ++d) {
      if (x) x[d] += N * v[k * dim + d];
      for (int e = 0; e < dim; ++e) J[d][e] += dN[e] * v[k * dim + d];
    }
  }
}

template <int dim>
__device__ inline double invert(const double J[dim][dim], double Ji[dim][dim]) {
  if (dim == 2) {
    const double det = J[0][0] * J[1][1] - J[0][1] * J[1][0];
    const double r = 1.0 / det;
    Ji[0][0] = J[1][1] * r; Ji[0][1] = -J[0][1] * r;
    Ji[1][0] = -J[1][0] * r; Ji[1][1] = J[0][0] * r;
    return det;
  } else {
    const double c00 = J[1][1] * J[2][2] - J[1][2] * J[2][1];
    const double c01 = J[1][2] * J[2][0] - J[1][0] * J[2][2];
    const double c02 = J[1][0] * J[2][1] - J[1][1] * J[2][0];
    const double det = J[0][0] * c00 + J[0][1] * c01 + J[0][2] * c02;
    const double r = 1.0 / det;
    Ji[0][0] = c00 * r;
    Ji[1][0] = c01 * r;
    Ji[2][0] = c02 * r;
    Ji[0][1] = (J[0][2] * J[2][1] - J[0][1] * J[2][2]) * r;
    Ji[1][1] = (J[0][0] * J[2][2] - J[0][2] * J[2][0]) * r;
    Ji[2][1] = (J[0][1] * J[2][0] - J[0][0] * J[2][1]) * r;
    Ji[0][2] = (J[0][1] * J[1][2] - J[0][2] * J[1][1]) * r;
    Ji[1][2] = (J[0][2] * J[1][0] - J[0][0] * J[1][2]) * r;
    Ji[2][2] = (J[0][0] * J[1][1] - J[0][1] * J[1][0]) * r;
    return det;
  }
}

// metric (layout: metric_offset) = w_q det(J) (J^-1 J^-T)_s (s over the upper triangle), jxw[cell][q]
template <int dim, typename Number>
__global__ void general_geometry_from_vertices(const double *vertices, const double *qp1d,
                                               const double *qw1d, int n, uint64_t n_cells,
                                               Number *metric, Number *jxw) {
  const int nq = dim == 2 ? n * n : n * n * n;
  const uint64_t gid = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (gid >= n_cells * nq) return;
  const uint64_t cell = gid / nq;
  const int q = (int)(gid - cell * nq);
  int qi[3] = {q % n, (q / n) % n, q / (n * n)};
  double xi[dim], w = 1.0;
  for (int d = 0; d < dim; ++d) { xi[d] = qp1d[qi[d]]; w *= qw1d[qi[d]]; }
  double J[dim][dim], Ji[dim][dim];
  q1_jacobian<dim>(vertices + cell * (1 << dim) * dim, xi, J, nullptr);
  const double det = invert<dim>(J, Ji);
  // Ji[e][d] = d xi_e / d x_d
  int s = 0;
  for (int e = 0; e < dim; ++e)
    for (int f = e; f < dim; ++f, ++s) {
      double m = 0.0;
      for (int d = 0; d < dim; ++d) m += Ji[e][d] * Ji[f][d];
      metric[metric_offset<dim>(n, cell, s, q)] = Number(m * det * w);
    }
  jxw[cell * nq + q] = Number(det * w);
}

// same from the arrays Portable::MatrixFree stores (inv_jacobian(q,cell,d,e), JxW(q,cell))
template <int dim, typename Number>
__global__ void general_geometry_from_jacobians(const double *inv_jac, const double *jxw_in,
                                                int nq, int n1d, uint64_t n_cells,
                                                Number *metric, Number *jxw) {
  const uint64_t gid = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (gid >= n_cells * nq) return;
  const uint64_t cell = gid / nq;
  const int q = (int)(gid - cell * nq);
  const double *Ji = inv_jac + gid * dim * dim; // [e][d]
  const double jw = jxw_in[gid];
  int s = 0;
  for (int e = 0; e < dim; ++e)
    for (int f = e; f < dim; ++f, ++s) {
      double m = 0.0;
      for (int d = 0; d < dim; ++d) m += Ji[e * dim + d] * Ji[f * dim + d];
      metric[metric_offset<dim>(n1d, cell, s, q)] = Number(m * jw);
    }
  jxw[cell * nq + q] = Number(jw);
}

template <int dim>
__global__ void quadrature_points_kernel(const double *vertices, const double *qp1d, int n,
                                         uint64_t n_cells, double *out) {
  const int nq = dim == 2 ? n * n : n * n * n;
  const uint64_t gid = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (gid >= n_cells * nq) return;
  const uint64_t cell = gid / nq;
  const int q = (int)(gid - cell * nq);
  int qi[3] = {q % n, (q / n) % n, q / (n * n)};
  double xi[dim], J[dim][dim], x[dim];
  for (int d = 0; d < dim; ++d) xi[d] = qp1d[qi[d]];
  q1_jacobian<dim>(vertices + cell * (1 << dim) * dim, xi, J, x);
  for (int d = 0; d < dim; ++d) out[gid * dim + d] = x[d];
}

// entries flagged B200MF_L2G_CONSTRAINED whose index part is out of range (e.g. deal.II's
// numbers::invalid_unsigned_int) are redirected to index 0 so that the branch-free kernels
// may form the address of every entry
__global__ void sanitize_l2g_kernel(uint32_t *l2g, uint64_t count, uint32_t n_local) {
  const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= count) return;
  const uint32_t v = l2g[i];
  if ((v & B200MF_L2G_CONSTRAINED) && (v & ~B200MF_L2G_CONSTRAINED) >= n_local)
    l2g[i] = B200MF_L2G_CONSTRAINED;
}

template <typename Number>
__global__ void copy_constrained_kernel(Number *dst, const Number *src, const uint32_t *idx,
                                        uint64_t n, double *dot_accum) {
  const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  double acc = 0.0;
  if (i < n) {
    const Number v = src[idx[i]];
    dst[idx[i]] = v;
    acc = double(v) * double(v);
  }
  if (dot_accum != nullptr) {
    acc = block_sum(acc);
    if (threadIdx.x == 0 && acc != 0.0) atomicAdd(dot_accum, acc);
  }
}
template <typename Number>
__global__ void set_constrained_kernel(Number *dst, Number value, const uint32_t *idx,
                                       uint64_t n) {
  const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) dst[idx[i]] = value;
}

// ---------------------------------------------------------------------------------------
template <typename T>
static int dev_alloc_copy(T **dptr, const T *host, size_t count, Setup &s, uint64_t *bucket) {
  B200MF_CUDA_CHECK(cudaMalloc((void **)dptr, std::max<size_t>(count, 1) * sizeof(T)));
  if (count) B200MF_CUDA_CHECK(cudaMemcpy(*dptr, host, count * sizeof(T), cudaMemcpyHostToDevice));
  s.device_bytes += count * sizeof(T);
  if (bucket) *bucket += count * sizeof(T);
  return B200MF_OK;
}

template <typename Number>
static int upload_converted(void **dptr, const std::vector<double> &v, Setup &s, uint64_t *bucket) {
  std::vector<Number> tmp(v.begin(), v.end());
  Number *d = nullptr;
  int rc = dev_alloc_copy<Number>(&d, tmp.data(), tmp.size(), s, bucket);
  *dptr = d;
  return rc;
}

// classify cells from their Q1 vertices and build the compressed affine table
// (cf. MappingInfo cell types, matrix_free/mapping_info.templates.h:428-573)
static int classify_and_compress(const b200mf_setup_desc &d, Setup &s,
                                 std::vector<double> &table, std::vector<uint32_t> &geom_id) {
  const int dim = d.dim, nv = 1 << dim;
  int kind = B200MF_CELLS_CARTESIAN;
  const double tol = 1e-12;
  // pass 1: classification
  for (uint64_t c = 0; c < d.n_cells && kind != B200MF_CELLS_GENERAL; ++c) {
    const double *v = d.cell_vertices + c * nv * dim;
    double E[3][3], scale = 0.0;
    for (int e = 0; e < dim; ++e)
      for (int k = 0; k < dim; ++k) {
        E[k][e] = v[(1 << e) * dim + k] - v[k]; // J[k][e] = dx_k/dxi_e
        scale = std::max(scale, std::fabs(E[k][e]));
      }
    bool affine = true;
    for (int k = 0; k < nv && affine; ++k)
      for (int x = 0; x < dim; ++x) {
        double pred = v[x];
        for (int e = 0; e < dim; ++e)
          if ((k >> e) & 1) pred += E[x][e];
        if (std::fabs(pred - v[k * dim + x]) > tol * scale) { affine = false; break; }
      }
    if (!affine) { kind = B200MF_CELLS_GENERAL; break; }
    for (int e = 0; e < dim; ++e)
      for (int k = 0; k < dim; ++k)
        if (k != e && std::fabs(E[k][e]) > tol * scale) kind = B200MF_CELLS_AFFINE;
  }
  s.cell_kind = kind;
  if (kind == B200MF_CELLS_GENERAL) return B200MF_OK;

  // pass 2: table of distinct Jacobians
  const int NS = dim * (dim + 1) / 2;
  const int entry = (kind == B200MF_CELLS_CARTESIAN ? dim : NS) + 1;
  std::map<std::vector<int64_t>, uint32_t> seen;
  geom_id.resize(d.n_cells);
  std::vector<int64_t> key(dim * dim);
  double last[9];
  uint32_t last_id = 0;
  bool have_last = false;
  for (uint64_t c = 0; c < d.n_cells; ++c) {
    const double *v = d.cell_vertices + c * nv * dim;
    double J[3][3], flat[9];
    for (int e = 0; e < dim; ++e)
      for (int k = 0; k < dim; ++k) {
        J[k][e] = v[(1 << e) * dim + k] - v[k];
        flat[k * dim + e] = J[k][e];
      }
    if (have_last) {
      bool same = true;
      for (int i = 0; i < dim * dim; ++i)
        if (std::fabs(flat[i] - last[i]) > 1e-13 * std::fabs(last[0]) + 1e-300) { same = false; break; }
      if (same) { geom_id[c] = last_id; continue; }
    }
    double scale = 0.0;
    for (int i = 0; i < dim * dim; ++i) scale = std::max(scale, std::fabs(flat[i]));
    // quantise relative to a power-of-two scale so that round-off equal cells collide
    int ex;
    std::frexp(scale, &ex);
    for (int i = 0; i < dim * dim; ++i)
      key[i] = (int64_t)std::llround(std::ldexp(flat[i], 40 - ex));
    key.resize(dim * dim + 1);
    key[dim * dim] = ex;
    auto it = seen.find(key);
    uint32_t id;
    if (it == seen.end()) {
      id = (uint32_t)seen.size();
      seen.emplace(key, id);
      // entry values
      double Ji[3][3], det;
      if (dim == 2) {
        det = J[0][0] * J[1][1] - J[0][1] * J[1][0];
        Ji[0][0] = J[1][1] / det; Ji[0][1] = -J[0][1] / det;
        Ji[1][0] = -J[1][0] / det; Ji[1][1] = J[0][0] / det;
      } else {
        const double c00 = J[1][1] * J[2][2] - J[1][2] * J[2][1];
        const double c01 = J[1][2] * J[2][0] - J[1][0] * J[2][2];
        const double c02 = J[1][0] * J[2][1] - J[1][1] * J[2][0];
        det = J[0][0] * c00 + J[0][1] * c01 + J[0][2] * c02;
        Ji[0][0] = c00 / det; Ji[1][0] = c01 / det; Ji[2][0] = c02 / det;
        Ji[0][1] = (J[0][2] * J[2][1] - J[0][1] * J[2][2]) / det;
        Ji[1][1] = (J[0][0] * J[2][2] - J[0][2] * J[2][0]) / det;
        Ji[2][1] = (J[0][1] * J[2][0] - J[0][0] * J[2][1]) / det;
        Ji[0][2] = (J[0][1] * J[1][2] - J[0][2] * J[1][1]) / det;
        Ji[1][2] = (J[0][2] * J[1][0] - J[0][0] * J[1][2]) / det;
        Ji[2][2] = (J[0][0] * J[1][1] - J[0][1] * J[1][0]) / det;
      }
      if (kind == B200MF_CELLS_CARTESIAN) {
        for (int e = 0; e < dim; ++e) table.push_back(Ji[e][e] * Ji[e][e] * det);
      } else {
        for (int e = 0; e < dim; ++e)
          for (int f = e; f < dim; ++f) {
            double m = 0.0;
            for (int x = 0; x < dim; ++x) m += Ji[e][x] * Ji[f][x];
            table.push_back(m * det);
          }
      }
      table.push_back(det);
    } else {
      id = it->second;
    }
    key.resize(dim * dim);
    geom_id[c] = id;
    std::memcpy(last, flat, sizeof(double) * dim * dim);
    last_id = id;
    have_last = true;
  }
  s.n_geom = table.size() / entry;
  return B200MF_OK;
}

} // namespace b200mf

using namespace b200mf;

namespace b200mf {
int copy_constrained_impl(const Setup &s, void *dst, const void *src, cudaStream_t st,
                                 double *dot_accum) {
  if (s.n_constrained == 0) return B200MF_OK;
  const unsigned blocks = (unsigned)((s.n_constrained + 255) / 256);
  if (s.number == B200MF_F64)
    copy_constrained_kernel<double><<<blocks, 256, 0, st>>>(
        (double *)dst, (const double *)src, s.d_constrained, s.n_constrained, dot_accum);
  else
    copy_constrained_kernel<float><<<blocks, 256, 0, st>>>(
        (float *)dst, (const float *)src, s.d_constrained, s.n_constrained, dot_accum);
  count_launch();
  B200MF_CUDA_CHECK(cudaGetLastError());
  return B200MF_OK;
}

int set_constrained_impl(const Setup &s, void *dst, double value, cudaStream_t st) {
  if (s.n_constrained == 0) return B200MF_OK;
  const unsigned blocks = (unsigned)((s.n_constrained + 255) / 256);
  if (s.number == B200MF_F64)
    set_constrained_kernel<double><<<blocks, 256, 0, st>>>((double *)dst, value, s.d_constrained,
                                                           s.n_constrained);
  else
    set_constrained_kernel<float><<<blocks, 256, 0, st>>>((float *)dst, (float)value,
                                                          s.d_constrained, s.n_constrained);
  count_launch();
  B200MF_CUDA_CHECK(cudaGetLastError());
  return B200MF_OK;
}

int vmult_prepare_impl(const Setup &s, const b200mf_operator &op, void *dst, cudaStream_t st) {
  if (selective_zero(s, op)) {
    // the brick kernel stores every dof that only one brick touches: zero the others
    if (s.n_zero_list == 0) return B200MF_OK;
    const unsigned blocks = (unsigned)((s.n_zero_list + 255) / 256);
    if (s.number == B200MF_F64)
      set_constrained_kernel<double><<<blocks, 256, 0, st>>>((double *)dst, 0.0, s.d_zero_list, s.n_zero_list);
    else
      set_constrained_kernel<float><<<blocks, 256, 0, st>>>((float *)dst, 0.0f, s.d_zero_list, s.n_zero_list);
    count_launch();
    B200MF_CUDA_CHECK(cudaGetLastError());
    return B200MF_OK;
  }
  B200MF_CUDA_CHECK(cudaMemsetAsync(dst, 0, (s.n_owned + s.n_ghost) * number_size(s.number), st));
  return B200MF_OK;
}

bool coloured_enabled(const Setup &s, const b200mf_operator &op) {
  return s.colouring.ready && s.colouring.enabled && bricks_enabled(s, op);
}

// zero what no brick stores (ghost section, dofs of the per-cell kernels, untouched dofs)
int coloured_prepare(const Setup &s, void *dst, cudaStream_t st) {
  const Setup::Colouring &K = s.colouring;
  const size_t ns = number_size(s.number);
  if (s.n_ghost)
    B200MF_CUDA_CHECK(cudaMemsetAsync(static_cast<char *>(dst) + s.n_owned * ns, 0, s.n_ghost * ns, st));
  if (K.n_zero) {
    const unsigned blocks = (unsigned)((K.n_zero + 255) / 256);
    if (s.number == B200MF_F64)
      set_constrained_kernel<double><<<blocks, 256, 0, st>>>((double *)dst, 0.0, K.d_zero, K.n_zero);
    else
      set_constrained_kernel<float><<<blocks, 256, 0, st>>>((float *)dst, 0.0f, K.d_zero, K.n_zero);
    count_launch();
    B200MF_CUDA_CHECK(cudaGetLastError());
  }
  return B200MF_OK;
}

int launch_coloured(const Setup &s, const b200mf_operator &op, void *dst, const void *src, int piece,
                    cudaStream_t st, double *dot_accum) {
  const Setup::Colouring &K = s.colouring;
  for (const auto &r : K.general) {
    if (piece >= 0 && r.piece != piece) continue;
    int rc = launch_cells(s, op, dst, src, r.begin, r.end, st, false, dot_accum);
    if (rc != B200MF_OK) return rc;
  }
  for (const auto &l : K.launches) {
    if (piece >= 0 && l.piece != piece) continue;
    int rc = launch_bricks(s, op, dst, src, l.offset, l.count, st, dot_accum, true, l.geom, K.d_list);
    if (rc != B200MF_OK) return rc;
  }
  return B200MF_OK;
}

// bulk brick path usable for this call?  (bulk copies need 16-byte aligned vectors)
bool bulk_enabled(const Setup &s, const b200mf_operator &op, const void *dst, const void *src) {
  static const bool legacy = std::getenv("B200MF_KERNEL") != nullptr && std::string(std::getenv("B200MF_KERNEL")) == "brick";
  return s.bulk.ready && s.bulk.enabled && !legacy && bricks_enabled(s, op) && dst != src &&
         (reinterpret_cast<uintptr_t>(dst) & 15) == 0 && (reinterpret_cast<uintptr_t>(src) & 15) == 0;
}

// vmult through the bulk brick kernel: zero what no brick stores, cells outside bricks with the
// per-cell kernels (atomics), then every brick in one launch
int vmult_bulk_impl(const Setup &s, const b200mf_operator &op, void *dst, const void *src,
                    cudaStream_t st, double *dot_accum) {
  const Setup::Bulk &B = s.bulk;
  const size_t ns = number_size(s.number);
  if (s.n_ghost)
    B200MF_CUDA_CHECK(cudaMemsetAsync(static_cast<char *>(dst) + s.n_owned * ns, 0, s.n_ghost * ns, st));
  if (B.n_zero) {
    const unsigned blocks = (unsigned)((B.n_zero + 255) / 256);
    if (s.number == B200MF_F64)
      set_constrained_kernel<double><<<blocks, 256, 0, st>>>((double *)dst, 0.0, B.d_zero, B.n_zero);
    else
      set_constrained_kernel<float><<<blocks, 256, 0, st>>>((float *)dst, 0.0f, B.d_zero, B.n_zero);
    count_launch();
    B200MF_CUDA_CHECK(cudaGetLastError());
  }
  for (const auto &r : B.general_ranges) {
    int rc = launch_cells(s, op, dst, src, r.first, r.second, st, false, dot_accum);
    if (rc != B200MF_OK) return rc;
  }
  return launch_bulk(s, op, dst, src, st, dot_accum);
}

int vmult_impl(const Setup &s, const b200mf_operator &op, void *dst, const void *src,
               cudaStream_t st, double *dot_accum) {
  if (bulk_enabled(s, op, dst, src)) {
    int rc = vmult_bulk_impl(s, op, dst, src, st, dot_accum);
    if (rc != B200MF_OK) return rc;
    return copy_constrained_impl(s, dst, src, st, dot_accum);
  }
  if (coloured_enabled(s, op)) {
    int rc = coloured_prepare(s, dst, st);
    if (rc != B200MF_OK) return rc;
    rc = launch_coloured(s, op, dst, src, -1, st, dot_accum);
    if (rc != B200MF_OK) return rc;
    return copy_constrained_impl(s, dst, src, st, dot_accum);
  }
  int rc0 = vmult_prepare_impl(s, op, dst, st);
  if (rc0 != B200MF_OK) return rc0;
  int rc = launch_cell_loop(s, op, dst, src, 0, s.n_cells, st, dot_accum, true);
  if (rc != B200MF_OK) return rc;
  return copy_constrained_impl(s, dst, src, st, dot_accum);
}
} // namespace b200mf


extern "C" {

const char *b200mf_last_error(void) { return g_error.c_str(); }
int b200mf_version(void) { return B200MF_VERSION; }
uint64_t b200mf_kernel_launch_count(void) { return g_launch_count.load(); }

int b200mf_setup_create(const b200mf_setup_desc *d, b200mf_setup **out) {
  B200MF_REQUIRE(d && out, "null argument");
  B200MF_REQUIRE(d->dim == 2 || d->dim == 3, "dim must be 2 or 3 (got %d)", d->dim);
  B200MF_REQUIRE(d->degree >= 1 && d->degree <= 8, "degree must be in 1..8 (got %d)", d->degree);
  // AssertThrow(n_q_points_1d >= fe_degree+1), portable_matrix_free.templates.h:1243.  More points than
  // fe_degree+1 (over-integration) run the non-collocation kernel of overint.cu
  B200MF_REQUIRE(d->n_q_points_1d >= d->degree + 1 && d->n_q_points_1d <= 12,
                 "n_q_points_1d must be in [degree+1, 12] (got %d for degree %d)", d->n_q_points_1d,
                 d->degree);
  const bool overint = d->n_q_points_1d != d->degree + 1;
  if (overint && (d->shape_values || d->shape_gradients_collocation || d->quadrature_weights)) {
    set_error("n_q_points_1d > degree+1 uses the engine's own shape data (caller-provided arrays are for degree+1 "
              "points)");
    return B200MF_ERR_UNSUPPORTED;
  }
  B200MF_REQUIRE(d->number == B200MF_F64 || d->number == B200MF_F32, "bad number type");
  B200MF_REQUIRE(d->local_to_global || d->n_cells == 0, "local_to_global is null");
  B200MF_REQUIRE(d->n_owned_dofs + d->n_ghost_dofs < 0x80000000ull,
                 "more than 2^31-1 local dofs per process are not supported");
  int dev_count = 0;
  if (cudaGetDeviceCount(&dev_count) != cudaSuccess || dev_count == 0) {
    set_error("no CUDA device available: the engine has no CPU fallback");
    return B200MF_ERR_CUDA;
  }
  b200mf_setup *h = new b200mf_setup();
  // every early return below (B200MF_CUDA_CHECK / B200MF_REQUIRE / TRY) goes through this guard:
  // it frees whatever the half-built setup already owns, incl. the temporaries d_qp / d_qw
  double *d_qp = nullptr, *d_qw = nullptr;
  struct Guard {
    b200mf_setup *&h;
    double *&qp, *&qw;
    ~Guard() {
      cudaFree(qp);
      cudaFree(qw);
      if (h) b200mf_setup_destroy(h);
    }
  } guard{h, d_qp, d_qw};
  Setup &s = h->impl;
  s.dim = d->dim; s.degree = d->degree; s.n = d->degree + 1; s.number = d->number;
  s.n_q_1d = d->n_q_points_1d;
  s.n_cells = d->n_cells; s.n_owned = d->n_owned_dofs; s.n_ghost = d->n_ghost_dofs;
  s.n_constrained = d->n_constrained_dofs; s.n_cells_interior = d->n_cells_interior;
  s.dofs_per_cell = 1;
  for (int i = 0; i < s.dim; ++i) s.dofs_per_cell *= s.n;
  const int n = s.n, nq = s.dofs_per_cell;
  const int Q = s.n_q_1d; // quadrature points per direction (= n unless over-integrating)
  int qpc = 1;            // ... and per cell
  for (int i = 0; i < s.dim; ++i) qpc *= Q;
  int rc = B200MF_OK;
#define TRY(x) do { rc = (x); if (rc != B200MF_OK) return rc; } while (0)

  // --- shape data
  build_fe_q_shape_data(s.degree, s.shape_values, s.shape_grad_colloc, s.q_weights,
                        s.q_points_1d, s.subface);
  if (d->shape_values) s.shape_values.assign(d->shape_values, d->shape_values + n * n);
  if (d->shape_gradients_collocation)
    s.shape_grad_colloc.assign(d->shape_gradients_collocation,
                               d->shape_gradients_collocation + n * n);
  if (d->quadrature_weights) s.q_weights.assign(d->quadrature_weights, d->quadrature_weights + n);
  if (d->subface_interpolation_matrix)
    s.subface.assign(d->subface_interpolation_matrix, d->subface_interpolation_matrix + n * n);

  // --- indices
  // the index list is padded with kL2gPadCells cells of "constrained" entries so that the
  // plane kernel can copy whole cell groups past the last cell without bounds checks
  {
    const size_t real = (size_t)d->n_cells * nq, pad = (size_t)kL2gPadCells * nq;
    B200MF_CUDA_CHECK(cudaMalloc((void **)&s.d_l2g, (real + pad) * sizeof(uint32_t)));
    if (real)
      B200MF_CUDA_CHECK(cudaMemcpy(s.d_l2g, d->local_to_global, real * sizeof(uint32_t),
                                   cudaMemcpyHostToDevice));
    {
      // padding entries: "constrained" marker on index 0 (read as 0, scatter adds 0 to dst[0])
      std::vector<uint32_t> padv(pad, B200MF_L2G_CONSTRAINED);
      B200MF_CUDA_CHECK(cudaMemcpy(s.d_l2g + real, padv.data(), pad * sizeof(uint32_t),
                                   cudaMemcpyHostToDevice));
    }
    if (real) {
      sanitize_l2g_kernel<<<(unsigned)((real + 255) / 256), 256>>>(
          s.d_l2g, real, (uint32_t)(d->n_owned_dofs + d->n_ghost_dofs));
      count_launch();
      B200MF_CUDA_CHECK(cudaGetLastError());
    }
    s.device_bytes += (real + pad) * sizeof(uint32_t);
    s.index_bytes += (real + pad) * sizeof(uint32_t);
  }
  if (d->constraint_mask) {
    s.masked_begin = d->n_cells;
    s.masked_end = 0;
    for (uint64_t c = 0; c < d->n_cells; ++c)
      if (d->constraint_mask[c]) {
        s.any_mask = true;
        s.masked_begin = std::min(s.masked_begin, c);
        s.masked_end = c + 1;
      }
    if (s.any_mask)
      TRY(dev_alloc_copy<uint16_t>(&s.d_mask, d->constraint_mask, d->n_cells, s, &s.index_bytes));
  }
  if (d->n_constrained_dofs)
    TRY(dev_alloc_copy<uint32_t>(&s.d_constrained, d->constrained_dofs, d->n_constrained_dofs, s,
                                 &s.index_bytes));
  if (s.number == B200MF_F64) TRY(upload_converted<double>(&s.d_weights, s.subface, s, nullptr));
  else                        TRY(upload_converted<float>(&s.d_weights, s.subface, s, nullptr));
  {
    // tables of the sum-factorised diagonal: SS = S.*S, GG = (SD).*(SD), SG = S.*(SD), [i*n+q]
    std::vector<double> tb(3 * n * n);
    for (int i = 0; i < n; ++i)
      for (int q = 0; q < n; ++q) {
        double g = 0.0;
        for (int r = 0; r < n; ++r) g += s.shape_values[i * n + r] * s.shape_grad_colloc[r * n + q];
        const double sv = s.shape_values[i * n + q];
        tb[i * n + q] = sv * sv;
        tb[n * n + i * n + q] = g * g;
        tb[2 * n * n + i * n + q] = sv * g;
      }
    if (s.number == B200MF_F64) TRY(upload_converted<double>(&s.d_diag_tables, tb, s, nullptr));
    else                        TRY(upload_converted<float>(&s.d_diag_tables, tb, s, nullptr));
  }

  if (overint) {
    std::vector<double> S, Dq;
    build_overint_shape_data(s.degree, Q, S, Dq, s.q_weights, s.q_points_1d);
    std::vector<double> tb(S);
    tb.insert(tb.end(), Dq.begin(), Dq.end());
    tb.insert(tb.end(), s.q_weights.begin(), s.q_weights.end());
    if (s.number == B200MF_F64) TRY(upload_converted<double>(&s.d_overint_tables, tb, s, nullptr));
    else                        TRY(upload_converted<float>(&s.d_overint_tables, tb, s, nullptr));
  }

  // --- geometry
  TRY(dev_alloc_copy<double>(&d_qp, s.q_points_1d.data(), Q, s, nullptr));
  TRY(dev_alloc_copy<double>(&d_qw, s.q_weights.data(), Q, s, nullptr));
  const size_t NS = s.dim * (s.dim + 1) / 2;
  const size_t ns = number_size(s.number);
  const uint64_t total_q = d->n_cells * (uint64_t)qpc;
  const unsigned blocks = (unsigned)((total_q + 255) / 256);
  if (d->geometry == B200MF_GEOMETRY_Q1_VERTICES) {
    B200MF_REQUIRE(d->cell_vertices || d->n_cells == 0, "cell_vertices is null");
    std::vector<double> table;
    std::vector<uint32_t> geom_id;
    TRY(classify_and_compress(*d, s, table, geom_id));
    const size_t nvd = (size_t)(1 << s.dim) * s.dim;
    if (s.cell_kind == B200MF_CELLS_GENERAL) {
      double *d_vert = nullptr;
      B200MF_CUDA_CHECK(cudaMalloc((void **)&d_vert, std::max<size_t>(d->n_cells * nvd, 1) * 8));
      B200MF_CUDA_CHECK(cudaMemcpy(d_vert, d->cell_vertices, d->n_cells * nvd * 8,
                                   cudaMemcpyHostToDevice));
      B200MF_CUDA_CHECK(cudaMalloc(&s.d_metric, std::max<size_t>(total_q * NS * ns, 1)));
      B200MF_CUDA_CHECK(cudaMalloc(&s.d_jxw, std::max<size_t>(total_q * ns, 1)));
      s.geometry_bytes += total_q * (NS + 1) * ns;
      if (total_q) {
        if (s.dim == 2 && s.number == B200MF_F64)
          general_geometry_from_vertices<2, double><<<blocks, 256>>>(d_vert, d_qp, d_qw, Q, d->n_cells, (double *)s.d_metric, (double *)s.d_jxw);
        else if (s.dim == 2)
          general_geometry_from_vertices<2, float><<<blocks, 256>>>(d_vert, d_qp, d_qw, Q, d->n_cells, (float *)s.d_metric, (float *)s.d_jxw);
        else if (s.number == B200MF_F64)
          general_geometry_from_vertices<3, double><<<blocks, 256>>>(d_vert, d_qp, d_qw, Q, d->n_cells, (double *)s.d_metric, (double *)s.d_jxw);
        else
          general_geometry_from_vertices<3, float><<<blocks, 256>>>(d_vert, d_qp, d_qw, Q, d->n_cells, (float *)s.d_metric, (float *)s.d_jxw);
        count_launch();
      }
      B200MF_CUDA_CHECK(cudaDeviceSynchronize());
      cudaFree(d_vert);
    } else {
      if (s.number == B200MF_F64) TRY(upload_converted<double>(&s.d_geom_table, table, s, &s.geometry_bytes));
      else                        TRY(upload_converted<float>(&s.d_geom_table, table, s, &s.geometry_bytes));
      if (s.n_geom > 1)
        TRY(dev_alloc_copy<uint32_t>(&s.d_geom_id, geom_id.data(), geom_id.size(), s, &s.geometry_bytes));
      if (s.cell_kind == B200MF_CELLS_CARTESIAN && s.n_geom >= 1 && s.n_geom <= 64 && s.dim == 3 && !overint) {
        for (int i = 0; i < 4; ++i) s.geom0[i] = table[i];
        s.h_geom_table = table;
        if (s.n_geom > 1) s.h_geom_id = geom_id;
        TRY(build_bricks(*d, s));
        s.h_geom_id.clear();
        s.h_geom_id.shrink_to_fit();
      }
    }
    // keep the vertices for quadrature point queries while they are small
    if (d->n_cells * nvd * 8 <= (512ull << 20)) {
      s.h_vertices.assign(d->cell_vertices, d->cell_vertices + d->n_cells * nvd);
      s.has_vertices = true;
    }
  } else if (d->geometry == B200MF_GEOMETRY_JACOBIANS) {
    B200MF_REQUIRE(d->inv_jacobian && d->JxW, "inv_jacobian / JxW is null");
    s.cell_kind = B200MF_CELLS_GENERAL;
    double *d_ij = nullptr, *d_jw = nullptr;
    const size_t dd = (size_t)s.dim * s.dim;
    B200MF_CUDA_CHECK(cudaMalloc((void **)&d_ij, std::max<size_t>(total_q * dd, 1) * 8));
    B200MF_CUDA_CHECK(cudaMalloc((void **)&d_jw, std::max<size_t>(total_q, 1) * 8));
    B200MF_CUDA_CHECK(cudaMemcpy(d_ij, d->inv_jacobian, total_q * dd * 8, cudaMemcpyHostToDevice));
    B200MF_CUDA_CHECK(cudaMemcpy(d_jw, d->JxW, total_q * 8, cudaMemcpyHostToDevice));
    B200MF_CUDA_CHECK(cudaMalloc(&s.d_metric, std::max<size_t>(total_q * NS * ns, 1)));
    B200MF_CUDA_CHECK(cudaMalloc(&s.d_jxw, std::max<size_t>(total_q * ns, 1)));
    s.geometry_bytes += total_q * (NS + 1) * ns;
    if (total_q) {
      if (s.dim == 2 && s.number == B200MF_F64)
        general_geometry_from_jacobians<2, double><<<blocks, 256>>>(d_ij, d_jw, qpc, Q, d->n_cells, (double *)s.d_metric, (double *)s.d_jxw);
      else if (s.dim == 2)
        general_geometry_from_jacobians<2, float><<<blocks, 256>>>(d_ij, d_jw, qpc, Q, d->n_cells, (float *)s.d_metric, (float *)s.d_jxw);
      else if (s.number == B200MF_F64)
        general_geometry_from_jacobians<3, double><<<blocks, 256>>>(d_ij, d_jw, qpc, Q, d->n_cells, (double *)s.d_metric, (double *)s.d_jxw);
      else
        general_geometry_from_jacobians<3, float><<<blocks, 256>>>(d_ij, d_jw, qpc, Q, d->n_cells, (float *)s.d_metric, (float *)s.d_jxw);
      count_launch();
    }
    B200MF_CUDA_CHECK(cudaDeviceSynchronize());
    cudaFree(d_ij);
    cudaFree(d_jw);
  } else {
    set_error("unknown geometry input kind %d", d->geometry);
    return B200MF_ERR_INVALID;
  }
  s.device_bytes += s.geometry_bytes;
  // "atomics or a colouring chosen by measurement" (north star; the reference offers graph
  // colouring or atomics, portable_matrix_free.templates.h:1060-1185): when the setup has both the
  // index-map brick path (memset + atomics on shared dofs) and the bulk brick path
  // (first-toucher-stores, no memset), time a few vmults of each on scratch vectors and keep the
  // faster one.  B200MF_BULK=0/1 forces the choice.
  if (s.bulk.ready || s.colouring.ready) {
    // candidates: 0 = index maps + memset + atomics, 1 = coloured launches (no atomics, no memset),
    // 2 = bulk tables + first-toucher-stores in one launch
    const char *force = std::getenv("B200MF_BRICK_PATH");
    auto select = [&](int path) {
      s.colouring.enabled = path == 1;
      s.bulk.enabled = path == 2;
    };
    if (force != nullptr) {
      select(std::atoi(force));
    } else {
      const size_t bytes = (s.n_owned + s.n_ghost) * number_size(s.number);
      void *va = nullptr, *vb = nullptr;
      int best = 0;
      if (cudaMalloc(&va, bytes) == cudaSuccess && cudaMalloc(&vb, bytes) == cudaSuccess) {
        cudaMemset(va, 0, bytes);
        b200mf_operator op{nullptr, nullptr, 1.0, 0.0};
        cudaEvent_t e0, e1;
        cudaEventCreate(&e0);
        cudaEventCreate(&e1);
        float ms[3] = {0.f, 0.f, 0.f};
        for (int mode = 0; mode < 3 && rc == B200MF_OK; ++mode) {
          if ((mode == 1 && !s.colouring.ready) || (mode == 2 && !s.bulk.ready)) continue;
          select(mode);
          for (int i = 0; i < 2 && rc == B200MF_OK; ++i) rc = vmult_impl(s, op, vb, va, nullptr, nullptr);
          cudaEventRecord(e0, nullptr);
          for (int i = 0; i < 4 && rc == B200MF_OK; ++i) rc = vmult_impl(s, op, vb, va, nullptr, nullptr);
          cudaEventRecord(e1, nullptr);
          cudaEventSynchronize(e1);
          cudaEventElapsedTime(&ms[mode], e0, e1);
          ms[mode] /= 4;
          if (rc == B200MF_OK && ms[mode] < ms[best]) best = mode;
        }
        cudaEventDestroy(e0);
        cudaEventDestroy(e1);
        s.bulk.tuned_ms[0] = ms[0];
        s.colouring.tuned_ms = ms[1];
        s.bulk.tuned_ms[1] = ms[2];
        rc = B200MF_OK;
      } else {
        (void)cudaGetLastError();
      }
      cudaFree(va);
      cudaFree(vb);
      select(best);
    }
  }
  B200MF_CUDA_CHECK(cudaMalloc((void **)&s.d_scratch, 4096 * sizeof(double)));
  B200MF_CUDA_CHECK(cudaMallocHost((void **)&s.h_pinned, 64 * sizeof(double)));
  B200MF_CUDA_CHECK(cudaGetLastError());
#undef TRY
  *out = h;
  h = nullptr; // released: the guard only frees the temporaries now
  return B200MF_OK;
}

int b200mf_setup_destroy(b200mf_setup *h) {
  if (!h) return B200MF_OK;
  Setup *s = &h->impl;
  cudaFree(s->d_l2g); cudaFree(s->d_mask); cudaFree(s->d_geom_id); cudaFree(s->d_geom_table);
  cudaFree(s->d_metric); cudaFree(s->d_jxw); cudaFree(s->d_constrained); cudaFree(s->d_weights);
  cudaFree(s->d_overint_tables);
  cudaFree(s->d_diag_tables);
  cudaFree(s->d_qpoints); cudaFree(s->d_scratch); cudaFree(s->d_brick_map); cudaFree(s->d_zero_list);
  cudaFree(s->colouring.d_list); cudaFree(s->colouring.d_zero); cudaFree(s->d_brick_strided);
  free_bulk(*s);
  if (s->h_pinned) cudaFreeHost(s->h_pinned);
  for (void *w : s->d_work) cudaFree(w);
  for (void *w : s->d_stage) cudaFree(w);
  for (int i = 0; i < 2; ++i) { cudaFree(s->d_pipe_in[i]); cudaFree(s->d_pipe_out[i]); }
  for (int i = 0; i < 3; ++i) {
    if (s->pipe_stream[i]) cudaStreamDestroy(s->pipe_stream[i]);
    for (int j = 0; j < 2; ++j)
      if (s->pipe_event[i][j]) cudaEventDestroy(s->pipe_event[i][j]);
  }
  delete h;
  return B200MF_OK;
}

int b200mf_setup_get_info(const b200mf_setup *h, b200mf_setup_info *info) {
  B200MF_REQUIRE(h && info, "null argument");
  const Setup &s = h->impl;
  info->dim = s.dim; info->degree = s.degree; info->n_q_points_1d = s.n_q_1d; info->number = s.number;
  info->n_cells = s.n_cells; info->n_owned_dofs = s.n_owned; info->n_ghost_dofs = s.n_ghost;
  info->n_constrained_dofs = s.n_constrained; info->cell_kind = s.cell_kind;
  info->n_distinct_geometries = s.n_geom; info->device_bytes = s.device_bytes;
  info->geometry_bytes = s.geometry_bytes; info->index_bytes = s.index_bytes;
  info->n_bricks = s.n_bricks;
  info->cells_per_brick = (uint64_t)s.brick_b * s.brick_b * s.brick_b;
  return B200MF_OK;
}

int b200mf_get_quadrature_points(const b200mf_setup *h, double *out_host) {
  B200MF_REQUIRE(h && out_host, "null argument");
  const Setup &s = h->impl;
  B200MF_REQUIRE(s.has_vertices, "quadrature points need a setup created from Q1 vertices");
  const size_t nvd = (size_t)(1 << s.dim) * s.dim;
  const int Q = s.n_q_1d;
  const uint64_t total_q = s.n_cells * (uint64_t)(s.dim == 2 ? Q * Q : Q * Q * Q);
  double *d_vert = nullptr, *d_qp = nullptr, *d_out = nullptr;
  B200MF_CUDA_CHECK(cudaMalloc((void **)&d_vert, std::max<size_t>(s.n_cells * nvd, 1) * 8));
  B200MF_CUDA_CHECK(cudaMalloc((void **)&d_qp, Q * 8));
  B200MF_CUDA_CHECK(cudaMalloc((void **)&d_out, std::max<size_t>(total_q * s.dim, 1) * 8));
  B200MF_CUDA_CHECK(cudaMemcpy(d_vert, s.h_vertices.data(), s.n_cells * nvd * 8, cudaMemcpyHostToDevice));
  B200MF_CUDA_CHECK(cudaMemcpy(d_qp, s.q_points_1d.data(), Q * 8, cudaMemcpyHostToDevice));
  const unsigned blocks = (unsigned)((total_q + 255) / 256);
  if (total_q) {
    if (s.dim == 2) quadrature_points_kernel<2><<<blocks, 256>>>(d_vert, d_qp, Q, s.n_cells, d_out);
    else            quadrature_points_kernel<3><<<blocks, 256>>>(d_vert, d_qp, Q, s.n_cells, d_out);
    count_launch();
  }
  B200MF_CUDA_CHECK(cudaMemcpy(out_host, d_out, total_q * s.dim * 8, cudaMemcpyDeviceToHost));
  cudaFree(d_vert); cudaFree(d_qp); cudaFree(d_out);
  return B200MF_OK;
}

int b200mf_cell_loop(const b200mf_setup *h, const b200mf_operator *op, void *dst, const void *src,
                     void *stream) {
  B200MF_REQUIRE(h && op && dst && src, "null argument");
  return launch_cell_loop(h->impl, *op, dst, src, 0, h->impl.n_cells, (cudaStream_t)stream, nullptr);
}

int b200mf_cell_loop_range(const b200mf_setup *h, const b200mf_operator *op, void *dst,
                           const void *src, uint64_t cell_begin, uint64_t cell_end, void *stream) {
  B200MF_REQUIRE(h && op && dst && src, "null argument");
  B200MF_REQUIRE(cell_begin <= cell_end && cell_end <= h->impl.n_cells, "bad cell range");
  return launch_cell_loop(h->impl, *op, dst, src, cell_begin, cell_end, (cudaStream_t)stream, nullptr);
}

int b200mf_cell_loop_range_dot(const b200mf_setup *h, const b200mf_operator *op, void *dst,
                               const void *src, uint64_t cell_begin, uint64_t cell_end,
                               double *dot_accum, void *stream) {
  B200MF_REQUIRE(h && op && dst && src, "null argument");
  B200MF_REQUIRE(cell_begin <= cell_end && cell_end <= h->impl.n_cells, "bad cell range");
  return launch_cell_loop(h->impl, *op, dst, src, cell_begin, cell_end, (cudaStream_t)stream, dot_accum);
}

int b200mf_debug_resolve_hanging_nodes(int dim, int degree, int number, uint16_t constraint_mask,
                                       int transpose, double *values_host) {
  B200MF_REQUIRE(values_host && (dim == 2 || dim == 3) && degree >= 1 && degree <= 8, "bad argument");
  B200MF_REQUIRE(number == B200MF_F64 || number == B200MF_F32, "bad number type");
  switch (degree + 1) {
    case 2: return debug_resolve_n2(dim, number, constraint_mask, transpose, values_host);
    case 3: return debug_resolve_n3(dim, number, constraint_mask, transpose, values_host);
    case 4: return debug_resolve_n4(dim, number, constraint_mask, transpose, values_host);
    case 5: return debug_resolve_n5(dim, number, constraint_mask, transpose, values_host);
    case 6: return debug_resolve_n6(dim, number, constraint_mask, transpose, values_host);
    case 7: return debug_resolve_n7(dim, number, constraint_mask, transpose, values_host);
    case 8: return debug_resolve_n8(dim, number, constraint_mask, transpose, values_host);
    default: return debug_resolve_n9(dim, number, constraint_mask, transpose, values_host);
  }
}

int b200mf_brick_probe(const b200mf_setup_desc *d, uint64_t *n_bricks, uint64_t *cells_per_brick,
                       uint64_t *n_complete_dofs) {
  B200MF_REQUIRE(d && n_bricks && cells_per_brick && n_complete_dofs, "null argument");
  B200MF_REQUIRE(d->dim == 3 && d->degree >= 1 && d->degree <= 8 && d->local_to_global, "bad descriptor");
  Setup s;
  s.dim = d->dim; s.degree = d->degree; s.n = d->degree + 1; s.number = d->number;
  s.n_cells = d->n_cells; s.n_owned = d->n_owned_dofs; s.n_ghost = d->n_ghost_dofs;
  s.cell_kind = B200MF_CELLS_CARTESIAN; s.n_geom = 1; // the probe looks at the index lists only
  if (d->constraint_mask)
    for (uint64_t c = 0; c < d->n_cells; ++c)
      if (d->constraint_mask[c]) s.any_mask = true;
  *n_complete_dofs = 0;
  int rc = build_bricks(*d, s, false, n_complete_dofs);
  *n_bricks = s.n_bricks;
  *cells_per_brick = (uint64_t)s.brick_b * s.brick_b * s.brick_b;
  return rc;
}

int b200mf_bulk_probe(const b200mf_setup_desc *d, b200mf_bulk_info *info) {
  B200MF_REQUIRE(d && info, "null argument");
  B200MF_REQUIRE(d->dim == 3 && d->degree >= 1 && d->degree <= 8 && d->local_to_global, "bad descriptor");
  std::memset(info, 0, sizeof(*info));
  Setup s;
  s.dim = d->dim; s.degree = d->degree; s.n = d->degree + 1; s.number = d->number;
  s.n_cells = d->n_cells; s.n_owned = d->n_owned_dofs; s.n_ghost = d->n_ghost_dofs;
  s.cell_kind = B200MF_CELLS_CARTESIAN; s.n_geom = 1;
  if (d->constraint_mask)
    for (uint64_t c = 0; c < d->n_cells; ++c)
      if (d->constraint_mask[c]) s.any_mask = true;
  BulkStats st{};
  uint64_t nc = 0;
  int rc = build_bricks(*d, s, false, &nc, &st);
  if (rc != B200MF_OK) return rc;
  info->n_bricks = st.n_bricks; info->n_patterns = st.n_patterns; info->n_own = st.n_own;
  info->n_first_scalar = st.n_first_scalar; info->n_later = st.n_later; info->n_zero = st.n_zero;
  info->n_general_cells = st.n_general_cells; info->n_boundary_bricks = st.n_boundary_bricks;
  info->usable = s.bulk.ready ? 1 : 0;
  return B200MF_OK;
}

int b200mf_setup_enable_bulk(b200mf_setup *h, int enable) {
  B200MF_REQUIRE(h, "null argument");
  h->impl.bulk.enabled = enable != 0;
  h->impl.colouring.enabled = false;
  return h->impl.bulk.ready ? 1 : 0;
}

int b200mf_setup_enable_strided(b200mf_setup *h, int enable) {
  B200MF_REQUIRE(h, "null argument");
  h->impl.strided_enabled = enable != 0;
  return h->impl.d_brick_strided != nullptr ? 1 : 0;
}

int b200mf_setup_select_brick_path(b200mf_setup *h, int path) {
  B200MF_REQUIRE(h && path >= 0 && path <= 2, "bad argument");
  Setup &s = h->impl;
  if ((path == 1 && !s.colouring.ready) || (path == 2 && !s.bulk.ready)) return -1;
  s.colouring.enabled = path == 1;
  s.bulk.enabled = path == 2;
  return path;
}

int b200mf_setup_get_bulk_info(const b200mf_setup *h, b200mf_bulk_info *info) {
  B200MF_REQUIRE(h && info, "null argument");
  const BulkStats &st = h->impl.bulk.stats;
  info->n_bricks = st.n_bricks; info->n_patterns = st.n_patterns; info->n_own = st.n_own;
  info->n_first_scalar = st.n_first_scalar; info->n_later = st.n_later; info->n_zero = st.n_zero;
  info->n_general_cells = st.n_general_cells; info->n_boundary_bricks = st.n_boundary_bricks;
  info->usable = h->impl.bulk.ready ? 1 : 0;
  info->enabled = h->impl.bulk.ready && h->impl.bulk.enabled ? 1 : 0;
  info->tuned_ms_index_map = h->impl.bulk.tuned_ms[0];
  info->tuned_ms_bulk = h->impl.bulk.tuned_ms[1];
  info->tuned_ms_coloured = h->impl.colouring.tuned_ms;
  info->n_colours = h->impl.colouring.ready ? h->impl.colouring.n_colours : 0;
  info->n_coloured_launches = (int)h->impl.colouring.launches.size();
  info->n_zero_coloured = h->impl.colouring.n_zero;
  info->strided = (h->impl.d_brick_strided != nullptr && h->impl.strided_enabled) ? 1 : 0;
  info->path = (h->impl.bulk.ready && h->impl.bulk.enabled) ? 2 : ((h->impl.colouring.ready && h->impl.colouring.enabled) ? 1 : 0);
  return B200MF_OK;
}

int b200mf_vmult_prepare(const b200mf_setup *h, const b200mf_operator *op, void *dst, void *stream) {
  B200MF_REQUIRE(h && op && dst, "null argument");
  return vmult_prepare_impl(h->impl, *op, dst, (cudaStream_t)stream);
}

int b200mf_vmult_range(const b200mf_setup *h, const b200mf_operator *op, void *dst,
                       const void *src, uint64_t cell_begin, uint64_t cell_end,
                       double *dot_accum, void *stream) {
  B200MF_REQUIRE(h && op && dst && src, "null argument");
  B200MF_REQUIRE(cell_begin <= cell_end && cell_end <= h->impl.n_cells, "bad cell range");
  return launch_cell_loop(h->impl, *op, dst, src, cell_begin, cell_end, (cudaStream_t)stream,
                          dot_accum, true);
}

int b200mf_copy_constrained_values_dot(const b200mf_setup *h, void *dst, const void *src,
                                       double *dot_accum, void *stream) {
  B200MF_REQUIRE(h && dst && src, "null argument");
  return copy_constrained_impl(h->impl, dst, src, (cudaStream_t)stream, dot_accum);
}

int b200mf_copy_constrained_values(const b200mf_setup *h, void *dst, const void *src,
                                   void *stream) {
  B200MF_REQUIRE(h && dst && src, "null argument");
  return copy_constrained_impl(h->impl, dst, src, (cudaStream_t)stream, nullptr);
}

int b200mf_set_constrained_values(const b200mf_setup *h, void *dst, double value, void *stream) {
  B200MF_REQUIRE(h && dst, "null argument");
  return set_constrained_impl(h->impl, dst, value, (cudaStream_t)stream);
}

int b200mf_vmult(const b200mf_setup *h, const b200mf_operator *op, void *dst, const void *src,
                 void *stream) {
  B200MF_REQUIRE(h && op && dst && src, "null argument");
  return vmult_impl(h->impl, *op, dst, src, (cudaStream_t)stream, nullptr);
}

int b200mf_compute_diagonal(const b200mf_setup *h, const b200mf_operator *op, void *diag,
                            void *stream) {
  B200MF_REQUIRE(h && op && diag, "null argument");
  const Setup &s = h->impl;
  cudaStream_t st = (cudaStream_t)stream;
  B200MF_CUDA_CHECK(cudaMemsetAsync(diag, 0, (s.n_owned + s.n_ghost) * number_size(s.number), st));
  int rc = launch_compute_diagonal(s, *op, diag, st);
  if (rc != B200MF_OK) return rc;
  return b200mf_set_constrained_values(h, diag, 1.0, stream);
}

int b200mf_vmult_host(const b200mf_setup *h, const b200mf_operator *op, void *dst_host,
                      const void *src_host) {
  B200MF_REQUIRE(h && op && dst_host && src_host, "null argument");
  Setup &s = const_cast<Setup &>(h->impl);
  const size_t bytes = (s.n_owned + s.n_ghost) * number_size(s.number);
  for (int i = 0; i < 2; ++i)
    if (!s.d_stage[i]) B200MF_CUDA_CHECK(cudaMalloc(&s.d_stage[i], std::max<size_t>(bytes, 1)));
  // src_host / dst_host hold the n_owned locally owned entries (like the batch and CG entry points)
  const size_t owned_bytes = s.n_owned * number_size(s.number);
  B200MF_CUDA_CHECK(cudaMemcpyAsync(s.d_stage[0], src_host, owned_bytes, cudaMemcpyHostToDevice, 0));
  if (bytes > owned_bytes)
    B200MF_CUDA_CHECK(cudaMemsetAsync(static_cast<char *>(s.d_stage[0]) + owned_bytes, 0, bytes - owned_bytes, 0));
  int rc = b200mf_vmult(h, op, s.d_stage[1], s.d_stage[0], nullptr);
  if (rc != B200MF_OK) return rc;
  B200MF_CUDA_CHECK(cudaMemcpyAsync(dst_host, s.d_stage[1], s.n_owned * number_size(s.number),
                                    cudaMemcpyDeviceToHost, 0));
  B200MF_CUDA_CHECK(cudaStreamSynchronize(0));
  return B200MF_OK;
}

int b200mf_vmult_host_batch(const b200mf_setup *h, const b200mf_operator *op, int n_vectors,
                            void *const *dst_host, const void *const *src_host) {
  B200MF_REQUIRE(h && op && n_vectors >= 0 && (n_vectors == 0 || (dst_host && src_host)), "null argument");
  Setup &s = const_cast<Setup &>(h->impl);
  const size_t bytes = (s.n_owned + s.n_ghost) * number_size(s.number);
  const size_t owned_bytes = s.n_owned * number_size(s.number);
  for (int i = 0; i < 2; ++i) {
    if (!s.d_pipe_in[i]) B200MF_CUDA_CHECK(cudaMalloc(&s.d_pipe_in[i], std::max<size_t>(bytes, 1)));
    if (!s.d_pipe_out[i]) B200MF_CUDA_CHECK(cudaMalloc(&s.d_pipe_out[i], std::max<size_t>(bytes, 1)));
  }
  for (int i = 0; i < 3; ++i) {
    if (!s.pipe_stream[i]) B200MF_CUDA_CHECK(cudaStreamCreateWithFlags(&s.pipe_stream[i], cudaStreamNonBlocking));
    for (int j = 0; j < 2; ++j)
      if (!s.pipe_event[i][j])
        B200MF_CUDA_CHECK(cudaEventCreateWithFlags(&s.pipe_event[i][j], cudaEventDisableTiming));
  }
  cudaStream_t s_in = s.pipe_stream[0], s_op = s.pipe_stream[1], s_out = s.pipe_stream[2];
  if (s.n_ghost) { // ghost sections of the inputs are read by the cell loop: zero once
    for (int i = 0; i < 2; ++i)
      B200MF_CUDA_CHECK(cudaMemsetAsync(static_cast<char *>(s.d_pipe_in[i]) + owned_bytes, 0, bytes - owned_bytes, s_in));
  }
  // three-stage pipeline over two slots: upload of vector k+1 and download of vector k-1 (both
  // directions of the link at once) run while vector k is multiplied
  for (int k = 0; k < n_vectors; ++k) {
    const int slot = k & 1;
    B200MF_REQUIRE(dst_host[k] && src_host[k], "null vector in batch");
    B200MF_CUDA_CHECK(cudaStreamWaitEvent(s_in, s.pipe_event[1][slot], 0));   // input slot consumed (k-2)
    B200MF_CUDA_CHECK(cudaMemcpyAsync(s.d_pipe_in[slot], src_host[k], owned_bytes, cudaMemcpyHostToDevice, s_in));
    B200MF_CUDA_CHECK(cudaEventRecord(s.pipe_event[0][slot], s_in));
    B200MF_CUDA_CHECK(cudaStreamWaitEvent(s_op, s.pipe_event[0][slot], 0));   // input arrived
    B200MF_CUDA_CHECK(cudaStreamWaitEvent(s_op, s.pipe_event[2][slot], 0));   // output slot downloaded (k-2)
    int rc = vmult_impl(s, *op, s.d_pipe_out[slot], s.d_pipe_in[slot], s_op, nullptr);
    if (rc != B200MF_OK) return rc;
    B200MF_CUDA_CHECK(cudaEventRecord(s.pipe_event[1][slot], s_op));
    B200MF_CUDA_CHECK(cudaStreamWaitEvent(s_out, s.pipe_event[1][slot], 0));
    B200MF_CUDA_CHECK(cudaMemcpyAsync(dst_host[k], s.d_pipe_out[slot], owned_bytes, cudaMemcpyDeviceToHost, s_out));
    B200MF_CUDA_CHECK(cudaEventRecord(s.pipe_event[2][slot], s_out));
  }
  B200MF_CUDA_CHECK(cudaStreamSynchronize(s_out));
  B200MF_CUDA_CHECK(cudaStreamSynchronize(s_op));
  B200MF_CUDA_CHECK(cudaStreamSynchronize(s_in));
  return B200MF_OK;
}

} // extern "C"
