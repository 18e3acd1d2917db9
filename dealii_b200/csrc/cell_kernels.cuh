// Generic fused cell kernel (all degrees 1..8, dim 2/3, FP64/FP32, all geometry kinds):
//   gather -> [hanging-node interpolation] -> sum-factorised evaluate (collocation route)
//   -> quadrature-point operator -> integrate -> [transposed interpolation] -> atomic scatter.
//
// Replaces the Kokkos team kernel of the reference:
//   ApplyKernel::operator()        matrix_free/portable_matrix_free.templates.h:498-528
//   FEEvaluation::read_dof_values / evaluate / integrate / distribute_local_to_global
//                                  matrix_free/portable_fe_evaluation.h:363-535
//   EvaluatorTensorProduct::apply  matrix_free/portable_tensor_product_kernels.h:340-392
// Design: one thread owns one 1D line of a cell per sweep, loads it into
// registers, applies the (p+1)x(p+1) matrix in even-odd form with coefficients taken
// straight from the constant bank (kernel parameter space), and writes it back to shared
// memory; several cells share one CTA so that 128/256 threads are busy for every degree.
#pragma once
#include "internal.h"
#include "vector_ops.cuh"

namespace b200mf {

constexpr int ipow(int b, int e) { return e == 0 ? 1 : b * ipow(b, e - 1); }

template <int dim, int n>
struct BlockCfg {
  static constexpr int lines = ipow(n, dim - 1);
  static constexpr int target = lines >= 36 ? 256 : 128;
  static constexpr int cells = target / lines;                       // cells per CTA
  static constexpr int threads = ((cells * lines + 31) / 32) * 32;
  static constexpr int npc = ipow(n, dim);
  // values + dim gradient components; odd stride (in elements) between cells so that the
  // per-cell arrays of the cells sharing a warp start in different banks
  static constexpr int cell_stride = ((1 + dim) * npc) | 1;
  static constexpr int smem_elems = cells * cell_stride;
};

// out[q] = sum_i M[i][q] in[i], even-odd form; sym = +1 (shape values) / -1 (derivative)
template <typename Number, int n, int sym>
__device__ __forceinline__ void apply_eo(const EoMatrix<Number, n> &M, const Number (&in)[n],
                                         Number (&out)[n]) {
  constexpr int h = n / 2, hq = (n + 1) / 2;
  Number xe[h > 0 ? h : 1], xo[h > 0 ? h : 1];
#pragma unroll
  for (int i = 0; i < h; ++i) {
    xe[i] = in[i] + in[n - 1 - i];
    xo[i] = in[i] - in[n - 1 - i];
  }
#pragma unroll
  for (int q = 0; q < hq; ++q) {
    // accumulators start from a product (no zero-initialised registers)
    const Number *xa = (sym == 1) ? xe : xo; // operand of the E block
    const Number *xb = (sym == 1) ? xo : xe; // operand of the O block
    Number X = M.E[q] * xa[0];
#pragma unroll
    for (int i = 1; i < h; ++i) X += M.E[i * hq + q] * xa[i];
    if (sym == 1 && n % 2 == 1) X += M.mid[q] * in[h];
    if (q < h) {
      Number Y = M.O[q] * xb[0];
#pragma unroll
      for (int i = 1; i < h; ++i) Y += M.O[i * h + q] * xb[i];
      if (sym != 1 && n % 2 == 1) Y += M.mid[q] * in[h];
      out[q] = X + Y;
      out[n - 1 - q] = X - Y;
    } else {
      out[q] = X;
    }
  }
}

// base offset and stride of 1D line number `line` in direction `dir`
template <int dim, int n, int dir>
__device__ __forceinline__ void line_geometry(int line, int &base, int &stride) {
  if (dim == 2) {
    if (dir == 0) { base = n * line; stride = 1; }
    else          { base = line;     stride = n; }
  } else {
    const int a = line % n, b = line / n;
    if (dir == 0)      { base = n * a + n * n * b; stride = 1; }
    else if (dir == 1) { base = a + n * n * b;     stride = n; }
    else               { base = a + n * b;         stride = n * n; }
  }
}

template <int dim, int n, int dir, int sym, bool add, typename Number>
__device__ __forceinline__ void sweep(const EoMatrix<Number, n> &M, const Number *in, Number *out,
                                      int line) {
  int base, stride;
  line_geometry<dim, n, dir>(line, base, stride);
  Number r[n], o[n];
#pragma unroll
  for (int k = 0; k < n; ++k) r[k] = in[base + k * stride];
  apply_eo<Number, n, sym>(M, r, o);
#pragma unroll
  for (int k = 0; k < n; ++k) {
    if (add) out[base + k * stride] += o[k];
    else     out[base + k * stride] = o[k];
  }
}

template <int dim, int n, typename Number, int KIND>
struct CellKernelParams {
  ShapeData<Number, n> shape;
  const uint32_t *l2g;
  const uint16_t *mask;
  const uint32_t *geom_id;
  const Number *geom_table;
  const Number *metric;
  const Number *jxw;
  const Number *weights; // subface interpolation matrix
  const Number *src;
  Number *dst;
  OperatorArgs<Number> op;
  unsigned long long cell_begin, cell_end;
  // optional fused reduction: *dot_accum += sum_cells u_cell . (A_cell u_cell)  (= src . A src,
  // the p.Ap of CG, lac/solver_cg.h:739) accumulated in the scatter epilogue
  double *dot_accum;
};

constexpr int n_sym(int dim) { return dim * (dim + 1) / 2; }

// Quadrature-point operator on the points of one line in the last direction.
// get_gradient/submit_gradient/get_value/submit_value of
// matrix_free/portable_fe_evaluation.h:544,653-745 with the metric merged to
// JxW * J^-1 J^-T (and stored once per distinct affine cell).
template <int dim, int n, typename Number, int KIND>
__device__ __forceinline__ void quadrature_point_operation(
    const CellKernelParams<dim, n, Number, KIND> &p, Number *U, Number *G, unsigned long long cell,
    int line) {
  constexpr int npc = ipow(n, dim);
  constexpr int NS = n_sym(dim);
  int base, stride;
  line_geometry<dim, n, dim - 1>(line, base, stride);
  // weight of the line's fixed coordinates
  Number wline;
  if (dim == 2) wline = p.shape.w[line];
  else          wline = p.shape.w[line % n] * p.shape.w[line / n];

  Number m[NS > 0 ? NS : 1];
  Number det = Number(1);
  if (KIND != B200MF_CELLS_GENERAL) {
    const unsigned gi = p.geom_id ? p.geom_id[cell] : 0u;
    if (KIND == B200MF_CELLS_CARTESIAN) {
      const Number *t = p.geom_table + gi * (dim + 1);
#pragma unroll
      for (int d = 0; d < dim; ++d) m[d] = t[d];
      det = t[dim];
    } else {
      const Number *t = p.geom_table + gi * (NS + 1);
#pragma unroll
      for (int d = 0; d < NS; ++d) m[d] = t[d];
      det = t[NS];
    }
  }
  const bool has_mass = p.op.has_mass;
#pragma unroll
  for (int k = 0; k < n; ++k) {
    const int q = base + k * stride;
    const unsigned long long gq = cell * npc + q;
    Number cg = p.op.grad_const;
    if (p.op.grad_coef) cg *= p.op.grad_coef[gq];
    Number g[dim];
#pragma unroll
    for (int d = 0; d < dim; ++d) g[d] = G[d * npc + q];
    Number jxw;
    if (KIND == B200MF_CELLS_GENERAL) {
#pragma unroll
      for (int d = 0; d < NS; ++d) m[d] = p.metric[metric_offset<dim>(n, cell, d, q)];
      jxw = has_mass ? p.jxw[gq] : Number(0);
    } else {
      const Number wq = wline * p.shape.w[k];
      cg *= wq;
      jxw = det * wq;
    }
    if constexpr (KIND == B200MF_CELLS_CARTESIAN) {
#pragma unroll
      for (int d = 0; d < dim; ++d) G[d * npc + q] = g[d] * (m[d] * cg);
    } else if constexpr (dim == 2) {
      G[q] = cg * (m[0] * g[0] + m[1] * g[1]);
      G[npc + q] = cg * (m[1] * g[0] + m[2] * g[1]);
    } else {
      G[q] = cg * (m[0] * g[0] + m[1] * g[1] + m[2] * g[2]);
      G[npc + q] = cg * (m[1] * g[0] + m[3] * g[1] + m[4] * g[2]);
      G[2 * npc + q] = cg * (m[2] * g[0] + m[4] * g[1] + m[5] * g[2]);
    }
    if (has_mass) {
      Number cm = p.op.mass_const;
      if (p.op.mass_coef) cm += p.op.mass_coef[gq];
      U[q] = U[q] * (cm * jxw);
    } else {
      U[q] = Number(0);
    }
  }
}

// Hanging-node interpolation of one cell held in shared memory (in place).
// Same algebra as Portable::internal::resolve_hanging_nodes
// (matrix_free/portable_hanging_nodes_internal.h:124-459) driven by the ConstraintKinds
// bit mask (matrix_free/hanging_nodes_internal.h:40-60):
//   bit 0..2  subcell_x,y,z   bit 3..5  face_x,y,z   bit 6..8  edge_x,y,z
// One thread handles one line per pass; a pass interpolates along one direction.
template <int dim, int n, int dir, bool transpose, typename Number>
__device__ __forceinline__ void hanging_node_pass(const Number *W, unsigned mask, Number *U,
                                                  int line) {
  constexpr int p = n - 1;
  // coordinates orthogonal to `dir`
  int c1, c2 = 0;
  constexpr int d1 = (dim == 2) ? 1 - dir : (dir + 1) % 3;
  constexpr int d2 = (dim == 2) ? 0 : (dir + 2) % 3;
  if (dim == 2) {
    c1 = line;
  } else {
    const int a = line % n, b = line / n; // the two remaining coordinates, lower direction first
    // dir 0: (a,b)=(y,z); dir 1: (a,b)=(x,z); dir 2: (a,b)=(x,y)
    if (dir == 0)      { c1 = a; c2 = b; }  // d1 = y, d2 = z
    else if (dir == 1) { c1 = b; c2 = a; }  // d1 = z, d2 = x
    else               { c1 = a; c2 = b; }  // d1 = x, d2 = y
  }
  const bool on1 = (mask & (1u << d1)) ? (c1 == 0) : (c1 == p);
  bool constrained = (mask & (8u << d1)) && on1;
  if (dim == 3) {
    const bool on2 = (mask & (1u << d2)) ? (c2 == 0) : (c2 == p);
    constrained = constrained || ((mask & (8u << d2)) && on2) ||
                  ((mask & (64u << dir)) && on1 && on2);
  }
  if (!constrained) return;
  int base, stride;
  line_geometry<dim, n, dir>(line, base, stride);
  Number v[n], o[n];
#pragma unroll
  for (int k = 0; k < n; ++k) v[k] = U[base + k * stride];
  const bool first_child = (mask & (1u << dir)) != 0;
#pragma unroll
  for (int j = 0; j < n; ++j) {
    Number sum = Number(0);
#pragma unroll
    for (int i = 0; i < n; ++i) {
      int r = transpose ? i : j, c = transpose ? j : i;
      if (!first_child) { r = p - r; c = p - c; }
      sum += __ldg(W + r * n + c) * v[i];
    }
    o[j] = sum;
  }
#pragma unroll
  for (int k = 0; k < n; ++k) U[base + k * stride] = o[k];
}

template <int dim, int n, typename Number, bool transpose>
__device__ __noinline__ void resolve_hanging_nodes_cell(const Number *W, unsigned mask, Number *U,
                                                        int line, bool active) {
  if (active && mask) hanging_node_pass<dim, n, 0, transpose>(W, mask, U, line);
  __syncthreads();
  if (active && mask) hanging_node_pass<dim, n, 1, transpose>(W, mask, U, line);
  __syncthreads();
  if (dim == 3) {
    if (active && mask) hanging_node_pass<dim, n, (dim == 3 ? 2 : 0), transpose>(W, mask, U, line);
    __syncthreads();
  }
}

// Evaluate the cell operator on the cells of this CTA held in shared memory.
// On entry U = local dof values, on exit U = local result. Contains __syncthreads().
template <int dim, int n, typename Number, int KIND>
__device__ __forceinline__ void evaluate_cells(const CellKernelParams<dim, n, Number, KIND> &p,
                                               Number *U, Number *G, unsigned long long cell,
                                               int line, bool active) {
  constexpr int npc = ipow(n, dim);
  constexpr int LAST = dim - 1;
  const ShapeData<Number, n> &sh = p.shape;

  // ---- values at quadrature points, all but the last direction
  if (active) sweep<dim, n, 0, 1, false>(sh.S, U, U, line);
  __syncthreads();
  if (dim == 3) {
    if (active) sweep<dim, n, 1, 1, false>(sh.S, U, U, line);
    __syncthreads();
  }
  // ---- last direction: values, then derivative along the same line from registers
  if (active) {
    int base, stride;
    line_geometry<dim, n, LAST>(line, base, stride);
    Number r[n], o[n], g[n];
#pragma unroll
    for (int k = 0; k < n; ++k) r[k] = U[base + k * stride];
    apply_eo<Number, n, 1>(sh.S, r, o);
    apply_eo<Number, n, -1>(sh.D, o, g);
#pragma unroll
    for (int k = 0; k < n; ++k) {
      U[base + k * stride] = o[k];
      G[LAST * npc + base + k * stride] = g[k];
    }
  }
  __syncthreads();
  // ---- remaining reference derivatives
  if (active) {
    sweep<dim, n, 0, -1, false>(sh.D, U, G, line);
    if (dim == 3) sweep<dim, n, 1, -1, false>(sh.D, U, G + npc, line);
  }
  __syncthreads();
  // ---- quadrature point operation
  if (active) quadrature_point_operation<dim, n, Number, KIND>(p, U, G, cell, line);
  __syncthreads();
  // ---- integrate: transposed derivatives summed into the value array
  if (active) sweep<dim, n, 0, -1, true>(sh.Dt, G, U, line);
  __syncthreads();
  if (dim == 3) {
    if (active) sweep<dim, n, 1, -1, true>(sh.Dt, G + npc, U, line);
    __syncthreads();
  }
  if (active) {
    int base, stride;
    line_geometry<dim, n, LAST>(line, base, stride);
    Number r[n], o[n], t[n];
#pragma unroll
    for (int k = 0; k < n; ++k) r[k] = G[LAST * npc + base + k * stride];
    apply_eo<Number, n, -1>(sh.Dt, r, o);
#pragma unroll
    for (int k = 0; k < n; ++k) o[k] += U[base + k * stride];
    apply_eo<Number, n, 1>(sh.St, o, t);
#pragma unroll
    for (int k = 0; k < n; ++k) U[base + k * stride] = t[k];
  }
  __syncthreads();
  if (dim == 3) {
    if (active) sweep<dim, n, 1, 1, false>(sh.St, U, U, line);
    __syncthreads();
  }
  if (active) sweep<dim, n, 0, 1, false>(sh.St, U, U, line);
  __syncthreads();
}

template <int dim, int n, typename Number>
__device__ __forceinline__ void atomic_add_number(Number *addr, Number v) {
  atomicAdd(addr, v);
}

template <int dim, int n, typename Number, int KIND>
__global__ void __launch_bounds__(BlockCfg<dim, n>::threads)
cell_loop_kernel(const __grid_constant__ CellKernelParams<dim, n, Number, KIND> p) {
  using Cfg = BlockCfg<dim, n>;
  constexpr int npc = Cfg::npc;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  Number *sm = reinterpret_cast<Number *>(smem_raw);

  const int tid = threadIdx.x;
  const int cib = tid / Cfg::lines;
  const int line = tid - cib * Cfg::lines;
  const unsigned long long cell0 = p.cell_begin + (unsigned long long)blockIdx.x * Cfg::cells;
  const unsigned long long remaining = p.cell_end - cell0;
  const int ncell = remaining < (unsigned long long)Cfg::cells ? (int)remaining : Cfg::cells;
  const bool active = cib < ncell;
  Number *U = sm + (cib < Cfg::cells ? cib : 0) * Cfg::cell_stride;
  Number *G = U + npc;

  // ---- gather (read_dof_values): coalesced over the lexicographic index list
  const uint32_t *l2g = p.l2g + cell0 * npc;
  constexpr int KEEP = (Cfg::cells * npc + Cfg::threads - 1) / Cfg::threads;
  Number ukeep[KEEP];
#pragma unroll
  for (int j = 0; j < KEEP; ++j) {
    const int i = tid + j * Cfg::threads;
    ukeep[j] = Number(0);
    if (i < ncell * npc) {
      const int c = i / npc, k = i - c * npc;
      const uint32_t idx = l2g[i];
      const Number v = (idx & B200MF_L2G_CONSTRAINED) ? Number(0) : __ldg(p.src + idx);
      sm[c * Cfg::cell_stride + k] = v;
      ukeep[j] = v;
    }
  }
  __syncthreads();

  unsigned mask = 0;
  if (p.mask != nullptr) {
    mask = active ? p.mask[cell0 + cib] : 0u;
    if (__syncthreads_or(mask != 0))
      resolve_hanging_nodes_cell<dim, n, Number, false>(p.weights, mask, U, line, active);
  }

  evaluate_cells<dim, n, Number, KIND>(p, U, G, cell0 + cib, line, active);

  if (p.mask != nullptr) {
    if (__syncthreads_or(mask != 0))
      resolve_hanging_nodes_cell<dim, n, Number, true>(p.weights, mask, U, line, active);
  }

  // ---- scatter (distribute_local_to_global): FP atomics resolve shared dofs
  double dot = 0.0;
#pragma unroll
  for (int j = 0; j < KEEP; ++j) {
    const int i = tid + j * Cfg::threads;
    if (i < ncell * npc) {
      const int c = i / npc, k = i - c * npc;
      const uint32_t idx = l2g[i];
      if (!(idx & B200MF_L2G_CONSTRAINED)) {
        const Number r = sm[c * Cfg::cell_stride + k];
        atomicAdd(p.dst + idx, r);
        dot += double(ukeep[j]) * double(r);
      }
    }
  }
  if (p.dot_accum != nullptr) {
    dot = block_sum(dot);
    if (tid == 0) atomicAdd(p.dot_accum, dot);
  }
}

// MatrixFreeTools::compute_diagonal (matrix_free/tools.h:1392-1569): apply the cell operator
// to every local unit vector and keep the diagonal entry; shared dofs summed by atomics.
template <int dim, int n, typename Number, int KIND>
__global__ void __launch_bounds__(BlockCfg<dim, n>::threads)
cell_diagonal_kernel(const __grid_constant__ CellKernelParams<dim, n, Number, KIND> p) {
  using Cfg = BlockCfg<dim, n>;
  constexpr int npc = Cfg::npc;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  Number *sm = reinterpret_cast<Number *>(smem_raw);
  Number *diag_sm = sm + Cfg::smem_elems; // [cells][npc]

  const int tid = threadIdx.x;
  const int cib = tid / Cfg::lines;
  const int line = tid - cib * Cfg::lines;
  const unsigned long long cell0 = p.cell_begin + (unsigned long long)blockIdx.x * Cfg::cells;
  const unsigned long long remaining = p.cell_end - cell0;
  const int ncell = remaining < (unsigned long long)Cfg::cells ? (int)remaining : Cfg::cells;
  const bool active = cib < ncell;
  Number *U = sm + (cib < Cfg::cells ? cib : 0) * Cfg::cell_stride;
  Number *G = U + npc;

  unsigned mask = 0;
  if (p.mask != nullptr) mask = active ? p.mask[cell0 + cib] : 0u;
  const bool any_mask = p.mask != nullptr && __syncthreads_or(mask != 0);

  for (int i = 0; i < npc; ++i) {
    for (int k = tid; k < ncell * npc; k += Cfg::threads) {
      const int c = k / npc, j = k - c * npc;
      sm[c * Cfg::cell_stride + j] = (j == i) ? Number(1) : Number(0);
    }
    __syncthreads();
    if (any_mask) resolve_hanging_nodes_cell<dim, n, Number, false>(p.weights, mask, U, line, active);
    evaluate_cells<dim, n, Number, KIND>(p, U, G, cell0 + cib, line, active);
    if (any_mask) resolve_hanging_nodes_cell<dim, n, Number, true>(p.weights, mask, U, line, active);
    if (tid < ncell) diag_sm[tid * npc + i] = sm[tid * Cfg::cell_stride + i];
    __syncthreads();
  }
  const uint32_t *l2g = p.l2g + cell0 * npc;
  for (int i = tid; i < ncell * npc; i += Cfg::threads) {
    const uint32_t idx = l2g[i];
    if (!(idx & B200MF_L2G_CONSTRAINED)) atomicAdd(p.dst + idx, diag_sm[i]);
  }
}

// compute_diagonal by sum factorisation (cells without hanging-node masks).  With
// phi_i(q) = prod_a S[i_a][q_a] and d_d phi_i(q) = (SD)[i_d][q_d] prod_{a != d} S[i_a][q_a],
//   A_ii = sum_q sum_{d,e} M_de(q) d_d phi_i d_e phi_i + m(q) phi_i^2
// is, term by term, the contraction of a coefficient field with a tensor product of the 1D tables
// SS = S.*S, GG = (SD).*(SD), SG = S.*(SD): three sweeps per term, dim(dim+1)/2 + 1 terms, instead
// of the (p+1)^dim operator applications of the reference (matrix_free/tools.h:1392-1569, whose
// result it reproduces to round-off).  tables = [SS | GG | SG], each [i * n + q].
template <int dim, int n, int dir, typename Number>
__device__ __forceinline__ void plain_sweep(const Number *T, Number *A, int line) {
  int base, stride;
  line_geometry<dim, n, dir>(line, base, stride);
  Number r[n], o[n];
#pragma unroll
  for (int k = 0; k < n; ++k) r[k] = A[base + k * stride];
#pragma unroll
  for (int i = 0; i < n; ++i) {
    Number acc = T[i * n] * r[0];
#pragma unroll
    for (int q = 1; q < n; ++q) acc += T[i * n + q] * r[q];
    o[i] = acc;
  }
#pragma unroll
  for (int k = 0; k < n; ++k) A[base + k * stride] = o[k];
}

template <int dim, int n, typename Number, int KIND>
__global__ void __launch_bounds__(BlockCfg<dim, n>::threads)
cell_diagonal_sumfac_kernel(const __grid_constant__ CellKernelParams<dim, n, Number, KIND> p,
                            const Number *__restrict__ tables) {
  using Cfg = BlockCfg<dim, n>;
  constexpr int npc = Cfg::npc, NS = n_sym(dim), n2 = n * n;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  Number *sm = reinterpret_cast<Number *>(smem_raw);
  Number *T = sm + Cfg::smem_elems; // [3][n*n]
  const int tid = threadIdx.x;
  const int cib = tid / Cfg::lines;
  const int line = tid - cib * Cfg::lines;
  const unsigned long long cell0 = p.cell_begin + (unsigned long long)blockIdx.x * Cfg::cells;
  const unsigned long long remaining = p.cell_end - cell0;
  const int ncell = remaining < (unsigned long long)Cfg::cells ? (int)remaining : Cfg::cells;
  const bool active = cib < ncell;
  const unsigned long long cell = cell0 + (active ? cib : 0);
  Number *W = sm + (cib < Cfg::cells ? cib : 0) * Cfg::cell_stride; // work array of this cell
  Number *Dg = W + npc;                                              // its diagonal
  for (int i = tid; i < 3 * n2; i += Cfg::threads) T[i] = tables[i];
  const Number *SS = T, *GG = T + n2, *SG = T + 2 * n2;

  int base, stride;
  line_geometry<dim, n, dim - 1>(line, base, stride);
  Number wline;
  if (dim == 2) wline = p.shape.w[line];
  else          wline = p.shape.w[line % n] * p.shape.w[line / n];
  Number m[NS], det = Number(1);
#pragma unroll
  for (int s = 0; s < NS; ++s) m[s] = Number(0);
  if (KIND != B200MF_CELLS_GENERAL && active) {
    const unsigned gi = p.geom_id ? p.geom_id[cell] : 0u;
    if (KIND == B200MF_CELLS_CARTESIAN) {
      const Number *t = p.geom_table + gi * (dim + 1);
      // diagonal entries of the symmetric storage: 2D (0, 2), 3D (0, 3, 5)
      m[0] = t[0];
      m[dim == 2 ? 2 : 3] = t[1];
      if (dim == 3) m[5] = t[2];
      det = t[dim];
    } else {
      const Number *t = p.geom_table + gi * (NS + 1);
#pragma unroll
      for (int s = 0; s < NS; ++s) m[s] = t[s];
      det = t[NS];
    }
  }
  if (active) {
#pragma unroll
    for (int k = 0; k < n; ++k) Dg[base + k * stride] = Number(0);
  }
  __syncthreads();

  // terms: the NS entries (d <= e) of the metric, then the mass term
  int s = 0;
#pragma unroll
  for (int d = 0; d < dim; ++d)
#pragma unroll
    for (int e = d; e <= dim; ++e) {
      const bool mass = (e == dim);
      if (mass && d != dim - 1) continue; // one mass term, attached to the last d
      const bool skip = mass ? !p.op.has_mass : (KIND == B200MF_CELLS_CARTESIAN && d != e);
      const int ss = s;
      if (!mass) ++s;
      if (skip) continue;
      if (active) {
#pragma unroll
        for (int k = 0; k < n; ++k) {
          const int q = base + k * stride;
          const unsigned long long gq = cell * npc + q;
          Number c;
          if (mass) {
            Number cm = p.op.mass_const;
            if (p.op.mass_coef) cm += p.op.mass_coef[gq];
            const Number jxw = (KIND == B200MF_CELLS_GENERAL) ? p.jxw[gq] : det * wline * p.shape.w[k];
            c = cm * jxw;
          } else {
            Number cg = p.op.grad_const * (d == e ? Number(1) : Number(2));
            if (p.op.grad_coef) cg *= p.op.grad_coef[gq];
            const Number mq = (KIND == B200MF_CELLS_GENERAL) ? p.metric[metric_offset<dim>(n, cell, ss, q)]
                                                            : m[ss] * wline * p.shape.w[k];
            c = cg * mq;
          }
          W[q] = c;
        }
      }
      __syncthreads();
      auto table = [&](int a) -> const Number * {
        if (mass) return SS;
        if (a == d && a == e) return GG;
        if (a == d || a == e) return SG;
        return SS;
      };
      if (active) plain_sweep<dim, n, 0>(table(0), W, line);
      __syncthreads();
      if (active) plain_sweep<dim, n, 1>(table(1), W, line);
      __syncthreads();
      if (dim == 3) {
        if (active) plain_sweep<dim, n, (dim == 3 ? 2 : 0)>(table(2), W, line);
        __syncthreads();
      }
      if (active) {
#pragma unroll
        for (int k = 0; k < n; ++k) Dg[base + k * stride] += W[base + k * stride];
      }
      __syncthreads();
    }
  const uint32_t *l2g = p.l2g + cell0 * npc;
  for (int i = tid; i < ncell * npc; i += Cfg::threads) {
    const int c = i / npc, k = i - c * npc;
    const uint32_t idx = l2g[i];
    if (!(idx & B200MF_L2G_CONSTRAINED)) atomicAdd(p.dst + idx, sm[c * Cfg::cell_stride + npc + k]);
  }
}

} // namespace b200mf
