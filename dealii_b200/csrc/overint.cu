// Cell loop for n_q_points_1d > fe_degree + 1 (over-integration), which Portable::MatrixFree allows
// (AssertThrow(n_q_points_1d >= fe_degree + 1), matrix_free/portable_matrix_free.templates.h:1243) and
// Portable::FEEvaluation<dim, fe_degree, n_q_points_1d> implements with the non-collocation route of
// matrix_free/portable_evaluation_kernels.h:418-503: interpolate the n = p+1 dof values per direction to the
// Q quadrature points, differentiate there (collocation derivative of the Lagrange basis in the Q Gauss
// points), apply the quadrature-point operator, and come back with the transposed matrices.
//
// Hanging-node constraints are resolved on the cell's dof values before and (transposed) after, with the
// routine of the collocation kernels.
//
// This is the correctness path of the engine for that case, not a tuned kernel: one CTA per cell, sizes are
// run-time values, all tensors live in shared memory.  The fast paths (bricks, plane kernel) cover
// n_q_points_1d == fe_degree + 1, which is what the reference's tutorials and benchmarks use.
#include "cell_kernels.cuh"
#include "vector_ops.cuh"

namespace b200mf {

constexpr int kOverThreads = 128;

template <typename Number>
struct OverintParams {
  const uint32_t *l2g;
  const uint16_t *mask;    // ConstraintKinds per cell or null
  const Number *weights;   // subface interpolation matrix [n][n]
  const uint32_t *geom_id;
  const Number *geom_table, *metric, *jxw;
  const Number *tables; // S[n][Q] (dof basis at the quadrature points) | Dq[Q][Q] (l_a'(x_b)) | w[Q]
  const Number *src;
  Number *dst;
  OperatorArgs<Number> op;
  unsigned long long cell_begin, cell_end;
  double *dot_accum;
  int n, Q, kind, diagonal;
};

// out[.., a, ..] (+)= sum_b mat[b * sb + a * sa] * in[.., b, ..] along direction dir; `ext` are the extents
// of `in`, the output has na entries along dir
template <typename Number>
__device__ void contract(Number *out, const Number *in, const Number *mat, int sb, int sa, int dir,
                         const int *ext, int na, bool add) {
  const int nb = ext[dir];
  int oe[3] = {ext[0], ext[1], ext[2]};
  oe[dir] = na;
  const int total = oe[0] * oe[1] * oe[2];
  const int istride = dir == 0 ? 1 : (dir == 1 ? ext[0] : ext[0] * ext[1]);
  for (int o = threadIdx.x; o < total; o += blockDim.x) {
    int c[3] = {o % oe[0], (o / oe[0]) % oe[1], o / (oe[0] * oe[1])};
    const int a = c[dir];
    c[dir] = 0;
    const int base = c[0] + ext[0] * (c[1] + ext[1] * c[2]);
    Number acc = add ? out[o] : Number(0);
    for (int b = 0; b < nb; ++b) acc += mat[b * sb + a * sa] * in[base + b * istride];
    out[o] = acc;
  }
}

// u0 (n^dim, in place of r0 on exit is NOT done: result goes to `res`): the cell operator
template <int dim, typename Number>
__device__ void apply_cell(const OverintParams<Number> &p, unsigned long long cell, const Number *u0, Number *A,
                           Number *B, Number *G, const Number *S, const Number *Dq, const Number *w,
                           Number **result) {
  const int n = p.n, Q = p.Q;
  const int qpc = dim == 2 ? Q * Q : Q * Q * Q;
  constexpr int NS = dim * (dim + 1) / 2;
  int ext[3] = {n, n, dim == 3 ? n : 1};
  // ---- values at the quadrature points
  const Number *in = u0;
  Number *cur = A, *other = B;
  for (int d = 0; d < dim; ++d) {
    contract(cur, in, S, Q, 1, d, ext, Q, false);
    ext[d] = Q;
    __syncthreads();
    in = cur;
    Number *t = cur; cur = other; other = t;
  }
  Number *V = const_cast<Number *>(in); // values, in `other`
  // ---- reference gradients at the quadrature points
  for (int d = 0; d < dim; ++d) contract(G + d * qpc, V, Dq, Q, 1, d, ext, Q, false);
  __syncthreads();
  // ---- quadrature-point operator (get_gradient / submit_gradient / get_value / submit_value)
  for (int q = threadIdx.x; q < qpc; q += blockDim.x) {
    const int qi[3] = {q % Q, (q / Q) % Q, q / (Q * Q)};
    const unsigned long long gq = cell * qpc + q;
    Number cg = p.op.grad_const;
    if (p.op.grad_coef) cg *= p.op.grad_coef[gq];
    Number m[NS], jxw;
    if (p.kind == B200MF_CELLS_GENERAL) {
      for (int s = 0; s < NS; ++s) m[s] = p.metric[metric_offset<dim>(Q, cell, s, q)];
      jxw = p.jxw[gq];
    } else {
      Number wq = w[qi[0]] * w[qi[1]];
      if (dim == 3) wq *= w[qi[2]];
      const unsigned gi = p.geom_id ? p.geom_id[cell] : 0u;
      Number det;
      if (p.kind == B200MF_CELLS_CARTESIAN) {
        const Number *t = p.geom_table + gi * (dim + 1);
        // diagonal metric in the slots of the symmetric one
        for (int s = 0; s < NS; ++s) m[s] = Number(0);
        if (dim == 2) { m[0] = t[0]; m[2] = t[1]; }
        else          { m[0] = t[0]; m[3] = t[1]; m[5] = t[2]; }
        det = t[dim];
      } else {
        const Number *t = p.geom_table + gi * (NS + 1);
        for (int s = 0; s < NS; ++s) m[s] = t[s];
        det = t[NS];
      }
      cg *= wq; // the table's metric already carries det(J)
      jxw = wq * det;
    }
    Number g[3] = {G[q], G[qpc + q], dim == 3 ? G[2 * qpc + q] : Number(0)};
    if (dim == 2) {
      G[q] = cg * (m[0] * g[0] + m[1] * g[1]);
      G[qpc + q] = cg * (m[1] * g[0] + m[2] * g[1]);
    } else {
      G[q] = cg * (m[0] * g[0] + m[1] * g[1] + m[2] * g[2]);
      G[qpc + q] = cg * (m[1] * g[0] + m[3] * g[1] + m[4] * g[2]);
      G[2 * qpc + q] = cg * (m[2] * g[0] + m[4] * g[1] + m[5] * g[2]);
    }
    if (p.op.has_mass) {
      Number cm = p.op.mass_const;
      if (p.op.mass_coef) cm += p.op.mass_coef[gq];
      V[q] = V[q] * (cm * jxw);
    } else {
      V[q] = Number(0);
    }
  }
  __syncthreads();
  // ---- integrate: test with the gradients (Dq transposed), then back to the dof basis (S transposed)
  for (int d = 0; d < dim; ++d) {
    contract(V, G + d * qpc, Dq, 1, Q, d, ext, Q, true);
    __syncthreads();
  }
  in = V;
  cur = (V == A) ? B : A;
  other = V;
  for (int d = 0; d < dim; ++d) {
    contract(cur, in, S, 1, Q, d, ext, n, false);
    ext[d] = n;
    __syncthreads();
    in = cur;
    Number *t = cur; cur = other; other = t;
  }
  *result = const_cast<Number *>(in);
}

// resolve_hanging_nodes on the cell's n^dim values (the routine of cell_kernels.cuh needs n at compile time)
template <int dim, typename Number, bool transpose>
__device__ void resolve_runtime(int n, const Number *W, unsigned mask, Number *U) {
  const int lines = dim == 2 ? n : n * n;
  const int line = threadIdx.x;
  const bool active = line < lines;
  switch (n) {
    case 2: resolve_hanging_nodes_cell<dim, 2, Number, transpose>(W, mask, U, line, active); break;
    case 3: resolve_hanging_nodes_cell<dim, 3, Number, transpose>(W, mask, U, line, active); break;
    case 4: resolve_hanging_nodes_cell<dim, 4, Number, transpose>(W, mask, U, line, active); break;
    case 5: resolve_hanging_nodes_cell<dim, 5, Number, transpose>(W, mask, U, line, active); break;
    case 6: resolve_hanging_nodes_cell<dim, 6, Number, transpose>(W, mask, U, line, active); break;
    case 7: resolve_hanging_nodes_cell<dim, 7, Number, transpose>(W, mask, U, line, active); break;
    case 8: resolve_hanging_nodes_cell<dim, 8, Number, transpose>(W, mask, U, line, active); break;
    default: resolve_hanging_nodes_cell<dim, 9, Number, transpose>(W, mask, U, line, active); break;
  }
}

template <int dim, typename Number>
__global__ void __launch_bounds__(kOverThreads) overint_kernel(const OverintParams<Number> p) {
  extern __shared__ __align__(16) unsigned char over_smem[];
  const int n = p.n, Q = p.Q;
  const int npc = dim == 2 ? n * n : n * n * n, qpc = dim == 2 ? Q * Q : Q * Q * Q;
  Number *S = reinterpret_cast<Number *>(over_smem), *Dq = S + n * Q, *w = Dq + Q * Q;
  Number *u0 = w + Q, *A = u0 + npc, *B = A + qpc, *G = B + qpc, *acc = G + dim * qpc; // acc: npc (diagonal)
  for (int i = threadIdx.x; i < n * Q + Q * Q + Q; i += blockDim.x) S[i] = p.tables[i];
  __syncthreads();
  for (unsigned long long cell = p.cell_begin + blockIdx.x; cell < p.cell_end; cell += gridDim.x) {
    const uint32_t *l2g = p.l2g + cell * npc;
    const unsigned mask = p.mask ? p.mask[cell] : 0u; // uniform over the CTA
    Number *res = nullptr;
    if (!p.diagonal) {
      for (int i = threadIdx.x; i < npc; i += blockDim.x) {
        const uint32_t idx = l2g[i];
        const Number v = (idx & B200MF_L2G_CONSTRAINED) ? Number(0) : p.src[idx];
        u0[i] = v;
        acc[i] = v; // the values as read, for src . (A src)
      }
      __syncthreads();
      if (mask) resolve_runtime<dim, Number, false>(n, p.weights, mask, u0);
      apply_cell<dim, Number>(p, cell, u0, A, B, G, S, Dq, w, &res);
      if (mask) resolve_runtime<dim, Number, true>(n, p.weights, mask, res);
      double dot = 0.0;
      for (int i = threadIdx.x; i < npc; i += blockDim.x) {
        const uint32_t idx = l2g[i];
        if (!(idx & B200MF_L2G_CONSTRAINED)) {
          atomicAdd(p.dst + idx, res[i]);
          dot += double(acc[i]) * double(res[i]);
        }
      }
      if (p.dot_accum != nullptr) {
        dot = block_sum(dot);
        if (threadIdx.x == 0) atomicAdd(p.dot_accum, dot);
      }
      __syncthreads();
    } else {
      // MatrixFreeTools::compute_diagonal: the cell operator on every local unit vector (matrix_free/tools.h)
      for (int j = 0; j < npc; ++j) {
        for (int i = threadIdx.x; i < npc; i += blockDim.x) u0[i] = i == j ? Number(1) : Number(0);
        __syncthreads();
        if (mask) resolve_runtime<dim, Number, false>(n, p.weights, mask, u0);
        apply_cell<dim, Number>(p, cell, u0, A, B, G, S, Dq, w, &res);
        if (mask) resolve_runtime<dim, Number, true>(n, p.weights, mask, res);
        if (threadIdx.x == 0) acc[j] = res[j];
        __syncthreads();
      }
      for (int i = threadIdx.x; i < npc; i += blockDim.x) {
        const uint32_t idx = l2g[i];
        if (!(idx & B200MF_L2G_CONSTRAINED)) atomicAdd(p.dst + idx, acc[i]);
      }
      __syncthreads();
    }
  }
}

template <int dim, typename Number>
static int launch_overint_t(const Setup &s, const b200mf_operator &op, void *dst, const void *src, uint64_t cb,
                            uint64_t ce, cudaStream_t st, double *dot, bool diagonal) {
  OverintParams<Number> p;
  p.l2g = s.d_l2g; p.geom_id = s.d_geom_id;
  p.mask = s.any_mask ? s.d_mask : nullptr; p.weights = (const Number *)s.d_weights;
  p.geom_table = (const Number *)s.d_geom_table; p.metric = (const Number *)s.d_metric; p.jxw = (const Number *)s.d_jxw;
  p.tables = (const Number *)s.d_overint_tables;
  p.src = (const Number *)src; p.dst = (Number *)dst;
  p.op.grad_coef = (const Number *)op.grad_coefficient; p.op.mass_coef = (const Number *)op.mass_coefficient;
  p.op.grad_const = Number(op.grad_constant); p.op.mass_const = Number(op.mass_constant);
  p.op.has_mass = (op.mass_coefficient != nullptr || op.mass_constant != 0.0) ? 1 : 0;
  p.cell_begin = cb; p.cell_end = ce; p.dot_accum = dot;
  p.n = s.n; p.Q = s.n_q_1d; p.kind = s.cell_kind; p.diagonal = diagonal ? 1 : 0;
  const int n = s.n, Q = s.n_q_1d;
  const size_t npc = dim == 2 ? n * n : n * n * n, qpc = dim == 2 ? Q * Q : Q * Q * Q;
  const size_t smem = (n * Q + Q * Q + Q + 2 * npc + (2 + dim) * qpc) * sizeof(Number);
  B200MF_REQUIRE(smem <= 200 * 1024, "n_q_points_1d = %d needs %zu bytes of shared memory per cell", Q, smem);
  if (smem > 48 * 1024)
    B200MF_CUDA_CHECK(cudaFuncSetAttribute(overint_kernel<dim, Number>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  const unsigned grid = (unsigned)std::min<uint64_t>(ce - cb, 148ull * 8);
  overint_kernel<dim, Number><<<grid, kOverThreads, smem, st>>>(p);
  count_launch();
  B200MF_CUDA_CHECK(cudaGetLastError());
  return B200MF_OK;
}

int launch_overint(const Setup &s, const b200mf_operator &op, void *dst, const void *src, uint64_t cb, uint64_t ce,
                   cudaStream_t st, double *dot, bool diagonal) {
  if (ce <= cb) return B200MF_OK;
  if (s.dim == 2)
    return s.number == B200MF_F64 ? launch_overint_t<2, double>(s, op, dst, src, cb, ce, st, dot, diagonal)
                                  : launch_overint_t<2, float>(s, op, dst, src, cb, ce, st, dot, diagonal);
  return s.number == B200MF_F64 ? launch_overint_t<3, double>(s, op, dst, src, cb, ce, st, dot, diagonal)
                                : launch_overint_t<3, float>(s, op, dst, src, cb, ce, st, dot, diagonal);
}

} // namespace b200mf
