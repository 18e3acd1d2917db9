// Instantiates the cell kernels for one value of n = degree + 1 (compile with -DB200MF_N=n);
// the build compiles this file once per degree so the eight degrees build in parallel.
#include "cell_kernels.cuh"

#ifndef B200MF_N
#error "compile with -DB200MF_N=<degree+1>"
#endif

namespace b200mf {

namespace {

template <int dim, int n, typename Number, int KIND>
int launch_one(const Setup &s, const b200mf_operator &op, void *dst, const void *src,
               uint64_t cell_begin, uint64_t cell_end, cudaStream_t stream, bool diagonal,
               double *dot_accum) {
  using Cfg = BlockCfg<dim, n>;
  CellKernelParams<dim, n, Number, KIND> p;
  fill_shape_data<Number, n>(s, p.shape);
  p.l2g = s.d_l2g;
  p.mask = s.any_mask ? s.d_mask : nullptr;
  p.geom_id = s.d_geom_id;
  p.geom_table = static_cast<const Number *>(s.d_geom_table);
  p.metric = static_cast<const Number *>(s.d_metric);
  p.jxw = static_cast<const Number *>(s.d_jxw);
  p.weights = static_cast<const Number *>(s.d_weights);
  p.src = static_cast<const Number *>(src);
  p.dst = static_cast<Number *>(dst);
  p.op.grad_coef = static_cast<const Number *>(op.grad_coefficient);
  p.op.mass_coef = static_cast<const Number *>(op.mass_coefficient);
  p.op.grad_const = Number(op.grad_constant);
  p.op.mass_const = Number(op.mass_constant);
  p.op.has_mass = (op.mass_coefficient != nullptr || op.mass_constant != 0.0) ? 1 : 0;
  p.cell_begin = cell_begin;
  p.cell_end = cell_end;
  p.dot_accum = dot_accum;
  if (KIND == B200MF_CELLS_GENERAL && p.op.has_mass && p.jxw == nullptr) {
    set_error("mass term requested but the setup holds no JxW array");
    return B200MF_ERR_INVALID;
  }
  if (cell_end <= cell_begin) return B200MF_OK;
  const uint64_t n_cells = cell_end - cell_begin;
  const unsigned grid = (unsigned)((n_cells + Cfg::cells - 1) / Cfg::cells);
  size_t smem = sizeof(Number) * Cfg::smem_elems;
  if (diagonal) {
    smem += sizeof(Number) * Cfg::cells * Cfg::npc;
    auto kernel = cell_diagonal_kernel<dim, n, Number, KIND>;
    B200MF_CUDA_CHECK(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                           (int)smem));
    kernel<<<grid, Cfg::threads, smem, stream>>>(p);
  } else {
    auto kernel = cell_loop_kernel<dim, n, Number, KIND>;
    B200MF_CUDA_CHECK(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                           (int)smem));
    kernel<<<grid, Cfg::threads, smem, stream>>>(p);
  }
  count_launch();
  B200MF_CUDA_CHECK(cudaGetLastError());
  return B200MF_OK;
}

template <int dim, int n, typename Number>
int launch_kind(const Setup &s, const b200mf_operator &op, void *dst, const void *src,
                uint64_t b, uint64_t e, cudaStream_t st, bool diag, double *dot) {
  switch (s.cell_kind) {
    case B200MF_CELLS_CARTESIAN:
      return launch_one<dim, n, Number, B200MF_CELLS_CARTESIAN>(s, op, dst, src, b, e, st, diag, dot);
    case B200MF_CELLS_AFFINE:
      return launch_one<dim, n, Number, B200MF_CELLS_AFFINE>(s, op, dst, src, b, e, st, diag, dot);
    default:
      return launch_one<dim, n, Number, B200MF_CELLS_GENERAL>(s, op, dst, src, b, e, st, diag, dot);
  }
}

template <int n>
int launch_n(const Setup &s, const b200mf_operator &op, void *dst, const void *src, uint64_t b,
             uint64_t e, cudaStream_t st, bool diag, double *dot) {
  if (s.dim == 2) {
    return s.number == B200MF_F64 ? launch_kind<2, n, double>(s, op, dst, src, b, e, st, diag, dot)
                                  : launch_kind<2, n, float>(s, op, dst, src, b, e, st, diag, dot);
  }
  return s.number == B200MF_F64 ? launch_kind<3, n, double>(s, op, dst, src, b, e, st, diag, dot)
                                : launch_kind<3, n, float>(s, op, dst, src, b, e, st, diag, dot);
}

} // namespace

#define B200MF_CAT2(a, b) a##b
#define B200MF_CAT(a, b) B200MF_CAT2(a, b)

int B200MF_CAT(launch_cells_n, B200MF_N)(const Setup &s, const b200mf_operator &op, void *dst,
                                         const void *src, uint64_t b, uint64_t e,
                                         cudaStream_t st, bool diag, double *dot) {
  return launch_n<B200MF_N>(s, op, dst, src, b, e, st, diag, dot);
}

} // namespace b200mf
