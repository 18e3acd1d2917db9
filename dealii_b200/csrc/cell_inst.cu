// Instantiates the cell kernels for one value of n = degree + 1 (compile with -DB200MF_N=n);
// the build compiles this file once per degree so the eight degrees build in parallel.
#include <algorithm>
#include <cstdlib>

#include "cell_kernels.cuh"
#include "cell_kernels_plane.cuh"
#include "brick_kernel.cuh"
#include "bulk_kernel.cuh"
#include <mutex>

#ifndef B200MF_N
#error "compile with -DB200MF_N=<degree+1>"
#endif

namespace b200mf {

namespace {

template <int dim, int n, typename Number, int KIND>
int launch_one(const Setup &s, const b200mf_operator &op, void *dst, const void *src,
               uint64_t cell_begin, uint64_t cell_end, cudaStream_t stream, bool diagonal,
               double *dot_accum, bool masked) {
  using Cfg = BlockCfg<dim, n>;
  CellKernelParams<dim, n, Number, KIND> p;
  fill_shape_data<Number, n>(s, p.shape);
  p.l2g = s.d_l2g;
  p.mask = (s.any_mask && masked) ? s.d_mask : nullptr;
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
  if constexpr (dim == 3 && n <= 6) {
    // fast path: plane-per-thread kernel (cells without hanging-node masks)
    // kernel choice by measurement (profiles/r01_kernel_choice.txt): the plane kernel wins
    // wherever its register planes fit; B200MF_KERNEL=v1|plane overrides for A/B runs
    static const int forced = [] {
      const char *e = std::getenv("B200MF_KERNEL");
      if (e == nullptr) return 0;
      return std::string(e) == "v1" ? 1 : (std::string(e) == "plane" ? 2 : 0);
    }();
    bool use_plane;
    if (KIND == B200MF_CELLS_GENERAL)
      use_plane = sizeof(Number) == 8 ? (n <= 4) : (n <= 3 || n == 5);
    else
      use_plane = n <= 5;
    if (forced == 1) use_plane = false;
    if (forced == 2) use_plane = true;
    if (!diagonal && p.mask == nullptr && use_plane) {
      using PCfg = PlaneCfg<n, Number>;
      auto kernel = dot_accum != nullptr ? cell_loop_plane_kernel<n, Number, KIND, true>
                                         : cell_loop_plane_kernel<n, Number, KIND, false>;
      // persistent grid: SMs x resident CTAs per SM, cached per device (the dynamic shared memory
      // opt-in and the occupancy are per-device properties)
      static std::mutex plane_mtx;
      static int resident_ctas_v[64][2] = {};
      int dev = 0;
      B200MF_CUDA_CHECK(cudaGetDevice(&dev));
      if (dev < 0 || dev >= 64) dev = 0;
      std::lock_guard<std::mutex> plane_lock(plane_mtx);
      int &resident_ctas = resident_ctas_v[dev][dot_accum != nullptr ? 1 : 0];
      if (resident_ctas == 0) {
        B200MF_CUDA_CHECK(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                               (int)PCfg::smem_bytes));
        int sms = 0, per_sm = 0;
        B200MF_CUDA_CHECK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
        B200MF_CUDA_CHECK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, PCfg::threads,
                                                                        PCfg::smem_bytes));
        if (per_sm < 1) per_sm = 1;
        resident_ctas = sms * per_sm;
      }
      const uint64_t needed = (n_cells + PCfg::cells - 1) / PCfg::cells;
      const unsigned pgrid = (unsigned)std::min<uint64_t>(needed, (uint64_t)resident_ctas);
      kernel<<<pgrid, PCfg::threads, PCfg::smem_bytes, stream>>>(p);
      count_launch();
      B200MF_CUDA_CHECK(cudaGetLastError());
      return B200MF_OK;
    }
  }
  if (diagonal && p.mask == nullptr && s.d_diag_tables != nullptr) {
    // compute_diagonal by sum factorisation (see cell_kernels.cuh)
    smem += sizeof(Number) * 3 * n * n;
    auto kernel = cell_diagonal_sumfac_kernel<dim, n, Number, KIND>;
    B200MF_CUDA_CHECK(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                           (int)smem));
    kernel<<<grid, Cfg::threads, smem, stream>>>(p, static_cast<const Number *>(s.d_diag_tables));
  } else if (diagonal) {
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
                uint64_t b, uint64_t e, cudaStream_t st, bool diag, double *dot, bool masked) {
  switch (s.cell_kind) {
    case B200MF_CELLS_CARTESIAN:
      return launch_one<dim, n, Number, B200MF_CELLS_CARTESIAN>(s, op, dst, src, b, e, st, diag, dot, masked);
    case B200MF_CELLS_AFFINE:
      return launch_one<dim, n, Number, B200MF_CELLS_AFFINE>(s, op, dst, src, b, e, st, diag, dot, masked);
    default:
      return launch_one<dim, n, Number, B200MF_CELLS_GENERAL>(s, op, dst, src, b, e, st, diag, dot, masked);
  }
}

template <int n>
int launch_n(const Setup &s, const b200mf_operator &op, void *dst, const void *src, uint64_t b,
             uint64_t e, cudaStream_t st, bool diag, double *dot, bool masked) {
  if (s.dim == 2) {
    return s.number == B200MF_F64 ? launch_kind<2, n, double>(s, op, dst, src, b, e, st, diag, dot, masked)
                                  : launch_kind<2, n, float>(s, op, dst, src, b, e, st, diag, dot, masked);
  }
  return s.number == B200MF_F64 ? launch_kind<3, n, double>(s, op, dst, src, b, e, st, diag, dot, masked)
                                : launch_kind<3, n, float>(s, op, dst, src, b, e, st, diag, dot, masked);
}

// resolve_hanging_nodes of ONE cell, for the golden-vector test of the device code path
template <int dim, int n, typename Number>
__global__ void __launch_bounds__(BlockCfg<dim, n>::threads)
resolve_hanging_nodes_kernel(Number *values, const Number *weights, unsigned mask, int transpose) {
  using Cfg = BlockCfg<dim, n>;
  constexpr int npc = Cfg::npc;
  __shared__ Number U[npc];
  const int tid = threadIdx.x;
  const bool active = tid < Cfg::lines; // cell 0 of the CTA
  for (int i = tid; i < npc; i += Cfg::threads) U[i] = values[i];
  __syncthreads();
  if (transpose) resolve_hanging_nodes_cell<dim, n, Number, true>(weights, mask, U, tid, active);
  else           resolve_hanging_nodes_cell<dim, n, Number, false>(weights, mask, U, tid, active);
  for (int i = tid; i < npc; i += Cfg::threads) values[i] = U[i];
}

template <int dim, int n, typename Number>
int debug_resolve_one(unsigned mask, int transpose, double *values_host) {
  constexpr int npc = BlockCfg<dim, n>::npc;
  std::vector<double> sv, sg, qw, qp, sub;
  build_fe_q_shape_data(n - 1, sv, sg, qw, qp, sub);
  std::vector<Number> w(sub.begin(), sub.end()), v(values_host, values_host + npc);
  Number *dw = nullptr, *dv = nullptr;
  B200MF_CUDA_CHECK(cudaMalloc((void **)&dw, sizeof(Number) * n * n));
  B200MF_CUDA_CHECK(cudaMalloc((void **)&dv, sizeof(Number) * npc));
  B200MF_CUDA_CHECK(cudaMemcpy(dw, w.data(), sizeof(Number) * n * n, cudaMemcpyHostToDevice));
  B200MF_CUDA_CHECK(cudaMemcpy(dv, v.data(), sizeof(Number) * npc, cudaMemcpyHostToDevice));
  resolve_hanging_nodes_kernel<dim, n, Number><<<1, BlockCfg<dim, n>::threads>>>(dv, dw, mask, transpose);
  count_launch();
  B200MF_CUDA_CHECK(cudaGetLastError());
  B200MF_CUDA_CHECK(cudaMemcpy(v.data(), dv, sizeof(Number) * npc, cudaMemcpyDeviceToHost));
  for (int i = 0; i < npc; ++i) values_host[i] = double(v[i]);
  cudaFree(dw);
  cudaFree(dv);
  return B200MF_OK;
}

template <int p, typename Number, bool DOT, bool STRIDED>
int launch_bricks_one(const Setup &s, const b200mf_operator &op, void *dst, const void *src,
                      uint64_t brick_begin, uint64_t n_bricks, cudaStream_t stream, double *dot_accum,
                      bool overwrite, uint32_t geom, const uint32_t *list) {
  constexpr int b = brick_edge(p);
  using Cfg = BrickCfg<p, b, Number>;
  BrickKernelParams<p, Number> prm;
  fill_brick_matrices<Number, p + 1>(s, op, prm.mat, geom);
  prm.map = s.d_brick_map;
  prm.src = static_cast<const Number *>(src);
  prm.dst = static_cast<Number *>(dst);
  prm.dot_accum = dot_accum;
  prm.brick_begin = brick_begin;
  prm.list = list;
  prm.overwrite = list ? 2 : (overwrite ? 1 : 0);
  prm.strided = STRIDED ? s.d_brick_strided : nullptr;
  if constexpr (STRIDED && b >= 2) {
    // z slabs (two per brick) when the two full lattice arrays allow only two CTAs per SM: more
    // resident warps for a kernel that is latency bound once the index traffic is gone
    static const bool slabs_off = std::getenv("B200MF_NO_SLABS") != nullptr;
    if (Cfg::smem_bytes > 64 * 1024 && !slabs_off) {
      using SCfg = SlabCfg<p, b, 2, Number>;
      auto skernel = brick_strided_slab_kernel<p, b, 2, Number, DOT>;
      B200MF_CUDA_CHECK(cudaFuncSetAttribute(skernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SCfg::smem_bytes));
      skernel<<<(unsigned)n_bricks, SCfg::threads, SCfg::smem_bytes, stream>>>(prm);
      count_launch();
      B200MF_CUDA_CHECK(cudaGetLastError());
      return B200MF_OK;
    }
  }
  auto kernel = brick_cartesian_kernel<p, b, Number, DOT, STRIDED>;
  // (the attribute is per device: set it on every launch, it is a cheap driver call)
  B200MF_CUDA_CHECK(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         (int)Cfg::smem_bytes));
  kernel<<<(unsigned)n_bricks, Cfg::threads, Cfg::smem_bytes, stream>>>(prm);
  count_launch();
  B200MF_CUDA_CHECK(cudaGetLastError());
  return B200MF_OK;
}

// bulk brick path: every brick of the setup in one persistent launch (vmult mode: dst is written,
// not added to)
template <int p, typename Number, bool DOT>
int launch_bulk_one(const Setup &s, const b200mf_operator &op, void *dst, const void *src,
                    cudaStream_t stream, double *dot_accum) {
  constexpr int b = brick_edge(p);
  using Cfg = BulkCfg<p, b, Number>;
  const Setup::Bulk &B = s.bulk;
  if (B.L != Cfg::L || B.TP != Cfg::threads) {
    set_error("bulk tables were built for another brick shape");
    return B200MF_ERR_INVALID;
  }
  BulkKernelParams<p, Number> prm;
  fill_brick_matrices<Number, p + 1>(s, op, prm.mat);
  prm.desc = B.d_desc; prm.other = B.d_other; prm.phdr = B.d_phdr;
  prm.own_pos = B.d_own_pos; prm.other_pos = B.d_other_pos;
  prm.flags = B.d_flags; prm.ticket = B.d_ticket;
  prm.src = static_cast<const Number *>(src);
  prm.dst = static_cast<Number *>(dst);
  prm.dot_accum = dot_accum;
  prm.n_exec = B.n_exec;
  prm.epoch = ++s.bulk.epoch;
  prm.boundary_begin = prm.boundary_end = 0;
  prm.ghost_ready = nullptr;
  prm.boundary_done = nullptr;
  prm.debug = std::getenv("B200MF_BULK_DEBUG") ? (uint32_t)std::atoi(std::getenv("B200MF_BULK_DEBUG")) : 0u;
  if (B.sync_ghost_ready != nullptr) {
    prm.boundary_begin = (uint32_t)B.exec_boundary_begin;
    prm.boundary_end = (uint32_t)B.exec_boundary_end;
    prm.ghost_ready = B.sync_ghost_ready;
    prm.boundary_done = B.sync_boundary_done;
  }
  auto kernel = bulk_brick_kernel<p, b, Number, DOT>;
  // per-device launch configuration (dynamic shared memory opt-in, persistent grid size)
  static std::mutex mtx;
  static int grid_of_device[64] = {0};
  int dev = 0;
  B200MF_CUDA_CHECK(cudaGetDevice(&dev));
  int grid;
  {
    std::lock_guard<std::mutex> lock(mtx);
    if (dev < 0 || dev >= 64) dev = 0;
    if (grid_of_device[dev] == 0) {
      B200MF_CUDA_CHECK(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                             (int)Cfg::smem_bytes));
      int sms = 0, per_sm = 0;
      B200MF_CUDA_CHECK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
      B200MF_CUDA_CHECK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, Cfg::threads,
                                                                      Cfg::smem_bytes));
      if (per_sm < 1) per_sm = 1;
      grid_of_device[dev] = sms * per_sm;
    }
    grid = grid_of_device[dev];
  }
  if ((uint64_t)grid > B.n_exec) grid = (int)B.n_exec;
  kernel<<<(unsigned)grid, Cfg::threads, Cfg::smem_bytes, stream>>>(prm);
  count_launch();
  B200MF_CUDA_CHECK(cudaGetLastError());
  return B200MF_OK;
}

} // namespace

#define B200MF_CAT2(a, b) a##b
#define B200MF_CAT(a, b) B200MF_CAT2(a, b)

int B200MF_CAT(launch_cells_n, B200MF_N)(const Setup &s, const b200mf_operator &op, void *dst,
                                         const void *src, uint64_t b, uint64_t e,
                                         cudaStream_t st, bool diag, double *dot, bool masked) {
  return launch_n<B200MF_N>(s, op, dst, src, b, e, st, diag, dot, masked);
}

int B200MF_CAT(debug_resolve_n, B200MF_N)(int dim, int number, unsigned mask, int transpose,
                                          double *values_host) {
  if (dim == 2)
    return number == B200MF_F64 ? debug_resolve_one<2, B200MF_N, double>(mask, transpose, values_host)
                                : debug_resolve_one<2, B200MF_N, float>(mask, transpose, values_host);
  return number == B200MF_F64 ? debug_resolve_one<3, B200MF_N, double>(mask, transpose, values_host)
                              : debug_resolve_one<3, B200MF_N, float>(mask, transpose, values_host);
}

int B200MF_CAT(launch_bulk_n, B200MF_N)(const Setup &s, const b200mf_operator &op, void *dst,
                                        const void *src, cudaStream_t st, double *dot) {
  constexpr int p = B200MF_N - 1;
  if (s.number == B200MF_F64)
    return dot ? launch_bulk_one<p, double, true>(s, op, dst, src, st, dot)
               : launch_bulk_one<p, double, false>(s, op, dst, src, st, dot);
  return dot ? launch_bulk_one<p, float, true>(s, op, dst, src, st, dot)
             : launch_bulk_one<p, float, false>(s, op, dst, src, st, dot);
}

#if B200MF_N <= 9
int B200MF_CAT(launch_bricks_n, B200MF_N)(const Setup &s, const b200mf_operator &op, void *dst,
                                          const void *src, uint64_t brick_begin, uint64_t n_bricks,
                                          cudaStream_t st__, double *dot, bool ow, uint32_t geom,
                                          const uint32_t *list) {
  constexpr int p = B200MF_N - 1;
  const bool st = s.d_brick_strided != nullptr && s.strided_enabled && list == nullptr;
#define B200MF_LB(T, D) (st ? launch_bricks_one<p, T, D, true>(s, op, dst, src, brick_begin, n_bricks, st_, dot, ow, geom, list) \
                            : launch_bricks_one<p, T, D, false>(s, op, dst, src, brick_begin, n_bricks, st_, dot, ow, geom, list))
  cudaStream_t st_ = st__;
  if (s.number == B200MF_F64) return dot ? B200MF_LB(double, true) : B200MF_LB(double, false);
  return dot ? B200MF_LB(float, true) : B200MF_LB(float, false);
#undef B200MF_LB
}
#endif

} // namespace b200mf
