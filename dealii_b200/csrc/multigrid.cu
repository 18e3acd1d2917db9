// Geometric multigrid on the device: the V-cycle of step-37 with the engine's operator on every level.
//
// What it replaces in the reference (all host classes there):
//   Multigrid::level_v_step                    multigrid/multigrid.templates.h:112-171
//   PreconditionMG::vmult                      multigrid/multigrid.h ("copy_to_mg, cycle, copy_from_mg")
//   MGTransferMatrixFree::prolongate /
//     restrict_and_add                         multigrid/mg_transfer_matrix_free.templates.h
//                                              (cell-wise tensor-product embedding, fine-dof weights)
//   mg::SmootherRelaxation<PreconditionChebyshev>  apply() = vmult (zero guess), smooth() = step()
//   MGCoarseGridApplySmoother                  multigrid/mg_coarse.h (Chebyshev in solver mode on level 0)
//
// Levels are uniform refinements of one coarse mesh; the children of coarse cell c are the fine cells
// (c << dim) + k (Morton order, what b200mf_mesh_create numbers) unless the caller passes child tables.
// One CTA transfers one coarse cell: its n^dim values are expanded to the (2n-1)^dim lattice of its
// children by dim sweeps with the 1D embedding matrix (shared memory), and written through the fine
// level's index list -- restriction is the transposed sequence with the fine residual weighted by
// 1/valence, so that a fine dof shared by several cells counts once.
#include "solver_impl.cuh"

namespace b200mf {

template <typename Number>
__global__ void mg_valence_kernel(Number *count, const uint32_t *l2g, uint64_t entries) {
  for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < entries;
       i += (uint64_t)gridDim.x * blockDim.x) {
    const uint32_t idx = l2g[i];
    if (!(idx & B200MF_L2G_CONSTRAINED)) atomicAdd(count + idx, Number(1));
  }
}
template <typename Number>
__global__ void mg_invert_kernel(Number *v, uint64_t n) {
  for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n;
       i += (uint64_t)gridDim.x * blockDim.x)
    v[i] = v[i] != Number(0) ? Number(1) / v[i] : Number(0);
}
template <typename To, typename From>
__global__ void mg_convert_kernel(To *dst, const From *src, uint64_t n) {
  for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n;
       i += (uint64_t)gridDim.x * blockDim.x)
    dst[i] = To(src[i]);
}

constexpr int kMgThreads = 256;

// index of lattice node (X, Y, Z) of coarse cell c in the fine level's vector
template <int dim, int n>
__device__ __forceinline__ uint32_t mg_fine_index(const uint32_t *l2g_f, const uint32_t *child, uint64_t c,
                                                  int npc, int X, int Y, int Z, int &shared_dirs) {
  constexpr int p = n - 1;
  const int cx = X > p, cy = Y > p, cz = (dim == 3) ? (Z > p) : 0;
  shared_dirs = (X == p) + (Y == p) + ((dim == 3) ? (Z == p) : 0);
  const int k = cx + 2 * cy + 4 * cz;
  const int local = (X - p * cx) + n * ((Y - p * cy) + n * (Z - p * cz));
  const uint64_t fine = child ? child[c * (1u << dim) + k] : ((c << dim) + k);
  return l2g_f[fine * npc + local];
}

// the 1D embedding matrix travels as a kernel parameter: its entries are constant-bank operands of the
// fully unrolled line products
template <typename Number, int n>
struct MgMatrix {
  Number P[(2 * n - 1) * n];
};

// One thread owns one line of a sweep: n inputs from shared memory into registers, the M = 2n-1 outputs of
// the line with the matrix from the constant bank, M stores (prolongation; restriction is the transpose).
template <typename Number, int dim, int n>
__global__ void __launch_bounds__(kMgThreads)
mg_prolongate_kernel(Number *__restrict__ dst, const Number *__restrict__ src, const uint32_t *__restrict__ l2g_c,
                     const uint32_t *__restrict__ l2g_f, const uint32_t *__restrict__ child,
                     const __grid_constant__ MgMatrix<Number, n> mat, uint64_t n_coarse_cells) {
  extern __shared__ __align__(16) unsigned char mg_smem[];
  constexpr int M = 2 * n - 1, nz = dim == 3 ? n : 1, Mz = dim == 3 ? M : 1;
  constexpr int npc = n * n * nz, cap = M * M * Mz;
  Number *A = reinterpret_cast<Number *>(mg_smem), *B = A + cap;
  const int tid = threadIdx.x, nthr = blockDim.x;
  for (uint64_t c = blockIdx.x; c < n_coarse_cells; c += gridDim.x) {
    __syncthreads();
    for (int i = tid; i < npc; i += nthr) {
      const uint32_t idx = l2g_c[c * npc + i];
      A[i] = (idx & B200MF_L2G_CONSTRAINED) ? Number(0) : src[idx];
    }
    __syncthreads();
    // x: [z][y][i] -> [z][y][X], lines (z, y)
    for (int line = tid; line < n * nz; line += nthr) {
      Number u[n];
#pragma unroll
      for (int i = 0; i < n; ++i) u[i] = A[line * n + i];
#pragma unroll
      for (int X = 0; X < M; ++X) {
        Number acc = 0;
#pragma unroll
        for (int i = 0; i < n; ++i) acc += mat.P[X * n + i] * u[i];
        B[line * M + X] = acc;
      }
    }
    __syncthreads();
    // y: [z][j][X] -> [z][Y][X], lines (z, X)
    for (int line = tid; line < nz * M; line += nthr) {
      const int X = line % M, z = line / M;
      Number u[n];
#pragma unroll
      for (int j = 0; j < n; ++j) u[j] = B[(z * n + j) * M + X];
#pragma unroll
      for (int Y = 0; Y < M; ++Y) {
        Number acc = 0;
#pragma unroll
        for (int j = 0; j < n; ++j) acc += mat.P[Y * n + j] * u[j];
        A[(z * M + Y) * M + X] = acc;
      }
    }
    __syncthreads();
    const Number *R = A;
    if (dim == 3) {
      // z: [k][Y][X] -> [Z][Y][X], lines (Y, X)
      for (int line = tid; line < M * M; line += nthr) {
        Number u[n];
#pragma unroll
        for (int k = 0; k < n; ++k) u[k] = A[k * M * M + line];
#pragma unroll
        for (int Z = 0; Z < M; ++Z) {
          Number acc = 0;
#pragma unroll
          for (int k = 0; k < n; ++k) acc += mat.P[Z * n + k] * u[k];
          B[Z * M * M + line] = acc;
        }
      }
      __syncthreads();
      R = B;
    }
    for (int o = tid; o < cap; o += nthr) {
      const int X = o % M, Y = (o / M) % M, Z = o / (M * M);
      int shared_dirs;
      const uint32_t f = mg_fine_index<dim, n>(l2g_f, child, c, npc, X, Y, Z, shared_dirs);
      // nodes on the faces of the coarse cell are written by every coarse cell that has them: same value
      if (!(f & B200MF_L2G_CONSTRAINED)) dst[f] = R[o];
    }
  }
}

template <typename Number, int dim, int n>
__global__ void __launch_bounds__(kMgThreads)
mg_restrict_kernel(Number *__restrict__ dst, const Number *__restrict__ src, const Number *__restrict__ inv_valence,
                   const uint32_t *__restrict__ l2g_c, const uint32_t *__restrict__ l2g_f,
                   const uint32_t *__restrict__ child, const __grid_constant__ MgMatrix<Number, n> mat,
                   uint64_t n_coarse_cells) {
  extern __shared__ __align__(16) unsigned char mg_smem[];
  constexpr int M = 2 * n - 1, nz = dim == 3 ? n : 1, Mz = dim == 3 ? M : 1;
  constexpr int npc = n * n * nz, cap = M * M * Mz;
  Number *A = reinterpret_cast<Number *>(mg_smem), *B = A + cap;
  const int tid = threadIdx.x, nthr = blockDim.x;
  for (uint64_t c = blockIdx.x; c < n_coarse_cells; c += gridDim.x) {
    __syncthreads();
    Number *G = dim == 3 ? B : A;
    for (int o = tid; o < cap; o += nthr) {
      const int X = o % M, Y = (o / M) % M, Z = o / (M * M);
      int shared_dirs;
      const uint32_t f = mg_fine_index<dim, n>(l2g_f, child, c, npc, X, Y, Z, shared_dirs);
      // a node between children of this cell is seen 2^shared_dirs times inside it and 1/valence of its
      // value belongs to each fine cell
      G[o] = (f & B200MF_L2G_CONSTRAINED) ? Number(0) : src[f] * inv_valence[f] * Number(1 << shared_dirs);
    }
    __syncthreads();
    if (dim == 3) {
      // z^T: [Z][Y][X] -> [k][Y][X], lines (Y, X)
      for (int line = tid; line < M * M; line += nthr) {
        Number acc[n];
#pragma unroll
        for (int k = 0; k < n; ++k) acc[k] = 0;
#pragma unroll
        for (int Z = 0; Z < M; ++Z) {
          const Number v = B[Z * M * M + line];
#pragma unroll
          for (int k = 0; k < n; ++k) acc[k] += mat.P[Z * n + k] * v;
        }
#pragma unroll
        for (int k = 0; k < n; ++k) A[k * M * M + line] = acc[k];
      }
      __syncthreads();
    }
    // y^T: [k][Y][X] -> [k][j][X], lines (k, X)
    for (int line = tid; line < nz * M; line += nthr) {
      const int X = line % M, k = line / M;
      Number acc[n];
#pragma unroll
      for (int j = 0; j < n; ++j) acc[j] = 0;
#pragma unroll
      for (int Y = 0; Y < M; ++Y) {
        const Number v = A[(k * M + Y) * M + X];
#pragma unroll
        for (int j = 0; j < n; ++j) acc[j] += mat.P[Y * n + j] * v;
      }
#pragma unroll
      for (int j = 0; j < n; ++j) B[(k * n + j) * M + X] = acc[j];
    }
    __syncthreads();
    // x^T: [k][j][X] -> [k][j][i], lines (k, j), added into the coarse vector
    for (int line = tid; line < n * nz; line += nthr) {
      Number acc[n];
#pragma unroll
      for (int i = 0; i < n; ++i) acc[i] = 0;
#pragma unroll
      for (int X = 0; X < M; ++X) {
        const Number v = B[line * M + X];
#pragma unroll
        for (int i = 0; i < n; ++i) acc[i] += mat.P[X * n + i] * v;
      }
#pragma unroll
      for (int i = 0; i < n; ++i) {
        const uint32_t idx = l2g_c[c * npc + line * n + i];
        if (!(idx & B200MF_L2G_CONSTRAINED)) atomicAdd(dst + idx, acc[i]);
      }
    }
  }
}

struct MgLevel {
  Setup *s = nullptr;
  b200mf_operator op{};
  void *inv_diag = nullptr, *sol = nullptr, *defect = nullptr, *t = nullptr, *inv_valence = nullptr;
  uint32_t *d_child = nullptr; // children of the cells of the next coarser level (optional)
  const b200mf_partitioner *part = nullptr; // the level is partitioned over ranks (vectors carry ghost entries)
  double lmin = 1.0, lmax = 1.0, theta = 1.0, delta = 0.0;
  int degree = 1, eig_cg_iterations = 0;
};

struct Mg {
  int number = 0, dim = 0, n = 0;
  std::vector<MgLevel> levels;
  std::vector<double> P; // 1D embedding matrix [(2n-1)][n]
  uint64_t vmults = 0;
  void *top_in = nullptr, *top_out = nullptr; // conversion buffers for callers of the other number type
  // CUDA graphs of the V-cycle, one per (dst, src) pair it was called with (the CG calls it with the same two
  // work vectors every iteration): the coarse levels are launch-bound, a replay has no launch gaps
  struct Graph {
    const void *dst = nullptr, *src = nullptr;
    int calls = 0;
    cudaGraphExec_t exec = nullptr;
    uint64_t launches = 0, vmults = 0;
  };
  std::vector<Graph> graphs;
  bool graph_enabled = false;
  cudaStream_t graph_stream = nullptr;
  cudaEvent_t graph_in = nullptr, graph_out = nullptr;
};

template <typename Number>
static size_t mg_smem_bytes(const Mg &mg) {
  const int M = 2 * mg.n - 1;
  size_t cap = (size_t)M * M * (mg.dim == 3 ? M : 1);
  return 2 * cap * sizeof(Number);
}

// launch shape of the transfer kernels: small CTAs, many per SM (the sweeps of one coarse cell are short and
// separated by barriers: concurrency across cells hides them)
static int mg_threads() {
  static const int v = getenv("B200MF_MG_THREADS") ? atoi(getenv("B200MF_MG_THREADS")) : 128;
  return std::min(std::max(v, 32), kMgThreads);
}
static unsigned mg_grid(uint64_t n_cells) {
  static const int per_sm = getenv("B200MF_MG_CTAS_PER_SM") ? atoi(getenv("B200MF_MG_CTAS_PER_SM")) : 16;
  return (unsigned)std::min<uint64_t>(n_cells, 148ull * std::max(per_sm, 1));
}

template <typename Number, int dim, int n>
static int mg_transfer_launch(const Mg &mg, int fine_level, bool prolongate, Number *dst, const Number *src,
                              cudaStream_t st) {
  const MgLevel &f = mg.levels[fine_level], &c = mg.levels[fine_level - 1];
  const size_t smem = mg_smem_bytes<Number>(mg);
  const unsigned grid = mg_grid(c.s->n_cells);
  const int threads = mg_threads();
  if (smem > 48 * 1024) {
    // once per kernel and device would do; the call is cheap next to the launch
    B200MF_CUDA_CHECK(cudaFuncSetAttribute(mg_prolongate_kernel<Number, dim, n>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    B200MF_CUDA_CHECK(cudaFuncSetAttribute(mg_restrict_kernel<Number, dim, n>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  }
  MgMatrix<Number, n> mat;
  for (int i = 0; i < (2 * n - 1) * n; ++i) mat.P[i] = Number(mg.P[i]);
  if (prolongate)
    mg_prolongate_kernel<Number, dim, n><<<grid, threads, smem, st>>>(dst, src, c.s->d_l2g, f.s->d_l2g, f.d_child, mat,
                                                                      c.s->n_cells);
  else
    mg_restrict_kernel<Number, dim, n><<<grid, threads, smem, st>>>(dst, src, (const Number *)f.inv_valence, c.s->d_l2g,
                                                                    f.s->d_l2g, f.d_child, mat, c.s->n_cells);
  count_launch();
  B200MF_CUDA_CHECK(cudaGetLastError());
  return B200MF_OK;
}

template <typename Number>
static int mg_transfer(const Mg &mg, int fine_level, bool prolongate, Number *dst, const Number *src, cudaStream_t st) {
#define B200MF_MG_CASE(N)                                                                                      \
  case N:                                                                                                      \
    return mg.dim == 3 ? mg_transfer_launch<Number, 3, N>(mg, fine_level, prolongate, dst, src, st)            \
                       : mg_transfer_launch<Number, 2, N>(mg, fine_level, prolongate, dst, src, st);
  switch (mg.n) {
    B200MF_MG_CASE(2) B200MF_MG_CASE(3) B200MF_MG_CASE(4) B200MF_MG_CASE(5)
    B200MF_MG_CASE(6) B200MF_MG_CASE(7) B200MF_MG_CASE(8) B200MF_MG_CASE(9)
  }
#undef B200MF_MG_CASE
  set_error("unsupported degree");
  return B200MF_ERR_INVALID;
}

// On partitioned levels the cells of a rank read the coarse (fine) values of dofs other ranks own: the ghost
// section of src is filled before the kernel (update_ghost_values) and cleared after it; restriction also
// sends what it added to coarse ghost entries to their owners (compress(add)).  A rank's fine cells are the
// children of its own coarse cells, so the kernels themselves stay local.
template <typename Number>
static int mg_prolongate(const Mg &mg, int to_level, Number *dst, const Number *src, cudaStream_t st) {
  const MgLevel &f = mg.levels[to_level], &c = mg.levels[to_level - 1];
  int rc;
  if (c.part && (rc = b200mf_update_ghost_values(c.part, const_cast<Number *>(src), st)) != B200MF_OK) return rc;
  if ((rc = mg_transfer<Number>(mg, to_level, true, dst, src, st)) != B200MF_OK) return rc;
  if (c.part && (rc = b200mf_zero_out_ghost_values(c.part, const_cast<Number *>(src), st)) != B200MF_OK) return rc;
  if (f.part && (rc = b200mf_zero_out_ghost_values(f.part, dst, st)) != B200MF_OK) return rc;
  // dofs constrained on the fine level (Dirichlet) stay zero
  return set_constrained_impl(*f.s, dst, 0.0, st);
}

template <typename Number>
static int mg_restrict_and_add(const Mg &mg, int from_level, Number *dst, const Number *src, cudaStream_t st) {
  const MgLevel &f = mg.levels[from_level], &c = mg.levels[from_level - 1];
  int rc;
  if (f.part && (rc = b200mf_update_ghost_values(f.part, const_cast<Number *>(src), st)) != B200MF_OK) return rc;
  if ((rc = mg_transfer<Number>(mg, from_level, false, dst, src, st)) != B200MF_OK) return rc;
  if (f.part && (rc = b200mf_zero_out_ghost_values(f.part, const_cast<Number *>(src), st)) != B200MF_OK) return rc;
  if (c.part && (rc = b200mf_compress_add(c.part, dst, st)) != B200MF_OK) return rc;
  return B200MF_OK;
}

// Multigrid::level_v_step (multigrid.templates.h:112-171); sol/defect of the finest level may be the
// caller's vectors
template <typename Number>
static int mg_level_v_step(Mg &mg, int level, Number *sol, const Number *defect, cudaStream_t st) {
  MgLevel &L = mg.levels[level];
  Setup &s = *L.s;
  const uint64_t n = s.n_owned;
  const unsigned grid = vec_grid(n);
  Chebyshev<Number> smoother{s, L.op, (const Number *)L.inv_diag, L.degree};
  smoother.theta = L.theta;
  smoother.delta = L.delta;
  smoother.part = L.part;
  int rc;
  if (level == 0) {
    rc = smoother.apply(sol, defect, st); // MGCoarseGridApplySmoother
    mg.vmults += smoother.vmults;
    return rc;
  }
  if ((rc = smoother.apply(sol, defect, st)) != B200MF_OK) return rc; // pre_smooth->apply
  Number *t = (Number *)L.t;
  if ((rc = level_vmult(s, L.part, L.op, t, sol, st, nullptr)) != B200MF_OK) return rc;
  mg.vmults++;
  sadd2_kernel<Number><<<grid, kVecThreads, 0, st>>>(t, Number(-1), Number(1), defect, n); // t = defect - A sol
  count_launch();
  MgLevel &C = mg.levels[level - 1];
  B200MF_CUDA_CHECK(cudaMemsetAsync(C.defect, 0, (C.s->n_owned + C.s->n_ghost) * sizeof(Number), st));
  if ((rc = mg_restrict_and_add<Number>(mg, level, (Number *)C.defect, t, st)) != B200MF_OK) return rc;
  if ((rc = mg_level_v_step<Number>(mg, level - 1, (Number *)C.sol, (const Number *)C.defect, st)) != B200MF_OK)
    return rc;
  // prolongate_and_add
  if ((rc = mg_prolongate<Number>(mg, level, t, (const Number *)C.sol, st)) != B200MF_OK) return rc;
  sadd2_kernel<Number><<<grid, kVecThreads, 0, st>>>(sol, Number(1), Number(1), t, n);
  count_launch();
  rc = smoother.step(sol, defect, st); // post_smooth->smooth
  mg.vmults += smoother.vmults;
  return rc;
}

// PreconditionMG::vmult for vectors of type Outer (converted to the level number when different)
template <typename Outer, typename Number>
static int mg_vcycle_eager(Mg &mg, Outer *dst, const Outer *src, cudaStream_t st);

// The first call with a (dst, src) pair runs eagerly (it also warms every lazily initialised piece), the second
// is captured on the multigrid's own stream into a graph, later ones replay it.  The caller's stream is ordered
// before and after the replay with two events.  Anything unexpected while capturing switches graphs off.
template <typename Outer, typename Number>
static int mg_vcycle(Mg &mg, Outer *dst, const Outer *src, cudaStream_t st) {
  if (!mg.graph_enabled) return mg_vcycle_eager<Outer, Number>(mg, dst, src, st);
  Mg::Graph *g = nullptr;
  for (auto &e : mg.graphs)
    if (e.dst == dst && e.src == src) g = &e;
  if (!g) {
    if (mg.graphs.size() >= 4) return mg_vcycle_eager<Outer, Number>(mg, dst, src, st);
    mg.graphs.emplace_back();
    g = &mg.graphs.back();
    g->dst = dst;
    g->src = src;
  }
  if (++g->calls == 1) return mg_vcycle_eager<Outer, Number>(mg, dst, src, st);
  if (!g->exec) {
    const uint64_t launches_before = g_launch_count.load(), vmults_before = mg.vmults;
    cudaGraph_t graph = nullptr;
    bool ok = cudaStreamBeginCapture(mg.graph_stream, cudaStreamCaptureModeRelaxed) == cudaSuccess;
    const int rc = ok ? mg_vcycle_eager<Outer, Number>(mg, dst, src, mg.graph_stream) : B200MF_ERR_CUDA;
    if (ok) ok = cudaStreamEndCapture(mg.graph_stream, &graph) == cudaSuccess && graph != nullptr;
    ok = ok && rc == B200MF_OK && cudaGraphInstantiate(&g->exec, graph, 0) == cudaSuccess;
    if (graph) cudaGraphDestroy(graph);
    g->launches = g_launch_count.load() - launches_before;
    g->vmults = mg.vmults - vmults_before;
    if (!ok) {
      cudaGetLastError(); // clear the capture error; nothing ran
      mg.graph_enabled = false;
      g->exec = nullptr;
      return mg_vcycle_eager<Outer, Number>(mg, dst, src, st);
    }
  } else {
    count_launch(g->launches);
    mg.vmults += g->vmults;
  }
  B200MF_CUDA_CHECK(cudaEventRecord(mg.graph_in, st));
  B200MF_CUDA_CHECK(cudaStreamWaitEvent(mg.graph_stream, mg.graph_in, 0));
  B200MF_CUDA_CHECK(cudaGraphLaunch(g->exec, mg.graph_stream));
  B200MF_CUDA_CHECK(cudaEventRecord(mg.graph_out, mg.graph_stream));
  B200MF_CUDA_CHECK(cudaStreamWaitEvent(st, mg.graph_out, 0));
  return B200MF_OK;
}

template <typename Outer, typename Number>
static int mg_vcycle_eager(Mg &mg, Outer *dst, const Outer *src, cudaStream_t st) {
  const int top = (int)mg.levels.size() - 1;
  MgLevel &L = mg.levels[top];
  const uint64_t n = L.s->n_owned;
  if (std::is_same<Outer, Number>::value)
    return mg_level_v_step<Number>(mg, top, (Number *)dst, (const Number *)src, st);
  if (!mg.top_in) {
    const uint64_t bytes = std::max<uint64_t>(n + L.s->n_ghost, 1) * sizeof(Number);
    B200MF_CUDA_CHECK(cudaMalloc(&mg.top_in, bytes));
    B200MF_CUDA_CHECK(cudaMalloc(&mg.top_out, bytes));
    B200MF_CUDA_CHECK(cudaMemsetAsync(mg.top_in, 0, bytes, st));
    B200MF_CUDA_CHECK(cudaMemsetAsync(mg.top_out, 0, bytes, st));
  }
  mg_convert_kernel<Number, Outer><<<vec_grid(n), kVecThreads, 0, st>>>((Number *)mg.top_in, src, n);
  count_launch();
  int rc = mg_level_v_step<Number>(mg, top, (Number *)mg.top_out, (const Number *)mg.top_in, st);
  if (rc != B200MF_OK) return rc;
  mg_convert_kernel<Outer, Number><<<vec_grid(n), kVecThreads, 0, st>>>(dst, (const Number *)mg.top_out, n);
  count_launch();
  return B200MF_OK;
}

template <typename Outer, typename Number>
struct MgPreconditioner {
  Mg &mg;
  uint64_t vmults = 0;
  int apply(Outer *z, const Outer *r, cudaStream_t st) {
    const uint64_t before = mg.vmults;
    int rc = mg_vcycle<Outer, Number>(mg, z, r, st);
    vmults += mg.vmults - before;
    return rc;
  }
};

template <typename Number>
static int mg_setup_levels(Mg &mg, const b200mf_mg_desc &d, cudaStream_t st) {
  build_prolongation_1d(mg.n - 1, mg.P);
  const int nl = (int)mg.levels.size();
  for (int l = 0; l < nl; ++l) {
    MgLevel &L = mg.levels[l];
    Setup &s = *L.s;
    const uint64_t n = s.n_owned, nt = s.n_owned + s.n_ghost, bytes = std::max<uint64_t>(nt, 1) * sizeof(Number);
    for (void **v : {&L.inv_diag, &L.sol, &L.defect, &L.t, &L.inv_valence}) {
      B200MF_CUDA_CHECK(cudaMalloc(v, bytes));
      B200MF_CUDA_CHECK(cudaMemsetAsync(*v, 0, bytes, st));
    }
    int rc = ensure_work(s, 5);
    if (rc != B200MF_OK) return rc;
    // inverse diagonal (LaplaceOperator::compute_diagonal + get_matrix_diagonal_inverse; constrained rows 1)
    if ((rc = launch_compute_diagonal(s, L.op, L.inv_diag, st)) != B200MF_OK) return rc;
    if (L.part && (rc = b200mf_compress_add(L.part, L.inv_diag, st)) != B200MF_OK) return rc;
    if ((rc = set_constrained_impl(s, L.inv_diag, 1.0, st)) != B200MF_OK) return rc;
    mg_invert_kernel<Number><<<vec_grid(n), kVecThreads, 0, st>>>((Number *)L.inv_diag, n);
    // 1 / (number of cells of this level, on all ranks, that hold the dof); ghost entries keep a copy
    const uint64_t entries = s.n_cells * (uint64_t)s.dofs_per_cell;
    mg_valence_kernel<Number><<<vec_grid(entries), kVecThreads, 0, st>>>((Number *)L.inv_valence, s.d_l2g, entries);
    if (L.part && (rc = b200mf_compress_add(L.part, L.inv_valence, st)) != B200MF_OK) return rc;
    mg_invert_kernel<Number><<<vec_grid(n), kVecThreads, 0, st>>>((Number *)L.inv_valence, n);
    if (L.part && (rc = b200mf_update_ghost_values(L.part, L.inv_valence, st)) != B200MF_OK) return rc;
    count_launch(3);
    B200MF_CUDA_CHECK(cudaGetLastError());
    if (l > 0 && d.child_cells && d.child_cells[l - 1]) {
      const uint64_t count = mg.levels[l - 1].s->n_cells << mg.dim;
      B200MF_CUDA_CHECK(cudaMalloc((void **)&L.d_child, count * sizeof(uint32_t)));
      B200MF_CUDA_CHECK(cudaMemcpyAsync(L.d_child, d.child_cells[l - 1], count * sizeof(uint32_t), cudaMemcpyHostToDevice, st));
      B200MF_CUDA_CHECK(cudaStreamSynchronize(st));
    }
    // smoother parameters (step-37.cc:965-984): level 0 is the Chebyshev "solver"
    const bool coarse = (l == 0);
    // eig_cg_n_iterations = mg_matrices[0].m(): the global size of level 0
    double n_global = double(n);
    if (level_is_distributed(L.part)) {
      B200MF_CUDA_CHECK(cudaMemcpyAsync(s.d_scratch + 58, &n_global, sizeof(double), cudaMemcpyHostToDevice, st));
      if ((rc = level_allreduce(L.part, s.d_scratch + 58, 1, st)) != B200MF_OK) return rc;
      B200MF_CUDA_CHECK(cudaMemcpyAsync(&n_global, s.d_scratch + 58, sizeof(double), cudaMemcpyDeviceToHost, st));
      B200MF_CUDA_CHECK(cudaStreamSynchronize(st));
    }
    const int eig_its = coarse ? (int)std::min<double>(n_global, double(1u << 20)) : d.eig_cg_n_iterations;
    uint64_t extra = 0;
    if ((rc = estimate_eigenvalues<Number>(s, L.op, (const Number *)L.inv_diag, eig_its, level_first_owned(L.part),
                                           d.safety_factor > 0 ? d.safety_factor : 1.2, st, L.lmin, L.lmax, extra,
                                           &L.eig_cg_iterations, /*zero_constrained=*/false, L.part)) != B200MF_OK)
      return rc;
    mg.vmults += extra;
    B200MF_REQUIRE(std::isfinite(L.lmin) && std::isfinite(L.lmax) && L.lmin > 0.0 && L.lmax >= L.lmin,
                   "multigrid level %d: eigenvalue estimate [%g, %g] is not positive (singular level operator? "
                   "the levels need Dirichlet constraints)", l, L.lmin, L.lmax);
    const double range = coarse ? d.coarse_tolerance : d.smoothing_range;
    const double alpha = range > 1.0 ? L.lmax / range : std::min(0.9 * L.lmax, L.lmin);
    L.degree = d.smoother_degree;
    if (coarse) {
      // PreconditionChebyshev::estimate_eigenvalues, degree == invalid_unsigned_int (precondition.h:3952-3982)
      const double actual_range = L.lmax / alpha;
      const double sigma = (1. - std::sqrt(1. / actual_range)) / (1. + std::sqrt(1. / actual_range));
      const double eps = range;
      const double deg = std::log(1. / eps + std::sqrt(1. / eps / eps - 1.)) / std::log(1. / sigma);
      B200MF_REQUIRE(std::isfinite(deg) && deg < 1e5, "multigrid level 0: Chebyshev solver degree %g out of range "
                     "(condition number %g): use fewer cells on the coarsest level", deg, actual_range);
      L.degree = 1 + (int)(unsigned)deg;
    }
    L.delta = (L.lmax - alpha) * 0.5;
    L.theta = (L.lmax + alpha) * 0.5;
  }
  return B200MF_OK;
}

template <typename Outer, typename Number>
static int mg_cg_solve_impl(Mg &mg, Setup &sys, const b200mf_partitioner *sys_part, const b200mf_operator &op,
                            double tolerance, int max_iterations, Outer *x, const Outer *b,
                            b200mf_solver_result *result, cudaStream_t st) {
  MgPreconditioner<Outer, Number> prec{mg};
  CgOptions opt{tolerance, max_iterations, false, false, 1};
  CgOutcome out;
  int rc = cg_generic<Outer>(sys, op, x, b, prec, opt, out, st, sys_part);
  if (rc != B200MF_OK) return rc;
  B200MF_CUDA_CHECK(cudaStreamSynchronize(st));
  if (result) {
    result->iterations = out.iterations;
    result->residual = out.residual;
    result->initial_residual = out.initial_residual;
    const MgLevel &top = mg.levels.back();
    result->chebyshev_max_eigenvalue = top.lmax;
    result->chebyshev_min_eigenvalue = top.lmin;
    result->operator_applications = out.vmults;
  }
  if (!out.success) {
    set_error("CG did not converge: %d iterations, residual %g (SolverControl::NoConvergence)", out.iterations,
              out.residual);
    return B200MF_ERR_NOCONVERGENCE;
  }
  return B200MF_OK;
}

static void mg_free(Mg &mg) {
  for (auto &g : mg.graphs)
    if (g.exec) cudaGraphExecDestroy(g.exec);
  if (mg.graph_stream) cudaStreamDestroy(mg.graph_stream);
  if (mg.graph_in) cudaEventDestroy(mg.graph_in);
  if (mg.graph_out) cudaEventDestroy(mg.graph_out);
  for (MgLevel &L : mg.levels) {
    for (void *v : {L.inv_diag, L.sol, L.defect, L.t, L.inv_valence}) cudaFree(v);
    cudaFree(L.d_child);
  }
  cudaFree(mg.top_in);
  cudaFree(mg.top_out);
}

} // namespace b200mf

using namespace b200mf;

struct b200mf_mg {
  b200mf::Mg impl;
};

#define B200MF_MG_DISPATCH(number, CALL)                              \
  do {                                                                \
    if ((number) == B200MF_F64) { using T = double; CALL; }           \
    else { using T = float; CALL; }                                   \
  } while (0)

extern "C" {

int b200mf_mg_create(const b200mf_mg_desc *d, b200mf_mg **out, void *stream) {
  B200MF_REQUIRE(d && out, "null argument");
  B200MF_REQUIRE(d->n_levels >= 1 && d->levels && d->operators, "multigrid needs at least one level");
  B200MF_REQUIRE(d->smoother_degree >= 1, "smoother degree must be positive");
  B200MF_REQUIRE(d->coarse_tolerance > 0.0 && d->coarse_tolerance < 1.0,
                 "coarse_tolerance is the relative tolerance of the level-0 Chebyshev solver, in (0, 1)");
  auto *h = new b200mf_mg;
  Mg &mg = h->impl;
  struct Guard {
    b200mf_mg *h;
    ~Guard() { if (h) { mg_free(h->impl); delete h; } }
  } guard{h};
  mg.levels.resize(d->n_levels);
  for (int l = 0; l < d->n_levels; ++l) {
    B200MF_REQUIRE(d->levels[l], "null level setup");
    Setup &s = const_cast<Setup &>(d->levels[l]->impl);
    mg.levels[l].s = &s;
    mg.levels[l].op = d->operators[l];
    if (l == 0) { mg.number = s.number; mg.dim = s.dim; mg.n = s.n; }
    B200MF_REQUIRE(s.number == mg.number && s.dim == mg.dim && s.n == mg.n,
                   "all levels must share number type, dimension and degree");
    mg.levels[l].part = d->partitioners ? d->partitioners[l] : nullptr;
    B200MF_REQUIRE(!s.any_mask, "multigrid levels are meshes without hanging nodes");
    B200MF_REQUIRE(s.n_ghost == 0 || mg.levels[l].part, "a level with ghost dofs needs its partitioner");
    if (l > 0)
      B200MF_REQUIRE(s.n_cells == (mg.levels[l - 1].s->n_cells << mg.dim),
                     "level l+1 must be level l refined once (2^dim children per cell)");
  }
  int rc;
  B200MF_MG_DISPATCH(mg.number, rc = mg_setup_levels<T>(mg, *d, (cudaStream_t)stream));
  if (rc != B200MF_OK) return rc;
  B200MF_CUDA_CHECK(cudaStreamSynchronize((cudaStream_t)stream));
  // graphs: one rank only (no NCCL inside a capture), no level on the bulk path (its launch epoch is a kernel
  // argument), not switched off by B200MF_MG_GRAPH=0
  mg.graph_enabled = !(getenv("B200MF_MG_GRAPH") && atoi(getenv("B200MF_MG_GRAPH")) == 0);
  for (const MgLevel &L : mg.levels)
    if (level_is_distributed(L.part) || (L.s->bulk.ready && L.s->bulk.enabled)) mg.graph_enabled = false;
  if (mg.graph_enabled) {
    B200MF_CUDA_CHECK(cudaStreamCreateWithFlags(&mg.graph_stream, cudaStreamNonBlocking));
    B200MF_CUDA_CHECK(cudaEventCreateWithFlags(&mg.graph_in, cudaEventDisableTiming));
    B200MF_CUDA_CHECK(cudaEventCreateWithFlags(&mg.graph_out, cudaEventDisableTiming));
  }
  guard.h = nullptr;
  *out = h;
  return B200MF_OK;
}

int b200mf_mg_prolongation_matrix_1d(int degree, double *out) {
  B200MF_REQUIRE(out && degree >= 1 && degree <= 8, "degree must be in 1..8");
  std::vector<double> P;
  build_prolongation_1d(degree, P);
  std::copy(P.begin(), P.end(), out);
  return B200MF_OK;
}

void b200mf_mg_destroy(b200mf_mg *h) {
  if (!h) return;
  mg_free(h->impl);
  delete h;
}

int b200mf_mg_get_level_info(const b200mf_mg *h, int level, b200mf_mg_level_info *info) {
  B200MF_REQUIRE(h && info, "null argument");
  B200MF_REQUIRE(level >= 0 && level < (int)h->impl.levels.size(), "no such level");
  const MgLevel &L = h->impl.levels[level];
  info->eig_min = L.lmin;
  info->eig_max = L.lmax;
  info->degree = L.degree;
  info->eig_cg_iterations = L.eig_cg_iterations;
  info->n_dofs = L.s->n_owned;
  info->inverse_diagonal = L.inv_diag;
  return B200MF_OK;
}

int b200mf_mg_prolongate(const b200mf_mg *h, int to_level, void *dst, const void *src, void *stream) {
  B200MF_REQUIRE(h && dst && src, "null argument");
  B200MF_REQUIRE(to_level >= 1 && to_level < (int)h->impl.levels.size(), "no such level");
  B200MF_MG_DISPATCH(h->impl.number, return mg_prolongate<T>(h->impl, to_level, (T *)dst, (const T *)src, (cudaStream_t)stream));
}

int b200mf_mg_restrict_and_add(const b200mf_mg *h, int from_level, void *dst, const void *src, void *stream) {
  B200MF_REQUIRE(h && dst && src, "null argument");
  B200MF_REQUIRE(from_level >= 1 && from_level < (int)h->impl.levels.size(), "no such level");
  B200MF_MG_DISPATCH(h->impl.number, return mg_restrict_and_add<T>(h->impl, from_level, (T *)dst, (const T *)src, (cudaStream_t)stream));
}

int b200mf_mg_vcycle(b200mf_mg *h, int number, void *dst, const void *src, void *stream) {
  B200MF_REQUIRE(h && dst && src, "null argument");
  B200MF_REQUIRE(number == B200MF_F64 || number == B200MF_F32, "bad number type");
  Mg &mg = h->impl;
  cudaStream_t st = (cudaStream_t)stream;
  if (number == B200MF_F64) {
    if (mg.number == B200MF_F64) return mg_vcycle<double, double>(mg, (double *)dst, (const double *)src, st);
    return mg_vcycle<double, float>(mg, (double *)dst, (const double *)src, st);
  }
  if (mg.number == B200MF_F64) return mg_vcycle<float, double>(mg, (float *)dst, (const float *)src, st);
  return mg_vcycle<float, float>(mg, (float *)dst, (const float *)src, st);
}

int b200mf_mg_dist_cg_solve(b200mf_mg *h, const b200mf_setup *system, const b200mf_partitioner *system_partitioner,
                            const b200mf_operator *op, double tolerance, int max_iterations, void *x, const void *b,
                            b200mf_solver_result *result, void *stream) {
  B200MF_REQUIRE(h && system && op && x && b, "null argument");
  Mg &mg = h->impl;
  Setup &sys = const_cast<Setup &>(system->impl);
  B200MF_REQUIRE(sys.n_owned == mg.levels.back().s->n_owned && sys.n_ghost == mg.levels.back().s->n_ghost,
                 "the system operator must live on the finest multigrid level");
  B200MF_REQUIRE(sys.n_ghost == 0 || system_partitioner, "a system operator with ghost dofs needs its partitioner");
  cudaStream_t st = (cudaStream_t)stream;
  const b200mf_partitioner *sp = system_partitioner;
  if (sys.number == B200MF_F64) {
    if (mg.number == B200MF_F64)
      return mg_cg_solve_impl<double, double>(mg, sys, sp, *op, tolerance, max_iterations, (double *)x, (const double *)b, result, st);
    return mg_cg_solve_impl<double, float>(mg, sys, sp, *op, tolerance, max_iterations, (double *)x, (const double *)b, result, st);
  }
  if (mg.number == B200MF_F64)
    return mg_cg_solve_impl<float, double>(mg, sys, sp, *op, tolerance, max_iterations, (float *)x, (const float *)b, result, st);
  return mg_cg_solve_impl<float, float>(mg, sys, sp, *op, tolerance, max_iterations, (float *)x, (const float *)b, result, st);
}

int b200mf_mg_cg_solve(b200mf_mg *h, const b200mf_setup *system, const b200mf_operator *op, double tolerance,
                       int max_iterations, void *x, const void *b, b200mf_solver_result *result, void *stream) {
  return b200mf_mg_dist_cg_solve(h, system, nullptr, op, tolerance, max_iterations, x, b, result, stream);
}

} // extern "C"
