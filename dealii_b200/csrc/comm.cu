// Multi-GPU layer behind the C ABI: communicator, Partitioner, ghost exchange, the distributed
// cell loop and the distributed CG -- so that a C++ deal.II host runs on several GPUs without
// embedding Python.
//
// Replaces (paths relative to the deal.II tree):
//   Utilities::MPI::Partitioner            base/partitioner.h, source/base/partitioner.cc:185-330
//     (ghost_targets / import_targets / import_indices; format pinned by
//      tests/mpi/parallel_partitioner_03.mpirun=4.output, tests/test_partition.py)
//   export_to_ghosted_array_start/finish,
//   import_from_ghosted_array_start/finish base/partitioner.templates.h:39-198, 290-671
//   LA::d::Vector::update_ghost_values / compress(add) / zero_out_ghost_values
//                                          lac/la_parallel_vector.templates.h:1026-1332
//   Portable::MatrixFree::distributed_cell_loop  matrix_free/portable_matrix_free.templates.h:1567-1690
//   SolverCG on distributed vectors (Utilities::MPI::sum of the partial sums) lac/solver_cg.h:871-893
// Transport: NCCL point-to-point (ncclSend/ncclRecv grouped per exchange) over NVLink on a
// high-priority stream so that the transfer kernels get SM slots while the cell loop has CTAs
// queued; the CG scalars go through one ncclAllReduce per reduction.  NCCL is loaded with dlopen at
// b200mf_comm_create (a process that already loaded an NCCL -- e.g. through torch -- shares it).
#include <dlfcn.h>
#include <nccl.h>

#include <algorithm>
#include <cmath>
#include <cstring>

#include "vector_ops.cuh"

namespace b200mf {
namespace {

struct NcclApi {
  void *handle = nullptr;
  ncclResult_t (*GetUniqueId)(ncclUniqueId *) = nullptr;
  ncclResult_t (*CommInitRank)(ncclComm_t *, int, ncclUniqueId, int) = nullptr;
  ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
  ncclResult_t (*GroupStart)() = nullptr;
  ncclResult_t (*GroupEnd)() = nullptr;
  ncclResult_t (*Send)(const void *, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*Recv)(void *, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*AllReduce)(const void *, void *, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t,
                            cudaStream_t) = nullptr;
  ncclResult_t (*AllGather)(const void *, void *, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
  const char *(*GetErrorString)(ncclResult_t) = nullptr;
};

NcclApi g_nccl;

int load_nccl() {
  if (g_nccl.handle != nullptr) return B200MF_OK;
  const char *names[] = {"libnccl.so.2", "libnccl.so"};
  void *h = nullptr;
  for (const char *nm : names) {
    h = dlopen(nm, RTLD_NOW | RTLD_GLOBAL);
    if (h) break;
  }
  if (!h) {
    set_error("cannot load NCCL (libnccl.so.2): %s", dlerror());
    return B200MF_ERR_COMM;
  }
#define SYM(field, name)                                                     \
  *(void **)(&g_nccl.field) = dlsym(h, name);                                \
  if (g_nccl.field == nullptr) {                                             \
    set_error("NCCL symbol %s missing", name);                               \
    return B200MF_ERR_COMM;                                                  \
  }
  SYM(GetUniqueId, "ncclGetUniqueId")
  SYM(CommInitRank, "ncclCommInitRank")
  SYM(CommDestroy, "ncclCommDestroy")
  SYM(GroupStart, "ncclGroupStart")
  SYM(GroupEnd, "ncclGroupEnd")
  SYM(Send, "ncclSend")
  SYM(Recv, "ncclRecv")
  SYM(AllReduce, "ncclAllReduce")
  SYM(AllGather, "ncclAllGather")
  SYM(GetErrorString, "ncclGetErrorString")
#undef SYM
  g_nccl.handle = h;
  return B200MF_OK;
}

#define B200MF_NCCL_CHECK(expr)                                                          \
  do {                                                                                   \
    ncclResult_t r__ = (expr);                                                           \
    if (r__ != ncclSuccess) {                                                            \
      ::b200mf::set_error("%s:%d: %s failed: %s", __FILE__, __LINE__, #expr,             \
                          g_nccl.GetErrorString ? g_nccl.GetErrorString(r__) : "?");     \
      return B200MF_ERR_COMM;                                                            \
    }                                                                                    \
  } while (0)

} // namespace
} // namespace b200mf

using namespace b200mf;

struct b200mf_comm {
  ncclComm_t nccl = nullptr;
  int n_ranks = 1, rank = 0;
  cudaStream_t comm_stream = nullptr;
  cudaEvent_t ev[4] = {nullptr, nullptr, nullptr, nullptr};
  double *d_tmp = nullptr; // small device scratch for reductions
};

struct b200mf_partitioner {
  b200mf_comm *comm = nullptr; // null: host-only object (index algebra, no exchange)
  int number = B200MF_F64, n_ranks = 1, rank = 0;
  uint64_t n_owned = 0, n_ghost = 0, n_import = 0, first_owned = 0;
  std::vector<int> ghost_rank, import_rank;
  std::vector<uint64_t> ghost_count, import_count;
  std::vector<uint32_t> import_indices; // local owned indices, concatenated per import target
  uint32_t *d_import_idx = nullptr;
  void *d_buf = nullptr;
};

namespace {

// the host half of Utilities::MPI::Partitioner::set_ghost_indices (source/base/partitioner.cc:
// 185-330): owners of my ghosts from the owned ranges
int owners_of_ghosts(b200mf_partitioner &p, const uint64_t *rank_offsets, const uint64_t *ghost_global) {
  p.ghost_rank.clear();
  p.ghost_count.clear();
  uint64_t prev = 0;
  for (uint64_t i = 0; i < p.n_ghost; ++i) {
    const uint64_t g = ghost_global[i];
    B200MF_REQUIRE(i == 0 || g > prev, "ghost indices must be sorted and unique");
    prev = g;
    const uint64_t *it = std::upper_bound(rank_offsets, rank_offsets + p.n_ranks + 1, g);
    const int owner = (int)(it - rank_offsets) - 1;
    B200MF_REQUIRE(owner >= 0 && owner < p.n_ranks && owner != p.rank, "ghost index %llu is not owned by another rank",
                   (unsigned long long)g);
    if (p.ghost_rank.empty() || p.ghost_rank.back() != owner) {
      p.ghost_rank.push_back(owner);
      p.ghost_count.push_back(0);
    }
    ++p.ghost_count.back();
  }
  return B200MF_OK;
}

int upload_partitioner(b200mf_partitioner &p) {
  int dev_count = 0;
  if (cudaGetDeviceCount(&dev_count) != cudaSuccess || dev_count == 0) {
    (void)cudaGetLastError();
    return B200MF_OK; // host-only use (index algebra tests)
  }
  const size_t ns = p.number == B200MF_F64 ? 8 : 4;
  B200MF_CUDA_CHECK(cudaMalloc((void **)&p.d_import_idx, std::max<size_t>(p.n_import, 1) * sizeof(uint32_t)));
  B200MF_CUDA_CHECK(cudaMemcpy(p.d_import_idx, p.import_indices.data(), p.n_import * sizeof(uint32_t),
                               cudaMemcpyHostToDevice));
  B200MF_CUDA_CHECK(cudaMalloc(&p.d_buf, std::max<size_t>(p.n_import, 1) * ns));
  return B200MF_OK;
}

// one exchange: owner -> ghost (update) or ghost -> owner (compress transport into d_buf)
int exchange(const b200mf_partitioner &p, void *vec, bool to_ghosts, cudaStream_t st) {
  const size_t ns = p.number == B200MF_F64 ? 8 : 4;
  const ncclDataType_t ty = p.number == B200MF_F64 ? ncclFloat64 : ncclFloat32;
  char *ghost = static_cast<char *>(vec) + p.n_owned * ns;
  char *buf = static_cast<char *>(p.d_buf);
  B200MF_NCCL_CHECK(g_nccl.GroupStart());
  uint64_t off = 0;
  for (size_t i = 0; i < p.ghost_rank.size(); ++i) {
    if (to_ghosts) B200MF_NCCL_CHECK(g_nccl.Recv(ghost + off * ns, p.ghost_count[i], ty, p.ghost_rank[i], p.comm->nccl, st));
    else           B200MF_NCCL_CHECK(g_nccl.Send(ghost + off * ns, p.ghost_count[i], ty, p.ghost_rank[i], p.comm->nccl, st));
    off += p.ghost_count[i];
  }
  off = 0;
  for (size_t i = 0; i < p.import_rank.size(); ++i) {
    if (to_ghosts) B200MF_NCCL_CHECK(g_nccl.Send(buf + off * ns, p.import_count[i], ty, p.import_rank[i], p.comm->nccl, st));
    else           B200MF_NCCL_CHECK(g_nccl.Recv(buf + off * ns, p.import_count[i], ty, p.import_rank[i], p.comm->nccl, st));
    off += p.import_count[i];
  }
  B200MF_NCCL_CHECK(g_nccl.GroupEnd());
  return B200MF_OK;
}

} // namespace

extern "C" {

int b200mf_comm_get_unique_id(void *id_out) {
  B200MF_REQUIRE(id_out, "null argument");
  int rc = load_nccl();
  if (rc != B200MF_OK) return rc;
  ncclUniqueId id;
  B200MF_NCCL_CHECK(g_nccl.GetUniqueId(&id));
  std::memcpy(id_out, &id, sizeof(id));
  return B200MF_OK;
}

int b200mf_comm_create(const void *unique_id, int n_ranks, int rank, b200mf_comm **out) {
  B200MF_REQUIRE(unique_id && out && n_ranks >= 1 && rank >= 0 && rank < n_ranks, "bad argument");
  int rc = load_nccl();
  if (rc != B200MF_OK) return rc;
  b200mf_comm *c = new b200mf_comm();
  c->n_ranks = n_ranks;
  c->rank = rank;
  ncclUniqueId id;
  std::memcpy(&id, unique_id, sizeof(id));
  ncclResult_t r = g_nccl.CommInitRank(&c->nccl, n_ranks, id, rank);
  if (r != ncclSuccess) {
    set_error("ncclCommInitRank failed: %s", g_nccl.GetErrorString(r));
    delete c;
    return B200MF_ERR_COMM;
  }
  int lo = 0, hi = 0;
  cudaDeviceGetStreamPriorityRange(&lo, &hi); // hi = numerically lowest = highest priority
  if (cudaStreamCreateWithPriority(&c->comm_stream, cudaStreamNonBlocking, hi) != cudaSuccess ||
      cudaMalloc((void **)&c->d_tmp, 64 * sizeof(double)) != cudaSuccess) {
    set_error("cannot create the communication stream");
    b200mf_comm_destroy(c);
    return B200MF_ERR_CUDA;
  }
  for (auto &e : c->ev)
    if (cudaEventCreateWithFlags(&e, cudaEventDisableTiming) != cudaSuccess) {
      set_error("cannot create events");
      b200mf_comm_destroy(c);
      return B200MF_ERR_CUDA;
    }
  *out = c;
  return B200MF_OK;
}

int b200mf_comm_destroy(b200mf_comm *c) {
  if (!c) return B200MF_OK;
  if (c->nccl && g_nccl.CommDestroy) g_nccl.CommDestroy(c->nccl);
  for (auto &e : c->ev)
    if (e) cudaEventDestroy(e);
  if (c->comm_stream) cudaStreamDestroy(c->comm_stream);
  cudaFree(c->d_tmp);
  delete c;
  return B200MF_OK;
}

int b200mf_comm_allreduce_sum(b200mf_comm *c, double *device_values, int count, void *stream) {
  B200MF_REQUIRE(c && device_values && count >= 0, "bad argument");
  if (c->n_ranks == 1 || count == 0) return B200MF_OK;
  B200MF_NCCL_CHECK(g_nccl.AllReduce(device_values, device_values, (size_t)count, ncclFloat64, ncclSum, c->nccl,
                                     (cudaStream_t)stream));
  return B200MF_OK;
}

// ---- Partitioner
int b200mf_partitioner_create_host(int n_ranks, int rank, const uint64_t *rank_offsets,
                                   const uint64_t *ghost_global, uint64_t n_ghost, int number,
                                   const int *import_ranks, const uint64_t *import_counts, int n_import_ranks,
                                   const uint64_t *import_global, b200mf_partitioner **out) {
  B200MF_REQUIRE(rank_offsets && out && n_ranks >= 1 && rank >= 0 && rank < n_ranks, "bad argument");
  B200MF_REQUIRE(n_ghost == 0 || ghost_global, "ghost_global is null");
  b200mf_partitioner *p = new b200mf_partitioner();
  p->number = number; p->n_ranks = n_ranks; p->rank = rank;
  p->first_owned = rank_offsets[rank];
  p->n_owned = rank_offsets[rank + 1] - rank_offsets[rank];
  p->n_ghost = n_ghost;
  int rc = owners_of_ghosts(*p, rank_offsets, ghost_global);
  if (rc != B200MF_OK) { delete p; return rc; }
  uint64_t pos = 0;
  for (int i = 0; i < n_import_ranks; ++i) {
    if (import_counts[i] == 0) continue;
    p->import_rank.push_back(import_ranks[i]);
    p->import_count.push_back(import_counts[i]);
    for (uint64_t k = 0; k < import_counts[i]; ++k, ++pos) {
      const uint64_t g = import_global[pos];
      if (g < p->first_owned || g >= p->first_owned + p->n_owned) {
        set_error("rank %d asks rank %d for index %llu which it does not own", import_ranks[i], rank,
                  (unsigned long long)g);
        delete p;
        return B200MF_ERR_INVALID;
      }
      p->import_indices.push_back((uint32_t)(g - p->first_owned));
    }
  }
  p->n_import = p->import_indices.size();
  rc = upload_partitioner(*p);
  if (rc != B200MF_OK) { b200mf_partitioner_destroy(p); return rc; }
  *out = p;
  return B200MF_OK;
}

// the same with the ghost lists exchanged over the communicator (what the reference's consensus
// algorithm does, source/base/partitioner.cc:241-267)
int b200mf_partitioner_create(b200mf_comm *c, const uint64_t *rank_offsets, const uint64_t *ghost_global,
                              uint64_t n_ghost, int number, b200mf_partitioner **out) {
  B200MF_REQUIRE(c && rank_offsets && out, "null argument");
  const int R = c->n_ranks, me = c->rank;
  b200mf_partitioner tmp;
  tmp.n_ranks = R; tmp.rank = me; tmp.n_ghost = n_ghost;
  int rc = owners_of_ghosts(tmp, rank_offsets, ghost_global);
  if (rc != B200MF_OK) return rc;
  // counts[i][j] = number of ghosts rank i holds of rank j's dofs
  std::vector<uint64_t> mine(R, 0), all((size_t)R * R, 0);
  for (size_t i = 0; i < tmp.ghost_rank.size(); ++i) mine[tmp.ghost_rank[i]] = tmp.ghost_count[i];
  uint64_t *d_counts = nullptr, *d_ghost = nullptr, *d_import = nullptr;
  cudaStream_t st = c->comm_stream;
  B200MF_CUDA_CHECK(cudaMalloc((void **)&d_counts, (size_t)R * (R + 1) * sizeof(uint64_t)));
  B200MF_CUDA_CHECK(cudaMemcpyAsync(d_counts, mine.data(), R * sizeof(uint64_t), cudaMemcpyHostToDevice, st));
  if (R > 1) B200MF_NCCL_CHECK(g_nccl.AllGather(d_counts, d_counts + R, R, ncclUint64, c->nccl, st));
  else B200MF_CUDA_CHECK(cudaMemcpyAsync(d_counts + R, d_counts, R * sizeof(uint64_t), cudaMemcpyDeviceToDevice, st));
  B200MF_CUDA_CHECK(cudaMemcpyAsync(all.data(), d_counts + R, (size_t)R * R * sizeof(uint64_t), cudaMemcpyDeviceToHost, st));
  B200MF_CUDA_CHECK(cudaStreamSynchronize(st));
  std::vector<int> import_ranks;
  std::vector<uint64_t> import_counts;
  uint64_t n_import = 0;
  for (int i = 0; i < R; ++i)
    if (i != me && all[(size_t)i * R + me] > 0) {
      import_ranks.push_back(i);
      import_counts.push_back(all[(size_t)i * R + me]);
      n_import += all[(size_t)i * R + me];
    }
  B200MF_CUDA_CHECK(cudaMalloc((void **)&d_ghost, std::max<uint64_t>(n_ghost, 1) * sizeof(uint64_t)));
  B200MF_CUDA_CHECK(cudaMalloc((void **)&d_import, std::max<uint64_t>(n_import, 1) * sizeof(uint64_t)));
  B200MF_CUDA_CHECK(cudaMemcpyAsync(d_ghost, ghost_global, n_ghost * sizeof(uint64_t), cudaMemcpyHostToDevice, st));
  if (R > 1) {
    B200MF_NCCL_CHECK(g_nccl.GroupStart());
    uint64_t off = 0;
    for (size_t i = 0; i < tmp.ghost_rank.size(); ++i) {
      B200MF_NCCL_CHECK(g_nccl.Send(d_ghost + off, tmp.ghost_count[i], ncclUint64, tmp.ghost_rank[i], c->nccl, st));
      off += tmp.ghost_count[i];
    }
    off = 0;
    for (size_t i = 0; i < import_ranks.size(); ++i) {
      B200MF_NCCL_CHECK(g_nccl.Recv(d_import + off, import_counts[i], ncclUint64, import_ranks[i], c->nccl, st));
      off += import_counts[i];
    }
    B200MF_NCCL_CHECK(g_nccl.GroupEnd());
  }
  std::vector<uint64_t> import_global(n_import);
  B200MF_CUDA_CHECK(cudaMemcpyAsync(import_global.data(), d_import, n_import * sizeof(uint64_t), cudaMemcpyDeviceToHost, st));
  B200MF_CUDA_CHECK(cudaStreamSynchronize(st));
  cudaFree(d_counts); cudaFree(d_ghost); cudaFree(d_import);
  rc = b200mf_partitioner_create_host(R, me, rank_offsets, ghost_global, n_ghost, number, import_ranks.data(),
                                      import_counts.data(), (int)import_ranks.size(), import_global.data(), out);
  if (rc == B200MF_OK) (*out)->comm = c;
  return rc;
}

int b200mf_partitioner_destroy(b200mf_partitioner *p) {
  if (!p) return B200MF_OK;
  cudaFree(p->d_import_idx);
  cudaFree(p->d_buf);
  delete p;
  return B200MF_OK;
}

int b200mf_partitioner_get_info(const b200mf_partitioner *p, b200mf_partitioner_info *info) {
  B200MF_REQUIRE(p && info, "null argument");
  info->n_owned = p->n_owned; info->n_ghost = p->n_ghost; info->n_import = p->n_import;
  info->n_ghost_targets = (int)p->ghost_rank.size();
  info->n_import_targets = (int)p->import_rank.size();
  info->ghost_target_ranks = p->ghost_rank.data(); info->ghost_target_counts = p->ghost_count.data();
  info->import_target_ranks = p->import_rank.data(); info->import_target_counts = p->import_count.data();
  info->import_indices = p->import_indices.data();
  return B200MF_OK;
}

// ---- ghost exchange on a caller's stream (blocking the stream, not the host)
int b200mf_update_ghost_values(const b200mf_partitioner *p, void *vec, void *stream) {
  B200MF_REQUIRE(p && vec, "null argument");
  if (p->n_ranks == 1 || (p->ghost_rank.empty() && p->import_rank.empty())) return B200MF_OK;
  B200MF_REQUIRE(p->comm, "partitioner has no communicator");
  int rc = b200mf_ghost_pack(p->number, p->d_buf, vec, p->d_import_idx, p->n_import, stream);
  if (rc != B200MF_OK) return rc;
  return exchange(*p, vec, true, (cudaStream_t)stream);
}

int b200mf_zero_out_ghost_values(const b200mf_partitioner *p, void *vec, void *stream) {
  B200MF_REQUIRE(p && vec, "null argument");
  const size_t ns = p->number == B200MF_F64 ? 8 : 4;
  if (p->n_ghost)
    B200MF_CUDA_CHECK(cudaMemsetAsync(static_cast<char *>(vec) + p->n_owned * ns, 0, p->n_ghost * ns, (cudaStream_t)stream));
  return B200MF_OK;
}

int b200mf_compress_add(const b200mf_partitioner *p, void *vec, void *stream) {
  B200MF_REQUIRE(p && vec, "null argument");
  if (p->n_ranks > 1 && !(p->ghost_rank.empty() && p->import_rank.empty())) {
    B200MF_REQUIRE(p->comm, "partitioner has no communicator");
    int rc = exchange(*p, vec, false, (cudaStream_t)stream);
    if (rc != B200MF_OK) return rc;
    rc = b200mf_ghost_unpack_add(p->number, vec, p->d_buf, p->d_import_idx, p->n_import, stream);
    if (rc != B200MF_OK) return rc;
  }
  return b200mf_zero_out_ghost_values(p, vec, stream);
}

// ---- distributed vmult: dst = 0; update_ghost_values(src) || interior cells (first half); cells
// touching ghosts; compress(dst) || interior cells (second half); zero ghosts; copy_constrained_values
static int dist_vmult_impl(const Setup &s, const b200mf_partitioner &p, const b200mf_operator &op, void *dst,
                           void *src, cudaStream_t main, double *dot_accum) {
  const bool single = p.n_ranks == 1 || (p.ghost_rank.empty() && p.import_rank.empty());
  if (single) {
    int rc = vmult_impl(s, op, dst, src, main, dot_accum);
    return rc;
  }
  b200mf_comm *c = p.comm;
  B200MF_REQUIRE(c, "partitioner has no communicator");
  cudaStream_t cs = c->comm_stream;
  // cells [0, ni) touch no ghost dof.  A setup WITH ghost dofs whose caller reports 0 interior cells (every
  // block of cells touches the interface, or no split was given) runs all cells after the ghost update.
  const uint64_t nc = s.n_cells, ni = s.n_ghost > 0 ? std::min(s.n_cells_interior, nc) : nc;
  const uint64_t W = s.n_bricks ? (uint64_t)s.brick_b * s.brick_b * s.brick_b : 1;
  const bool coloured = coloured_enabled(s, op);
  const uint64_t half = coloured ? s.colouring.half : (ni / 2) / W * W; // pieces never cut a brick
  auto piece = [&](int which, uint64_t cb, uint64_t ce) -> int {
    return coloured ? launch_coloured(s, op, dst, src, which, main, dot_accum)
                    : launch_cell_loop(s, op, dst, src, cb, ce, main, dot_accum, true);
  };
  int rc = coloured ? coloured_prepare(s, dst, main) : vmult_prepare_impl(s, op, dst, main);
  if (rc != B200MF_OK) return rc;
  B200MF_CUDA_CHECK(cudaEventRecord(c->ev[0], main));
  B200MF_CUDA_CHECK(cudaStreamWaitEvent(cs, c->ev[0], 0));
  rc = b200mf_ghost_pack(p.number, p.d_buf, src, p.d_import_idx, p.n_import, cs);
  if (rc != B200MF_OK) return rc;
  rc = exchange(p, src, true, cs);
  if (rc != B200MF_OK) return rc;
  B200MF_CUDA_CHECK(cudaEventRecord(c->ev[1], cs));
  rc = piece(0, 0, half); // interior, part A
  if (rc != B200MF_OK) return rc;
  B200MF_CUDA_CHECK(cudaStreamWaitEvent(main, c->ev[1], 0));
  rc = piece(1, ni, nc); // cells touching ghosts
  if (rc != B200MF_OK) return rc;
  B200MF_CUDA_CHECK(cudaEventRecord(c->ev[2], main));
  B200MF_CUDA_CHECK(cudaStreamWaitEvent(cs, c->ev[2], 0));
  rc = exchange(p, dst, false, cs);
  if (rc != B200MF_OK) return rc;
  B200MF_CUDA_CHECK(cudaEventRecord(c->ev[3], cs));
  rc = piece(2, half, ni); // interior, part B
  if (rc != B200MF_OK) return rc;
  B200MF_CUDA_CHECK(cudaStreamWaitEvent(main, c->ev[3], 0));
  rc = b200mf_ghost_unpack_add(p.number, dst, p.d_buf, p.d_import_idx, p.n_import, main);
  if (rc != B200MF_OK) return rc;
  const size_t ns = number_size(s.number);
  B200MF_CUDA_CHECK(cudaMemsetAsync(static_cast<char *>(dst) + s.n_owned * ns, 0, s.n_ghost * ns, main));
  B200MF_CUDA_CHECK(cudaMemsetAsync(static_cast<char *>(src) + s.n_owned * ns, 0, s.n_ghost * ns, main));
  return copy_constrained_impl(s, dst, src, main, dot_accum);
}

int b200mf_dist_vmult(const b200mf_setup *h, const b200mf_partitioner *p, const b200mf_operator *op, void *dst,
                      void *src, void *stream) {
  B200MF_REQUIRE(h && p && op && dst && src, "null argument");
  B200MF_REQUIRE(p->n_owned == h->impl.n_owned && p->n_ghost == h->impl.n_ghost,
                 "partitioner and setup disagree about the vector layout");
  return dist_vmult_impl(h->impl, *p, *op, dst, src, (cudaStream_t)stream, nullptr);
}

// n_vectors independent distributed vmults on HOST vectors (n_owned elements each, page-locked),
// pipelined like b200mf_vmult_host_batch: upload of vector k+1 and download of vector k-1 overlap the
// (ghost-exchanging) vmult of vector k.  Every rank calls it with the same n_vectors.
int b200mf_dist_vmult_host_batch(const b200mf_setup *h, const b200mf_partitioner *p, const b200mf_operator *op,
                                 int n_vectors, void *const *dst_host, const void *const *src_host) {
  B200MF_REQUIRE(h && p && op && n_vectors >= 0 && (n_vectors == 0 || (dst_host && src_host)), "null argument");
  Setup &s = const_cast<Setup &>(h->impl);
  const size_t ns = number_size(s.number);
  const size_t bytes = (s.n_owned + s.n_ghost) * ns, owned_bytes = s.n_owned * ns;
  for (int i = 0; i < 2; ++i) {
    if (!s.d_pipe_in[i]) B200MF_CUDA_CHECK(cudaMalloc(&s.d_pipe_in[i], std::max<size_t>(bytes, 1)));
    if (!s.d_pipe_out[i]) B200MF_CUDA_CHECK(cudaMalloc(&s.d_pipe_out[i], std::max<size_t>(bytes, 1)));
  }
  for (int i = 0; i < 3; ++i) {
    if (!s.pipe_stream[i]) B200MF_CUDA_CHECK(cudaStreamCreateWithFlags(&s.pipe_stream[i], cudaStreamNonBlocking));
    for (int j = 0; j < 2; ++j)
      if (!s.pipe_event[i][j]) B200MF_CUDA_CHECK(cudaEventCreateWithFlags(&s.pipe_event[i][j], cudaEventDisableTiming));
  }
  cudaStream_t s_in = s.pipe_stream[0], s_op = s.pipe_stream[1], s_out = s.pipe_stream[2];
  if (s.n_ghost)
    for (int i = 0; i < 2; ++i)
      B200MF_CUDA_CHECK(cudaMemsetAsync(static_cast<char *>(s.d_pipe_in[i]) + owned_bytes, 0, bytes - owned_bytes, s_in));
  for (int k = 0; k < n_vectors; ++k) {
    const int slot = k & 1;
    B200MF_REQUIRE(dst_host[k] && src_host[k], "null vector in batch");
    B200MF_CUDA_CHECK(cudaStreamWaitEvent(s_in, s.pipe_event[1][slot], 0));
    B200MF_CUDA_CHECK(cudaMemcpyAsync(s.d_pipe_in[slot], src_host[k], owned_bytes, cudaMemcpyHostToDevice, s_in));
    B200MF_CUDA_CHECK(cudaEventRecord(s.pipe_event[0][slot], s_in));
    B200MF_CUDA_CHECK(cudaStreamWaitEvent(s_op, s.pipe_event[0][slot], 0));
    B200MF_CUDA_CHECK(cudaStreamWaitEvent(s_op, s.pipe_event[2][slot], 0));
    int rc = dist_vmult_impl(s, *p, *op, s.d_pipe_out[slot], s.d_pipe_in[slot], s_op, nullptr);
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

int b200mf_dist_compute_diagonal(const b200mf_setup *h, const b200mf_partitioner *p, const b200mf_operator *op,
                                 void *diag, void *stream) {
  B200MF_REQUIRE(h && p && op && diag, "null argument");
  int rc = b200mf_compute_diagonal(h, op, diag, stream);
  if (rc != B200MF_OK || p->n_ranks == 1) return rc;
  rc = b200mf_compress_add(p, diag, stream);
  if (rc != B200MF_OK) return rc;
  return b200mf_set_constrained_values(h, diag, 1.0, stream);
}

// ---- distributed SolverCG with Jacobi (or no) preconditioner: the algebra and stopping rule of
// b200mf_cg_solve / lac/solver_cg.h:703-763; the partial sums of an iteration are all-reduced where
// the reference calls Utilities::MPI::sum.  The residual is read back every check_every iterations.
int b200mf_dist_cg_solve(const b200mf_setup *h, const b200mf_partitioner *p, const b200mf_operator *op,
                         const b200mf_solver_desc *sd, void *x, const void *b, b200mf_solver_result *result,
                         void *stream) {
  B200MF_REQUIRE(h && p && op && sd && x && b && result, "null argument");
  B200MF_REQUIRE(sd->preconditioner == B200MF_PRECOND_NONE || sd->preconditioner == B200MF_PRECOND_JACOBI,
                 "the distributed solver supports PreconditionIdentity and Jacobi");
  Setup &s = const_cast<Setup &>(h->impl);
  cudaStream_t st = (cudaStream_t)stream;
  const uint64_t n = s.n_owned, nt = s.n_owned + s.n_ghost;
  const size_t ns = number_size(s.number);
  for (int i = 0; i < 3; ++i)
    if (!s.d_work[i]) B200MF_CUDA_CHECK(cudaMalloc(&s.d_work[i], std::max<uint64_t>(nt, 1) * ns));
  void *r = s.d_work[0], *pv = s.d_work[1], *v = s.d_work[2];
  B200MF_CUDA_CHECK(cudaMemsetAsync(pv, 0, nt * ns, st));
  B200MF_CUDA_CHECK(cudaMemsetAsync(v, 0, nt * ns, st));
  double *sc = s.d_scratch, *hp = s.h_pinned;
  B200MF_CUDA_CHECK(cudaMemsetAsync(sc, 0, 64 * sizeof(double), st));
  const void *d = sd->preconditioner == B200MF_PRECOND_JACOBI ? sd->inverse_diagonal : nullptr;
  B200MF_REQUIRE(sd->preconditioner != B200MF_PRECOND_JACOBI || d != nullptr, "inverse_diagonal is null");
  b200mf_comm *c = p->comm;
  const bool multi = p->n_ranks > 1;
  B200MF_REQUIRE(!multi || c, "partitioner has no communicator");
  auto allreduce = [&](double *ptr, int count) -> int {
    return multi ? b200mf_comm_allreduce_sum(c, ptr, count, st) : B200MF_OK;
  };
  int rc;
  std::memset(result, 0, sizeof(*result));
  // startup: r = b - A x unless x == 0 (solver_cg.h:640-652)
  if ((rc = b200mf_vec_dot_device(s.number, x, x, n, sc + 32, st)) != B200MF_OK) return rc;
  if ((rc = allreduce(sc + 32, 1)) != B200MF_OK) return rc;
  B200MF_CUDA_CHECK(cudaMemcpyAsync(hp, sc + 32, sizeof(double), cudaMemcpyDeviceToHost, st));
  B200MF_CUDA_CHECK(cudaStreamSynchronize(st));
  const bool x_zero = hp[0] == 0.0;
  if (!x_zero) {
    if ((rc = dist_vmult_impl(s, *p, *op, v, x, st, nullptr)) != B200MF_OK) return rc;
    result->operator_applications++;
  }
  if ((rc = b200mf_cg_init(s.number, r, pv, b, x_zero ? nullptr : v, d, n, sc, st)) != B200MF_OK) return rc;
  if ((rc = allreduce(sc + 8 + 1, 2)) != B200MF_OK) return rc;
  B200MF_CUDA_CHECK(cudaMemcpyAsync(hp, sc + 8, 3 * sizeof(double), cudaMemcpyDeviceToHost, st));
  B200MF_CUDA_CHECK(cudaStreamSynchronize(st));
  double res = std::sqrt(hp[1]);
  result->initial_residual = result->residual = res;
  const int check_every = sd->check_every > 0 ? sd->check_every : 1;
  int it = 0;
  bool converged = res <= sd->tolerance;
  while (!converged && it < sd->max_iterations) {
    ++it;
    double *cur = sc + 8 * (it % 3), *nxt = sc + 8 * ((it + 1) % 3);
    if ((rc = dist_vmult_impl(s, *p, *op, v, pv, st, cur)) != B200MF_OK) return rc;
    result->operator_applications++;
    if ((rc = allreduce(cur, 1)) != B200MF_OK) return rc;
    if ((rc = b200mf_cg_post(s.number, r, v, d, n, sc, it, st)) != B200MF_OK) return rc;
    if ((rc = allreduce(nxt + 1, 2)) != B200MF_OK) return rc;
    bool done = false;
    if (it % check_every == 0 || it >= sd->max_iterations) {
      B200MF_CUDA_CHECK(cudaMemcpyAsync(hp, nxt, 3 * sizeof(double), cudaMemcpyDeviceToHost, st));
      B200MF_CUDA_CHECK(cudaStreamSynchronize(st));
      res = std::sqrt(std::fabs(hp[1]));
      converged = res <= sd->tolerance;
      done = converged || it >= sd->max_iterations || std::isnan(res);
    }
    if (done) {
      if ((rc = b200mf_cg_final(s.number, x, pv, n, sc, it, st)) != B200MF_OK) return rc;
      break;
    }
    if ((rc = b200mf_cg_pre(s.number, x, pv, r, d, n, sc, it, st)) != B200MF_OK) return rc;
  }
  B200MF_CUDA_CHECK(cudaStreamSynchronize(st));
  result->iterations = it;
  result->residual = res;
  if (!converged) {
    set_error("CG did not converge: residual %.3e after %d iterations", res, it);
    return B200MF_ERR_NOCONVERGENCE;
  }
  return B200MF_OK;
}

} // extern "C"

// hooks for the solvers that run on one rank or on a partition alike (solver_impl.cuh, multigrid.cu)
namespace b200mf {
int level_vmult(const Setup &s, const b200mf_partitioner *p, const b200mf_operator &op, void *dst, void *src,
                cudaStream_t st, double *dot_accum) {
  return p ? dist_vmult_impl(s, *p, op, dst, src, st, dot_accum) : vmult_impl(s, op, dst, src, st, dot_accum);
}
int level_allreduce(const b200mf_partitioner *p, double *device_values, int count, cudaStream_t st) {
  if (!p || p->n_ranks == 1) return B200MF_OK;
  return b200mf_comm_allreduce_sum(p->comm, device_values, count, st);
}
uint64_t level_first_owned(const b200mf_partitioner *p) { return p ? p->first_owned : 0; }
bool level_is_distributed(const b200mf_partitioner *p) { return p && p->n_ranks > 1; }
} // namespace b200mf
