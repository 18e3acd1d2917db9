// Multi-GPU CG from a C++ host, no Python: one process per GPU, each calling only the C ABI of
// libb200mf.so -- what a deal.II application with MPI does, with the MPI calls replaced by the two
// things they are used for here (this image has no MPI): the ranks learn (rank, world) from the
// environment (RANK / WORLD_SIZE / LOCAL_RANK, as torchrun or mpirun wrappers export them) and rank 0
// hands the 128-byte NCCL id to the others through a file (an MPI_Bcast in a real application).
//
//   partitioned mesh (b200mf_mesh_create_partitioned, p4est-style numbering)
//   -> b200mf_setup_create_from_mesh -> b200mf_comm_create -> b200mf_partitioner_create
//   -> b200mf_dist_vmult (self-check: the Laplacian annihilates constants across the partition)
//   -> b200mf_dist_compute_diagonal -> b200mf_dist_cg_solve (Jacobi)
//   -> the same system with step-37's multigrid as preconditioner: partitioned level meshes 1..r, FP32 level
//      setups + partitioners, child tables from b200mf_partition_view::cell_morton_position,
//      b200mf_mg_create -> b200mf_mg_dist_cg_solve
//
// Build:  g++ -std=c++17 -Iinclude -I/usr/local/cuda/include examples/dist_cg.cc -Ldealii_b200 -lb200mf
//             -L/usr/local/cuda/lib64 -lcudart -o dist_cg
// Run:    for r in 0 1; do RANK=$r WORLD_SIZE=2 LOCAL_RANK=$r B200MF_ID_FILE=/tmp/id ./dist_cg & done; wait
#include <cuda_runtime_api.h>
#include <unistd.h>

#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <initializer_list>
#include <string>
#include <vector>

#include "b200mf.h"

#define CHECK(call)                                                                      \
  do {                                                                                   \
    int rc__ = (call);                                                                   \
    if (rc__ != B200MF_OK) {                                                             \
      std::fprintf(stderr, "rank %d: %s failed: %s\n", rank, #call, b200mf_last_error()); \
      return 1;                                                                          \
    }                                                                                    \
  } while (0)

static int env_int(const char *name, int fallback) {
  const char *v = std::getenv(name);
  return v ? std::atoi(v) : fallback;
}

int main() {
  const int rank = env_int("RANK", 0), world = env_int("WORLD_SIZE", 1), local = env_int("LOCAL_RANK", rank);
  const int degree = env_int("B200MF_DEGREE", 4), refinements = env_int("B200MF_REFINEMENTS", 4);
  if (cudaSetDevice(local) != cudaSuccess) {
    std::fprintf(stderr, "rank %d: no CUDA device %d (the engine has no CPU fallback)\n", rank, local);
    return 1;
  }
  // ---- communicator: rank 0 creates the id, the others read it (MPI_Bcast in a deal.II application)
  unsigned char id[B200MF_UNIQUE_ID_BYTES];
  const std::string id_file = std::getenv("B200MF_ID_FILE") ? std::getenv("B200MF_ID_FILE") : "/tmp/b200mf_nccl_id";
  if (rank == 0) {
    CHECK(b200mf_comm_get_unique_id(id));
    const std::string tmp = id_file + ".tmp";
    FILE *f = std::fopen(tmp.c_str(), "wb");
    std::fwrite(id, 1, sizeof id, f);
    std::fclose(f);
    std::rename(tmp.c_str(), id_file.c_str());
  } else {
    FILE *f = nullptr;
    for (int tries = 0; tries < 600 && (f = std::fopen(id_file.c_str(), "rb")) == nullptr; ++tries) usleep(100000);
    if (!f || std::fread(id, 1, sizeof id, f) != sizeof id) {
      std::fprintf(stderr, "rank %d: cannot read the NCCL id from %s\n", rank, id_file.c_str());
      return 1;
    }
    std::fclose(f);
  }
  b200mf_comm *comm = nullptr;
  CHECK(b200mf_comm_create(id, world, rank, &comm));

  // ---- this rank's part of the mesh: one cube per rank
  b200mf_partition_desc pd;
  std::memset(&pd, 0, sizeof pd);
  pd.mesh.dim = 3; pd.mesh.degree = degree; pd.mesh.cells_per_direction = 1 << refinements;
  pd.mesh.cell_order = B200MF_MESH_MORTON; pd.mesh.left = 0.0; pd.mesh.right = 1.0;
  pd.mesh.dirichlet_boundary = 1;
  // ONE cube cut into world = 2^k Morton chunks: the global problem does not depend on the rank count, so
  // the numbers printed below must agree between WORLD_SIZE = 1, 2, 4, 8
  if (world != 1 && world != 2 && world != 4 && world != 8) { std::fprintf(stderr, "WORLD_SIZE must be 1, 2, 4 or 8\n"); return 1; }
  for (int k = 0; k < 3; ++k) pd.coarse[k] = 1;
  pd.n_ranks = world; pd.rank = rank; pd.ghost_mode = B200MF_GHOSTS_TOUCHED;
  b200mf_mesh *mesh = nullptr;
  CHECK(b200mf_mesh_create_partitioned(&pd, &mesh));
  b200mf_partition_view pv;
  CHECK(b200mf_mesh_partition_view_get(mesh, &pv));
  b200mf_mesh_view mv;
  CHECK(b200mf_mesh_view_get(mesh, &mv));
  b200mf_setup *setup = nullptr;
  CHECK(b200mf_setup_create_from_mesh(mesh, B200MF_F64, &setup));
  b200mf_partitioner *part = nullptr;
  CHECK(b200mf_partitioner_create(comm, pv.rank_offsets, pv.ghost_global, pv.n_ghost, B200MF_F64, &part));

  const size_t n = pv.n_owned, nt = pv.n_owned + pv.n_ghost;
  double *x, *b, *y, *diag, *scal;
  for (double **v : {&x, &b, &y, &diag}) cudaMalloc(reinterpret_cast<void **>(v), nt * 8);
  cudaMalloc(reinterpret_cast<void **>(&scal), 64);
  cudaMemset(x, 0, nt * 8); cudaMemset(b, 0, nt * 8); cudaMemset(y, 0, nt * 8);
  const b200mf_operator op = {nullptr, nullptr, 1.0, 0.0};

  // ---- self-check of the distributed vmult: the energy src^T A src (src = 1 on the unconstrained dofs) is a
  // property of the global problem, not of the partition: tests/test_dist_example_gpu.py compares the value, the
  // CG iteration count and |x| between rank counts.
  CHECK(b200mf_vec_set(B200MF_F64, b, 1.0, n, nullptr));
  CHECK(b200mf_set_constrained_values(setup, b, 0.0, nullptr));
  CHECK(b200mf_dist_vmult(setup, part, &op, y, b, nullptr));
  cudaMemset(scal, 0, 64);
  CHECK(b200mf_vec_dot_device(B200MF_F64, y, b, n, scal, nullptr)); // 1^T A 1 over the unconstrained dofs
  CHECK(b200mf_comm_allreduce_sum(comm, scal, 1, nullptr));
  double energy = 0;
  cudaMemcpy(&energy, scal, 8, cudaMemcpyDeviceToHost);

  // ---- CG + Jacobi: rhs = 1 on unconstrained dofs, tolerance 1e-8 |b|
  CHECK(b200mf_dist_compute_diagonal(setup, part, &op, diag, nullptr));
  {
    std::vector<double> h(n);
    cudaMemcpy(h.data(), diag, n * 8, cudaMemcpyDeviceToHost);
    for (auto &v : h) v = 1.0 / v;
    cudaMemcpy(diag, h.data(), n * 8, cudaMemcpyHostToDevice);
  }
  cudaMemset(scal, 0, 64);
  CHECK(b200mf_vec_dot_device(B200MF_F64, b, b, n, scal, nullptr));
  CHECK(b200mf_comm_allreduce_sum(comm, scal, 1, nullptr));
  double bb = 0;
  cudaMemcpy(&bb, scal, 8, cudaMemcpyDeviceToHost);
  b200mf_solver_desc sd;
  std::memset(&sd, 0, sizeof sd);
  sd.preconditioner = B200MF_PRECOND_JACOBI; sd.inverse_diagonal = diag;
  sd.tolerance = 1e-8 * std::sqrt(bb); sd.max_iterations = 5000; sd.check_every = 1;
  b200mf_solver_result res;
  CHECK(b200mf_dist_cg_solve(setup, part, &op, &sd, x, b, &res, nullptr));
  cudaMemset(scal, 0, 64);
  CHECK(b200mf_vec_dot_device(B200MF_F64, x, x, n, scal, nullptr));
  CHECK(b200mf_comm_allreduce_sum(comm, scal, 1, nullptr));
  double xx = 0;
  cudaMemcpy(&xx, scal, 8, cudaMemcpyDeviceToHost);
  if (rank == 0)
    std::printf("dist_cg: %d ranks, %llu global dofs, 1^T A 1 = %.12e, CG iterations %d, |x| = %.12e, residual %.3e\n", world,
                (unsigned long long)pv.n_global_dofs, energy, res.iterations, std::sqrt(xx), res.residual);

  // ---- the same solve preconditioned by the geometric multigrid of step-37 on the partitioned hierarchy.
  // Level 1 (8 cells) is the coarsest level every rank count up to 8 can share; a rank's cells of level l+1 are
  // the children of its cells of level l (Morton chunks nest), found through cell_morton_position.
  {
    const int min_level = 1;
    std::vector<b200mf_mesh *> level_meshes;
    std::vector<b200mf_setup *> level_setups;
    std::vector<b200mf_partitioner *> level_parts;
    std::vector<b200mf_operator> level_ops;
    std::vector<std::vector<uint64_t>> positions;
    for (int level = min_level; level <= refinements; ++level) {
      b200mf_partition_desc ld = pd;
      ld.mesh.cells_per_direction = 1 << level;
      ld.mesh.mark_constrained_l2g = 1; // MatrixFreeOperators::Base: constrained dofs eliminated on the levels
      b200mf_mesh *lm = nullptr;
      CHECK(b200mf_mesh_create_partitioned(&ld, &lm));
      b200mf_partition_view lv;
      CHECK(b200mf_mesh_partition_view_get(lm, &lv));
      b200mf_mesh_view lmv;
      CHECK(b200mf_mesh_view_get(lm, &lmv));
      b200mf_setup *ls = nullptr;
      CHECK(b200mf_setup_create_from_mesh(lm, B200MF_F32, &ls));
      b200mf_partitioner *lp = nullptr;
      CHECK(b200mf_partitioner_create(comm, lv.rank_offsets, lv.ghost_global, lv.n_ghost, B200MF_F32, &lp));
      level_meshes.push_back(lm);
      level_setups.push_back(ls);
      level_parts.push_back(lp);
      level_ops.push_back(op);
      positions.emplace_back(lv.cell_morton_position, lv.cell_morton_position + lmv.n_cells);
    }
    std::vector<std::vector<uint32_t>> tables;
    std::vector<const uint32_t *> table_ptrs;
    for (size_t l = 0; l + 1 < positions.size(); ++l) {
      const auto &pc = positions[l], &pf = positions[l + 1];
      std::vector<uint32_t> inverse(pf.size());
      for (size_t i = 0; i < pf.size(); ++i) inverse[pf[i]] = (uint32_t)i;
      std::vector<uint32_t> table(pc.size() * 8);
      for (size_t i = 0; i < pc.size(); ++i)
        for (int k = 0; k < 8; ++k) table[i * 8 + k] = inverse[(pc[i] << 3) + k];
      tables.push_back(std::move(table));
    }
    for (const auto &t : tables) table_ptrs.push_back(t.data());
    b200mf_mg_desc md;
    std::memset(&md, 0, sizeof md);
    md.n_levels = (int)level_setups.size();
    md.levels = level_setups.data();
    md.operators = level_ops.data();
    md.child_cells = table_ptrs.empty() ? nullptr : table_ptrs.data();
    md.partitioners = level_parts.data();
    md.smoother_degree = 5; md.smoothing_range = 15.0; md.eig_cg_n_iterations = 10; md.coarse_tolerance = 1e-3;
    b200mf_mg *mg = nullptr;
    CHECK(b200mf_mg_create(&md, &mg, nullptr));
    cudaMemset(x, 0, nt * 8);
    b200mf_solver_result mres;
    CHECK(b200mf_mg_dist_cg_solve(mg, setup, part, &op, 1e-8 * std::sqrt(bb), 100, x, b, &mres, nullptr));
    cudaMemset(scal, 0, 64);
    CHECK(b200mf_vec_dot_device(B200MF_F64, x, x, n, scal, nullptr));
    CHECK(b200mf_comm_allreduce_sum(comm, scal, 1, nullptr));
    double xg = 0;
    cudaMemcpy(&xg, scal, 8, cudaMemcpyDeviceToHost);
    if (rank == 0)
      std::printf("dist_gmg: %d ranks, levels %d..%d, CG + multigrid iterations %d, |x| = %.12e, residual %.3e\n", world,
                  min_level, refinements, mres.iterations, std::sqrt(xg), mres.residual);
    b200mf_mg_destroy(mg);
    for (auto *p : level_parts) b200mf_partitioner_destroy(p);
    for (auto *ls : level_setups) b200mf_setup_destroy(ls);
    for (auto *lm : level_meshes) b200mf_mesh_destroy(lm);
  }
  b200mf_partitioner_destroy(part);
  b200mf_setup_destroy(setup);
  b200mf_mesh_destroy(mesh);
  b200mf_comm_destroy(comm);
  cudaFree(x); cudaFree(b); cudaFree(y); cudaFree(diag); cudaFree(scal);
  if (rank == 0) std::remove(id_file.c_str());
  return 0;
}
