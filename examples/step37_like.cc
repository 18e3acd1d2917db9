// Geometric multigrid from a C++ host over the C ABI, no Python and no deal.II: the solver of step-37
// (examples/step-37/step-37.cc:907-1075) -- FP64 CG preconditioned by one V-cycle with Chebyshev smoothers on
// FP32 levels, matrix-free transfer, Chebyshev coarse solver -- on meshes from the engine's generator.
//   b200mf_mesh_create (levels 0..r, Morton order = refine_global) -> b200mf_setup_create_from_mesh (FP32 per
//   level, FP64 for the system) -> b200mf_mg_create -> b200mf_mg_cg_solve
// Build:  g++ -std=c++17 -Iinclude examples/step37_like.cc -Ldealii_b200 -lb200mf -lcudart -o step37_like
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
#include <vector>

#include "b200mf.h"

#define CHECK(call)                                                                \
  do {                                                                             \
    if ((call) != B200MF_OK) {                                                     \
      std::fprintf(stderr, "%s failed: %s\n", #call, b200mf_last_error());         \
      return 1;                                                                    \
    }                                                                              \
  } while (0)

int main(int argc, char **argv) {
  const int degree = argc > 1 ? std::atoi(argv[1]) : 2, refinements = argc > 2 ? std::atoi(argv[2]) : 4;
  std::vector<b200mf_mesh *> meshes;
  std::vector<b200mf_setup *> levels;
  std::vector<b200mf_operator> operators;
  for (int level = 0; level <= refinements; ++level) {
    b200mf_mesh_desc md{};
    md.dim = 3; md.degree = degree; md.cells_per_direction = 1 << level; md.cell_order = B200MF_MESH_MORTON;
    md.left = 0.0; md.right = 1.0;
    md.dirichlet_boundary = 1; md.mark_constrained_l2g = 1; // MatrixFreeOperators::Base: constrained dofs eliminated
    b200mf_mesh *mesh = nullptr;
    CHECK(b200mf_mesh_create(&md, &mesh));
    b200mf_setup *setup = nullptr;
    CHECK(b200mf_setup_create_from_mesh(mesh, B200MF_F32, &setup));
    meshes.push_back(mesh);
    levels.push_back(setup);
    operators.push_back(b200mf_operator{nullptr, nullptr, 1.0, 0.0}); // Laplace
  }
  b200mf_mg_desc d{};
  d.n_levels = (int)levels.size();
  d.levels = levels.data();
  d.operators = operators.data();
  d.smoother_degree = 5; d.smoothing_range = 15.0; d.eig_cg_n_iterations = 10; // step-37.cc:965-975
  d.coarse_tolerance = 1e-3;
  b200mf_mg *mg = nullptr;
  CHECK(b200mf_mg_create(&d, &mg, nullptr));

  // the system operator in double on the finest mesh
  b200mf_setup *system = nullptr;
  CHECK(b200mf_setup_create_from_mesh(meshes.back(), B200MF_F64, &system));
  b200mf_mesh_view v{};
  CHECK(b200mf_mesh_view_get(meshes.back(), &v));
  const size_t n = v.n_dofs;
  double *x, *b, *r;
  cudaMalloc(reinterpret_cast<void **>(&x), n * 8);
  cudaMalloc(reinterpret_cast<void **>(&b), n * 8);
  cudaMalloc(reinterpret_cast<void **>(&r), n * 8);
  cudaMemset(x, 0, n * 8);
  CHECK(b200mf_vec_set(B200MF_F64, b, 1.0, n, nullptr));
  CHECK(b200mf_set_constrained_values(system, b, 0.0, nullptr));
  double bnorm = 0;
  CHECK(b200mf_vec_norm_2(B200MF_F64, b, n, &bnorm, nullptr));
  const b200mf_operator op = {nullptr, nullptr, 1.0, 0.0};
  b200mf_solver_result res{};
  CHECK(b200mf_mg_cg_solve(mg, system, &op, 1e-10 * bnorm, 100, x, b, &res, nullptr));
  // true residual
  CHECK(b200mf_vmult(system, &op, r, x, nullptr));
  CHECK(b200mf_vec_sadd(B200MF_F64, r, -1.0, 1.0, b, n, nullptr));
  double rnorm = 0;
  CHECK(b200mf_vec_norm_2(B200MF_F64, r, n, &rnorm, nullptr));
  std::printf("Q%d, %d levels, %llu DoFs: CG + multigrid converged in %d iterations, |b - A x| / |b| = %.3e\n", degree,
              d.n_levels, (unsigned long long)n, res.iterations, rnorm / bnorm);
  for (int l = 0; l < d.n_levels; ++l) {
    b200mf_mg_level_info info;
    CHECK(b200mf_mg_get_level_info(mg, l, &info));
    std::printf("  level %d: %llu DoFs, eigenvalue estimate [%.4f, %.4f], Chebyshev degree %d\n", l,
                (unsigned long long)info.n_dofs, info.eig_min, info.eig_max, info.degree);
  }
  const int ok = res.iterations <= 8 && rnorm <= 1e-9 * bnorm;
  cudaFree(x); cudaFree(b); cudaFree(r);
  b200mf_mg_destroy(mg);
  b200mf_setup_destroy(system);
  for (auto *s : levels) b200mf_setup_destroy(s);
  for (auto *m : meshes) b200mf_mesh_destroy(m);
  return ok ? 0 : 1;
}
