// Minimal C++ host program over the C ABI (through the header shim): the step-64 pattern
// (examples/step-64/step-64.cc:313-325, 595-622) -- reinit, vmult, compute_diagonal, SolverCG
// with Jacobi -- on a mesh from the engine's synthetic generator.  Build:
//   g++ -std=c++17 -Iinclude examples/step64_like.cc -Ldealii_b200 -lb200mf -lcudart -o step64_like
#include <cmath>
#include <cstdio>
#include <cuda_runtime.h>
#include <vector>

#include "b200mf_portable.hpp"

int main() {
  constexpr int dim = 3, degree = 3;
  b200mf_mesh_desc md{};
  md.dim = dim; md.degree = degree; md.cells_per_direction = 8; md.cell_order = B200MF_MESH_MORTON;
  md.left = 0.0; md.right = 1.0; md.dirichlet_boundary = 1;
  b200mf_mesh *mesh = nullptr;
  b200::check(b200mf_mesh_create(&md, &mesh));
  b200mf_mesh_view v{};
  b200::check(b200mf_mesh_view_get(mesh, &v));

  b200::ReinitData rd;
  rd.degree = degree; rd.n_cells = v.n_cells; rd.n_owned_dofs = v.n_dofs;
  rd.local_to_global = v.local_to_global; rd.cell_vertices = v.cell_vertices;
  rd.constrained_dofs = v.boundary_dofs; rd.n_constrained_dofs = v.n_boundary_dofs;
  b200::MatrixFree<dim, double> mf;
  mf.reinit(rd);
  b200::Operator<dim, double> A(mf, nullptr, nullptr, 1.0, /*mass_constant=*/1.0); // Helmholtz, a = 1

  const size_t n = mf.n_local_dofs();
  double *x, *b, *diag;
  cudaMalloc(&x, n * 8); cudaMalloc(&b, n * 8); cudaMalloc(&diag, n * 8);
  cudaMemset(x, 0, n * 8);
  b200::check(b200mf_vec_set(B200MF_F64, b, 1.0, n, nullptr));
  mf.set_constrained_values(0.0, b);
  A.compute_diagonal(diag);
  std::vector<double> h(n);
  cudaMemcpy(h.data(), diag, n * 8, cudaMemcpyDeviceToHost);
  for (auto &d : h) d = 1.0 / d;
  cudaMemcpy(diag, h.data(), n * 8, cudaMemcpyHostToDevice);

  b200::SolverControl control(1000, 1e-12 * std::sqrt(double(n)));
  b200::SolverCG<dim, double> cg(control);
  cg.solve(A, x, b, diag);
  std::printf("Solved in %d iterations, residual %.3e, %llu DoFs\n", control.last_step(),
              control.last_value(), (unsigned long long)v.n_dofs);
  cudaFree(x); cudaFree(b); cudaFree(diag);
  b200mf_mesh_destroy(mesh);
  return 0;
}
