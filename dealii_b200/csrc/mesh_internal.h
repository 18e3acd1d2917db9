// Shared between the host mesh generators (mesh_gen.cpp, mesh_adaptive.cpp).
#pragma once
#include <array>
#include <vector>

#include "internal.h"

struct b200mf_mesh {
  b200mf_mesh_desc desc;
  uint64_t n_cells = 0, n_dofs = 0;
  int dofs_per_cell = 0;
  std::vector<uint32_t> l2g;
  std::vector<double> vertices;
  std::vector<uint32_t> boundary; // constrained dofs: Dirichlet boundary (+ hanging nodes), sorted
  // partitioned meshes (b200mf_mesh_create_partitioned / _adaptive)
  bool partitioned = false;
  uint64_t n_global_dofs = 0, n_global_cells = 0, first_owned = 0, n_owned = 0, n_ghost = 0,
           n_cells_interior = 0;
  std::vector<uint64_t> rank_offsets, ghost_global, lattice_ids, cell_position;
  // adaptive meshes (b200mf_mesh_create_adaptive)
  std::vector<uint16_t> cell_mask;
  std::vector<uint32_t> active_index;
  std::vector<double> dof_coords;
  uint64_t n_hanging_dofs = 0;
};

namespace b200mf {
// offsets in [0,p]^dim of the dofs of FE_Q(p) in hierarchical order (vertices, lines, quads, hex)
std::vector<std::array<int, 3>> mesh_hierarchic_offsets(int dim, int p);
// Gauss-Lobatto support points of FE_Q(p) on [0, 1] (shape.cpp)
void build_fe_q_support_points(int degree, std::vector<double> &points);
} // namespace b200mf
