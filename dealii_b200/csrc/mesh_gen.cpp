// Synthetic hyper-cube mesh + FE_Q DoF numbering generator (host).
// Stands in for GridGenerator::hyper_cube + refine_global / subdivided_hyper_cube and
// DoFHandler::distribute_dofs of the reference for meshes its host setup cannot hold:
//   * active cell order: Morton (children of a cell are created consecutively, child c at
//     offset (c&1, c>>1&1, c>>2&1); include/deal.II/base/geometry_info.h) or lexicographic
//     (subdivided_hyper_rectangle, source/grid/grid_generator.cc);
//   * first-touch numbering cell by cell, per cell vertices -> lines -> quads -> interior
//     (source/dofs/dof_handler_policy.cc:1676-1719) in FE_Q's hierarchical order
//     (include/deal.II/fe/fe_tools.templates.h:2987-3152);
//   * cell index lists re-ordered lexicographically as Portable::MatrixFree stores them
//     (matrix_free/portable_matrix_free.templates.h:292-298).
// Bit-identical to oracle/mesh.py at every size the oracle can hold (tests/test_mesh.py).
#include <algorithm>
#include <array>
#include <cmath>
#include <cstring>

#include "internal.h"

struct b200mf_mesh {
  b200mf_mesh_desc desc;
  uint64_t n_cells = 0, n_dofs = 0;
  int dofs_per_cell = 0;
  std::vector<uint32_t> l2g;
  std::vector<double> vertices;
  std::vector<uint32_t> boundary;
};

namespace b200mf {
namespace {

// offsets in [0,p]^dim of the dofs of FE_Q(p) in hierarchical order
std::vector<std::array<int, 3>> hierarchic_offsets(int dim, int p) {
  std::vector<std::array<int, 3>> out;
  auto add = [&](int x, int y, int z) { out.push_back(std::array<int, 3>{{x, y, z}}); };
  if (dim == 2) {
    for (int v = 0; v < 4; ++v) add((v & 1) * p, (v >> 1 & 1) * p, 0);
    for (int x : {0, p}) for (int a = 1; a < p; ++a) add(x, a, 0);
    for (int y : {0, p}) for (int a = 1; a < p; ++a) add(a, y, 0);
    for (int b = 1; b < p; ++b) for (int a = 1; a < p; ++a) add(a, b, 0);
  } else {
    for (int v = 0; v < 8; ++v) add((v & 1) * p, (v >> 1 & 1) * p, (v >> 2 & 1) * p);
    for (int z : {0, p}) {
      for (int x : {0, p}) for (int a = 1; a < p; ++a) add(x, a, z);
      for (int y : {0, p}) for (int a = 1; a < p; ++a) add(a, y, z);
    }
    const int xy[4][2] = {{0, 0}, {p, 0}, {0, p}, {p, p}};
    for (auto &c : xy) for (int a = 1; a < p; ++a) add(c[0], c[1], a);
    for (int x : {0, p}) for (int a = 1; a < p; ++a) for (int b = 1; b < p; ++b) add(x, b, a);
    for (int y : {0, p}) for (int a = 1; a < p; ++a) for (int b = 1; b < p; ++b) add(a, y, b);
    for (int z : {0, p}) for (int a = 1; a < p; ++a) for (int b = 1; b < p; ++b) add(b, a, z);
    for (int a = 1; a < p; ++a) for (int b = 1; b < p; ++b) for (int c = 1; c < p; ++c) add(c, b, a);
  }
  return out;
}

inline void cell_coords(const b200mf_mesh_desc &d, uint64_t c, int N, int ijk[3]) {
  ijk[0] = ijk[1] = ijk[2] = 0;
  if (d.cell_order == B200MF_MESH_MORTON) {
    for (int level = 0; (1 << level) < N; ++level) {
      const unsigned child = (unsigned)(c >> (d.dim * level)) & ((1u << d.dim) - 1);
      for (int k = 0; k < d.dim; ++k) ijk[k] |= ((child >> k) & 1) << level;
    }
  } else {
    for (int k = 0; k < d.dim; ++k) { ijk[k] = (int)(c % N); c /= N; }
  }
}

} // namespace
} // namespace b200mf

using namespace b200mf;

extern "C" {

int b200mf_mesh_create(const b200mf_mesh_desc *d, b200mf_mesh **out) {
  B200MF_REQUIRE(d && out, "null argument");
  B200MF_REQUIRE(d->dim == 2 || d->dim == 3, "dim must be 2 or 3");
  B200MF_REQUIRE(d->degree >= 1 && d->degree <= 8, "degree must be in 1..8");
  const int N = d->cells_per_direction, p = d->degree, n = p + 1, dim = d->dim;
  B200MF_REQUIRE(N >= 1, "cells_per_direction must be positive");
  if (d->cell_order == B200MF_MESH_MORTON)
    B200MF_REQUIRE((N & (N - 1)) == 0, "Morton order needs a power-of-two cell count per direction");
  const uint64_t L = (uint64_t)N * p + 1;
  uint64_t lattice_size = 1, n_cells = 1;
  for (int k = 0; k < dim; ++k) { lattice_size *= L; n_cells *= N; }
  B200MF_REQUIRE(lattice_size < 0x7fffffffull, "mesh too large for 32-bit local indices");

  b200mf_mesh *m = new b200mf_mesh();
  m->desc = *d;
  m->n_cells = n_cells;
  int npc = 1;
  for (int k = 0; k < dim; ++k) npc *= n;
  m->dofs_per_cell = npc;
  const uint64_t stride[3] = {1, L, L * L};

  std::vector<int32_t> lattice(lattice_size, -1);
  const auto hier = hierarchic_offsets(dim, p);
  std::vector<uint64_t> hier_lin(npc), lex_lin(npc);
  for (int h = 0; h < npc; ++h)
    hier_lin[h] = hier[h][0] * stride[0] + hier[h][1] * stride[1] + (dim == 3 ? hier[h][2] * stride[2] : 0);
  for (int i = 0; i < npc; ++i) {
    const int a[3] = {i % n, (i / n) % n, i / (n * n)};
    lex_lin[i] = a[0] * stride[0] + a[1] * stride[1] + (dim == 3 ? a[2] * stride[2] : 0);
  }
  // pass 1: first-touch numbering (inherently sequential)
  int32_t next = 0;
  std::vector<uint64_t> base(n_cells);
  for (uint64_t c = 0; c < n_cells; ++c) {
    int ijk[3];
    cell_coords(*d, c, N, ijk);
    uint64_t b = 0;
    for (int k = 0; k < dim; ++k) b += (uint64_t)ijk[k] * p * stride[k];
    base[c] = b;
    for (int h = 0; h < npc; ++h) {
      int32_t &slot = lattice[b + hier_lin[h]];
      if (slot < 0) slot = next++;
    }
  }
  m->n_dofs = (uint64_t)next;
  // boundary dofs
  std::vector<uint8_t> is_boundary;
  if (d->dirichlet_boundary) {
    is_boundary.assign(m->n_dofs, 0);
    for (uint64_t idx = 0; idx < lattice_size; ++idx) {
      if (lattice[idx] < 0) continue;
      uint64_t r = idx;
      bool bnd = false;
      for (int k = 0; k < dim; ++k) { const uint64_t a = r % L; r /= L; bnd |= (a == 0 || a == L - 1); }
      if (bnd) is_boundary[lattice[idx]] = 1;
    }
    for (uint64_t i = 0; i < m->n_dofs; ++i)
      if (is_boundary[i]) m->boundary.push_back((uint32_t)i);
  }
  // pass 2: lexicographic cell index lists + vertices
  m->l2g.resize(n_cells * npc);
  const int nv = 1 << dim;
  m->vertices.resize(n_cells * nv * dim);
  const double h = (d->right - d->left) / N, len = d->right - d->left;
  const double pi = 3.14159265358979323846;
  const bool mark = d->mark_constrained_l2g && d->dirichlet_boundary;
#pragma omp parallel for schedule(static)
  for (int64_t c = 0; c < (int64_t)n_cells; ++c) {
    uint32_t *row = m->l2g.data() + (uint64_t)c * npc;
    for (int i = 0; i < npc; ++i) {
      uint32_t g = (uint32_t)lattice[base[c] + lex_lin[i]];
      if (mark && is_boundary[g]) g |= B200MF_L2G_CONSTRAINED;
      row[i] = g;
    }
    int ijk[3];
    cell_coords(*d, (uint64_t)c, N, ijk);
    for (int v = 0; v < nv; ++v) {
      double x[3] = {0, 0, 0};
      for (int k = 0; k < dim; ++k) x[k] = d->left + h * (ijk[k] + ((v >> k) & 1));
      if (d->deformation == B200MF_DEFORM_SINE) {
        double s = d->deformation_amplitude * len;
        for (int k = 0; k < dim; ++k) s *= std::sin(pi * (x[k] - d->left) / len);
        for (int k = 0; k < dim; ++k) x[k] += s;
      }
      for (int k = 0; k < dim; ++k) m->vertices[((uint64_t)c * nv + v) * dim + k] = x[k];
    }
  }
  *out = m;
  return B200MF_OK;
}

int b200mf_mesh_view_get(const b200mf_mesh *m, b200mf_mesh_view *v) {
  B200MF_REQUIRE(m && v, "null argument");
  v->n_cells = m->n_cells; v->n_dofs = m->n_dofs; v->n_boundary_dofs = m->boundary.size();
  v->dofs_per_cell = m->dofs_per_cell; v->vertices_per_cell = 1 << m->desc.dim; v->dim = m->desc.dim;
  v->local_to_global = m->l2g.data(); v->cell_vertices = m->vertices.data();
  v->boundary_dofs = m->boundary.data();
  return B200MF_OK;
}

int b200mf_mesh_destroy(b200mf_mesh *m) { delete m; return B200MF_OK; }

int b200mf_setup_create_from_mesh(const b200mf_mesh *m, int number, b200mf_setup **out) {
  B200MF_REQUIRE(m && out, "null argument");
  b200mf_setup_desc d;
  std::memset(&d, 0, sizeof(d));
  d.dim = m->desc.dim; d.degree = m->desc.degree; d.n_q_points_1d = m->desc.degree + 1;
  d.number = number; d.n_cells = m->n_cells; d.n_owned_dofs = m->n_dofs; d.n_ghost_dofs = 0;
  d.local_to_global = m->l2g.data(); d.geometry = B200MF_GEOMETRY_Q1_VERTICES;
  d.cell_vertices = m->vertices.data();
  d.constrained_dofs = m->boundary.data(); d.n_constrained_dofs = m->boundary.size();
  return b200mf_setup_create(&d, out);
}

} // extern "C"
