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

#include "mesh_internal.h"

namespace b200mf {

// offsets in [0,p]^dim of the dofs of FE_Q(p) in hierarchical order
std::vector<std::array<int, 3>> mesh_hierarchic_offsets(int dim, int p) {
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

namespace {
inline std::vector<std::array<int, 3>> hierarchic_offsets(int dim, int p) { return mesh_hierarchic_offsets(dim, p); }

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
  if (d->dof_numbering == 1) {
    // DoFRenumbering::lexicographic: support points sorted by the last coordinate first, i.e. the new
    // number of a dof is its rank in the lattice order x fastest (every lattice point of a uniformly
    // refined cube carries a dof)
    int32_t r = 0;
    for (uint64_t idx = 0; idx < lattice_size; ++idx)
      if (lattice[idx] >= 0) lattice[idx] = r++;
  }
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
  if (m->partitioned) {
    d.n_owned_dofs = m->n_owned; d.n_ghost_dofs = m->n_ghost; d.n_cells_interior = m->n_cells_interior;
  }
  d.local_to_global = m->l2g.data(); d.geometry = B200MF_GEOMETRY_Q1_VERTICES;
  d.cell_vertices = m->vertices.data();
  d.constrained_dofs = m->boundary.data(); d.n_constrained_dofs = m->boundary.size();
  if (!m->cell_mask.empty()) d.constraint_mask = m->cell_mask.data();
  return b200mf_setup_create(&d, out);
}

} // extern "C"

// ---------------------------------------------------------------------------------------
// Partitioned generator
// ---------------------------------------------------------------------------------------
namespace b200mf {
namespace {

struct Box { int lo[3], ext[3]; }; // in cells

struct PartGeom {
  int dim, p, n, levels, Nloc; // Nloc = cells per direction inside one coarse cell
  int coarse[3];
  int n_ranks;
  uint64_t cells_per_coarse, n_cells_global, chunk;
  std::vector<Box> boxes;
  std::vector<uint64_t> offsets; // [n_ranks+1]
  int Ncells[3];                 // global cells per direction
};

// global integer cell coordinates of global cell index c (coarse lexicographic x Morton)
inline void global_cell_coords(const PartGeom &g, uint64_t c, int ijk[3]) {
  const uint64_t q = c / g.cells_per_coarse, m = c % g.cells_per_coarse;
  int cc[3] = {(int)(q % g.coarse[0]), (int)((q / g.coarse[0]) % g.coarse[1]),
               (int)(q / ((uint64_t)g.coarse[0] * g.coarse[1]))};
  int loc[3] = {0, 0, 0};
  for (int level = 0; level < g.levels; ++level) {
    const unsigned child = (unsigned)(m >> (g.dim * level)) & ((1u << g.dim) - 1);
    for (int k = 0; k < g.dim; ++k) loc[k] |= ((child >> k) & 1) << level;
  }
  for (int k = 0; k < 3; ++k) ijk[k] = (k < g.dim) ? cc[k] * g.Nloc + loc[k] : 0;
}

// owned lattice interval of a box along direction k: [lo*p + (lo>0), (lo+ext)*p]
inline uint64_t owned_count(const PartGeom &g, const Box &b) {
  uint64_t c = 1;
  for (int k = 0; k < g.dim; ++k) c *= (uint64_t)b.ext[k] * g.p + (b.lo[k] == 0 ? 1 : 0);
  return c;
}

inline bool owns_point(const PartGeom &g, const Box &b, const int64_t P[3]) {
  for (int k = 0; k < g.dim; ++k) {
    const int64_t lo = (int64_t)b.lo[k] * g.p + (b.lo[k] > 0 ? 1 : 0), hi = (int64_t)(b.lo[k] + b.ext[k]) * g.p;
    if (P[k] < lo || P[k] > hi) return false;
  }
  return true;
}

// first-touch numbering of the points owned by rank s over the closure lattice of its box;
// lattice index = sum_k (P_k - lo_k*p) * stride_k with extents ext_k*p+1.  Returns count.
uint64_t number_box(const PartGeom &g, int s, const std::vector<std::array<int, 3>> &hier,
                    std::vector<int32_t> &latt) {
  const Box &b = g.boxes[s];
  uint64_t ext[3] = {1, 1, 1}, stride[3] = {1, 1, 1}, size = 1;
  for (int k = 0; k < g.dim; ++k) { ext[k] = (uint64_t)b.ext[k] * g.p + 1; stride[k] = size; size *= ext[k]; }
  latt.assign(size, -1);
  const int npc = (int)hier.size();
  std::vector<uint64_t> hier_lin(npc);
  // points on a low face that has a neighbour are not owned: mark them -2 lazily by coordinate
  int32_t next = 0;
  for (uint64_t c = (uint64_t)s * g.chunk; c < (uint64_t)(s + 1) * g.chunk; ++c) {
    int ijk[3];
    global_cell_coords(g, c, ijk);
    uint64_t base = 0;
    bool low_face[3];
    for (int k = 0; k < g.dim; ++k) {
      const int l = ijk[k] - b.lo[k];
      base += (uint64_t)l * g.p * stride[k];
      low_face[k] = (l == 0 && b.lo[k] > 0);
    }
    for (int h = 0; h < npc; ++h) {
      bool foreign = false;
      uint64_t off = 0;
      for (int k = 0; k < g.dim; ++k) {
        off += (uint64_t)hier[h][k] * stride[k];
        foreign |= (low_face[k] && hier[h][k] == 0);
      }
      if (foreign) continue;
      int32_t &slot = latt[base + off];
      if (slot < 0) slot = next++;
    }
  }
  return (uint64_t)next;
}

} // namespace
} // namespace b200mf

extern "C" {

int b200mf_mesh_create_partitioned(const b200mf_partition_desc *pd, b200mf_mesh **out) {
  B200MF_REQUIRE(pd && out, "null argument");
  const b200mf_mesh_desc *d = &pd->mesh;
  B200MF_REQUIRE(d->dim == 2 || d->dim == 3, "dim must be 2 or 3");
  B200MF_REQUIRE(d->degree >= 1 && d->degree <= 8, "degree must be in 1..8");
  B200MF_REQUIRE(d->cell_order == B200MF_MESH_MORTON, "partitioned meshes use Morton order");
  const int N = d->cells_per_direction, dim = d->dim, p = d->degree, n = p + 1;
  B200MF_REQUIRE(N >= 1 && (N & (N - 1)) == 0, "cells_per_direction must be a power of two");
  B200MF_REQUIRE(pd->n_ranks >= 1 && pd->rank >= 0 && pd->rank < pd->n_ranks, "bad rank / n_ranks");
  PartGeom g;
  g.dim = dim; g.p = p; g.n = n; g.Nloc = N; g.n_ranks = pd->n_ranks;
  g.levels = 0;
  while ((1 << g.levels) < N) ++g.levels;
  uint64_t n_coarse = 1;
  for (int k = 0; k < 3; ++k) {
    g.coarse[k] = (k < dim && pd->coarse[k] > 0) ? pd->coarse[k] : 1;
    n_coarse *= g.coarse[k];
    g.Ncells[k] = k < dim ? g.coarse[k] * N : 1;
  }
  g.cells_per_coarse = 1;
  for (int k = 0; k < dim; ++k) g.cells_per_coarse *= N;
  g.n_cells_global = n_coarse * g.cells_per_coarse;
  B200MF_REQUIRE(g.n_cells_global % pd->n_ranks == 0, "cell count not divisible by n_ranks");
  g.chunk = g.n_cells_global / pd->n_ranks;
  // boxes of all ranks from their first and last cell (chunks must be boxes)
  g.boxes.resize(pd->n_ranks);
  for (int s = 0; s < pd->n_ranks; ++s) {
    int a[3], b[3];
    global_cell_coords(g, (uint64_t)s * g.chunk, a);
    global_cell_coords(g, (uint64_t)(s + 1) * g.chunk - 1, b);
    uint64_t vol = 1;
    for (int k = 0; k < 3; ++k) {
      g.boxes[s].lo[k] = a[k];
      g.boxes[s].ext[k] = k < dim ? b[k] - a[k] + 1 : 1;
      B200MF_REQUIRE(g.boxes[s].ext[k] >= 1, "Morton chunks of this rank count are not boxes");
      vol *= g.boxes[s].ext[k];
    }
    B200MF_REQUIRE(vol == g.chunk, "Morton chunks of this rank count are not boxes "
                                   "(use n_ranks = coarse cells, or one coarse cell and 2^k ranks)");
  }
  g.offsets.assign(pd->n_ranks + 1, 0);
  for (int s = 0; s < pd->n_ranks; ++s) g.offsets[s + 1] = g.offsets[s] + owned_count(g, g.boxes[s]);
  const int me = pd->rank;
  const Box &mb = g.boxes[me];
  B200MF_REQUIRE(g.offsets[me + 1] - g.offsets[me] < 0x7fffffffull, "too many dofs per rank");

  b200mf_mesh *m = new b200mf_mesh();
  m->desc = *d;
  m->partitioned = true;
  m->n_cells = g.chunk;
  m->n_global_cells = g.n_cells_global;
  m->n_global_dofs = g.offsets[pd->n_ranks];
  m->rank_offsets = g.offsets;
  m->first_owned = g.offsets[me];
  int npc = 1;
  for (int k = 0; k < dim; ++k) npc *= n;
  m->dofs_per_cell = npc;
  const auto hier = hierarchic_offsets(dim, p);

  // ---- own numbering
  std::vector<int32_t> own;
  m->n_owned = number_box(g, me, hier, own);
  m->n_dofs = m->n_owned;
  if (m->n_owned != g.offsets[me + 1] - g.offsets[me]) {
    delete m;
    set_error("internal error: owned dof count mismatch");
    return B200MF_ERR_INVALID;
  }

  // ---- halo lattice: region of cells [lo-h, hi+h) clipped to the domain; local index per point
  const int halo = pd->ghost_mode == B200MF_GHOSTS_RELEVANT ? 1 : 0;
  int64_t rlo[3] = {0, 0, 0}, rext[3] = {1, 1, 1};
  uint64_t rstride[3] = {1, 1, 1}, rsize = 1;
  for (int k = 0; k < dim; ++k) {
    const int clo = std::max(0, mb.lo[k] - halo), chi = std::min(g.Ncells[k], mb.lo[k] + mb.ext[k] + halo);
    rlo[k] = (int64_t)clo * p;
    rext[k] = (int64_t)(chi - clo) * p + 1;
    rstride[k] = rsize;
    rsize *= (uint64_t)rext[k];
  }
  std::vector<int32_t> local(rsize, -1);
  uint64_t ostride[3] = {1, 1, 1}, oext[3] = {1, 1, 1};
  {
    uint64_t sz = 1;
    for (int k = 0; k < dim; ++k) { oext[k] = (uint64_t)mb.ext[k] * p + 1; ostride[k] = sz; sz *= oext[k]; }
  }
  // ghost candidates grouped by owner
  struct GhostPoint { uint64_t rpos; int64_t P[3]; };
  std::vector<std::vector<GhostPoint>> by_owner(pd->n_ranks);
  {
    int64_t P[3] = {0, 0, 0};
    for (uint64_t idx = 0; idx < rsize; ++idx) {
      uint64_t r = idx;
      bool in_own_closure = true, touched = true;
      uint64_t opos = 0;
      for (int k = 0; k < dim; ++k) {
        P[k] = rlo[k] + (int64_t)(r % (uint64_t)rext[k]);
        r /= (uint64_t)rext[k];
        const int64_t l = P[k] - (int64_t)mb.lo[k] * p;
        if (l < 0 || l >= (int64_t)oext[k]) in_own_closure = false;
        else opos += (uint64_t)l * ostride[k];
      }
      touched = in_own_closure;
      if (in_own_closure && own[opos] >= 0) { local[idx] = own[opos]; continue; }
      if (pd->ghost_mode == B200MF_GHOSTS_TOUCHED && !touched) continue;
      // a ghost: find the owner
      int owner = -1;
      for (int s = 0; s < pd->n_ranks; ++s)
        if (s != me && owns_point(g, g.boxes[s], P)) { owner = s; break; }
      if (owner < 0) { delete m; set_error("internal error: ghost point without owner"); return B200MF_ERR_INVALID; }
      GhostPoint gp; gp.rpos = idx; gp.P[0] = P[0]; gp.P[1] = P[1]; gp.P[2] = P[2];
      by_owner[owner].push_back(gp);
    }
  }
  // global numbers of the ghosts: replay the owner's numbering
  std::vector<std::pair<uint64_t, uint64_t>> ghosts; // (global, rpos)
  {
    std::vector<int32_t> other;
    for (int s = 0; s < pd->n_ranks; ++s) {
      if (by_owner[s].empty()) continue;
      number_box(g, s, hier, other);
      const Box &ob = g.boxes[s];
      uint64_t st[3] = {1, 1, 1}, sz = 1;
      for (int k = 0; k < dim; ++k) { st[k] = sz; sz *= (uint64_t)ob.ext[k] * p + 1; }
      for (const GhostPoint &gp : by_owner[s]) {
        uint64_t pos = 0;
        for (int k = 0; k < dim; ++k) pos += (uint64_t)(gp.P[k] - (int64_t)ob.lo[k] * p) * st[k];
        ghosts.emplace_back(g.offsets[s] + (uint64_t)other[pos], gp.rpos);
      }
      by_owner[s].clear();
      by_owner[s].shrink_to_fit();
    }
  }
  std::sort(ghosts.begin(), ghosts.end());
  m->n_ghost = ghosts.size();
  B200MF_REQUIRE(m->n_owned + m->n_ghost < 0x7fffffffull, "too many local dofs");
  m->ghost_global.resize(m->n_ghost);
  for (uint64_t i = 0; i < m->n_ghost; ++i) {
    m->ghost_global[i] = ghosts[i].first;
    local[ghosts[i].second] = (int32_t)(m->n_owned + i);
  }
  if (pd->want_lattice_ids) {
    // global lattice id = P0 + L0*(P1 + L1*P2)
    uint64_t L[3] = {1, 1, 1};
    for (int k = 0; k < dim; ++k) L[k] = (uint64_t)g.Ncells[k] * p + 1;
    m->lattice_ids.assign(m->n_owned + m->n_ghost, 0);
    for (uint64_t idx = 0; idx < rsize; ++idx) {
      if (local[idx] < 0) continue;
      uint64_t r = idx, id = 0, mul = 1;
      for (int k = 0; k < dim; ++k) {
        id += (uint64_t)(rlo[k] + (int64_t)(r % (uint64_t)rext[k])) * mul;
        mul *= L[k];
        r /= (uint64_t)rext[k];
      }
      m->lattice_ids[local[idx]] = id;
    }
  }
  // boundary (Dirichlet) dofs among the owned ones, and a flag per local dof
  std::vector<uint8_t> is_boundary;
  if (d->dirichlet_boundary) {
    is_boundary.assign(m->n_owned + m->n_ghost, 0);
    for (uint64_t idx = 0; idx < rsize; ++idx) {
      if (local[idx] < 0) continue;
      uint64_t r = idx;
      bool bnd = false;
      for (int k = 0; k < dim; ++k) {
        const int64_t P = rlo[k] + (int64_t)(r % (uint64_t)rext[k]);
        r /= (uint64_t)rext[k];
        bnd |= (P == 0 || P == (int64_t)g.Ncells[k] * p);
      }
      if (bnd) is_boundary[local[idx]] = 1;
    }
    for (uint64_t i = 0; i < m->n_owned; ++i)
      if (is_boundary[i]) m->boundary.push_back((uint32_t)i);
  }

  // ---- cell lists: interior cells (no ghost dof) first, Morton order kept inside each class.
  // In 3D the classification is done for whole aligned blocks of b^3 consecutive cells (the
  // bricks of brick_kernel.cuh, b = brick_edge(degree)): a block is "interior" if none of its
  // cells touches a ghost dof, so both classes keep their bricks intact.
  const uint64_t nc = g.chunk;
  std::vector<uint64_t> order(nc);
  {
    uint64_t W = 1;
    if (dim == 3) {
      const uint64_t b = (uint64_t)brick_edge(p);
      if (nc % (b * b * b) == 0 && ((uint64_t)me * g.chunk) % (b * b * b) == 0) W = b * b * b;
    }
    std::vector<uint8_t> touches(nc, 0);
#pragma omp parallel for schedule(static)
    for (int64_t c = 0; c < (int64_t)nc; ++c) {
      int ijk[3];
      global_cell_coords(g, (uint64_t)me * g.chunk + (uint64_t)c, ijk);
      bool touches_ghost = false;
      if (m->n_ghost) {
        // a cell touches a non-owned dof iff one of its 2^dim corner points is not owned
        // (ghost dofs sit on partition interfaces, which are unions of cell faces)
        for (int v = 0; v < (1 << dim) && !touches_ghost; ++v) {
          uint64_t pos = 0;
          for (int k = 0; k < dim; ++k)
            pos += (uint64_t)(((int64_t)ijk[k] + ((v >> k) & 1)) * p - rlo[k]) * rstride[k];
          touches_ghost = local[pos] >= (int32_t)m->n_owned;
        }
        // faces/edges on a low interface without a ghost corner cannot exist for box partitions
      }
      touches[c] = touches_ghost ? 1 : 0;
    }
    std::vector<uint64_t> boundary_cells;
    uint64_t ni = 0;
    for (uint64_t blk = 0; blk < nc / W; ++blk) {
      bool any = false;
      for (uint64_t c = blk * W; c < (blk + 1) * W; ++c) any |= touches[c] != 0;
      for (uint64_t c = blk * W; c < (blk + 1) * W; ++c) {
        if (any) boundary_cells.push_back(c);
        else order[ni++] = c;
      }
    }
    m->n_cells_interior = ni;
    for (uint64_t c : boundary_cells) order[ni++] = c;
  }
  m->cell_position = order;
  std::vector<uint64_t> lex_lin(npc);
  for (int i = 0; i < npc; ++i) {
    const int a[3] = {i % n, (i / n) % n, i / (n * n)};
    lex_lin[i] = a[0] * rstride[0] + a[1] * rstride[1] + (dim == 3 ? a[2] * rstride[2] : 0);
  }
  m->l2g.resize(nc * npc);
  const int nv = 1 << dim;
  m->vertices.resize(nc * nv * dim);
  const double h = (d->right - d->left) / N;
  const double pi = 3.14159265358979323846;
  const bool mark = d->mark_constrained_l2g && d->dirichlet_boundary;
#pragma omp parallel for schedule(static)
  for (int64_t i = 0; i < (int64_t)nc; ++i) {
    const uint64_t c = order[i];
    int ijk[3];
    global_cell_coords(g, (uint64_t)me * g.chunk + c, ijk);
    uint64_t base = 0;
    for (int k = 0; k < dim; ++k) base += (uint64_t)((int64_t)ijk[k] * p - rlo[k]) * rstride[k];
    uint32_t *row = m->l2g.data() + (uint64_t)i * npc;
    for (int j = 0; j < npc; ++j) {
      uint32_t li = (uint32_t)local[base + lex_lin[j]];
      if (mark && is_boundary[li]) li |= B200MF_L2G_CONSTRAINED;
      row[j] = li;
    }
    for (int v = 0; v < nv; ++v) {
      double x[3] = {0, 0, 0};
      for (int k = 0; k < dim; ++k) x[k] = d->left + h * (ijk[k] + ((v >> k) & 1));
      if (d->deformation == B200MF_DEFORM_SINE) {
        // the displacement vanishes on the boundary of every coarse cube
        const double len = d->right - d->left;
        double s = d->deformation_amplitude * len;
        for (int k = 0; k < dim; ++k) s *= std::sin(pi * (x[k] - d->left) / len);
        for (int k = 0; k < dim; ++k) x[k] += s;
      }
      for (int k = 0; k < dim; ++k) m->vertices[((uint64_t)i * nv + v) * dim + k] = x[k];
    }
  }
  *out = m;
  return B200MF_OK;
}

int b200mf_mesh_partition_view_get(const b200mf_mesh *m, b200mf_partition_view *v) {
  B200MF_REQUIRE(m && v, "null argument");
  B200MF_REQUIRE(m->partitioned, "mesh was not created by b200mf_mesh_create_partitioned");
  v->n_global_dofs = m->n_global_dofs; v->n_global_cells = m->n_global_cells;
  v->first_owned_global = m->first_owned; v->n_owned = m->n_owned; v->n_ghost = m->n_ghost;
  v->n_cells_interior = m->n_cells_interior;
  v->rank_offsets = m->rank_offsets.data();
  v->ghost_global = m->ghost_global.data();
  v->lattice_ids = m->lattice_ids.empty() ? nullptr : m->lattice_ids.data();
  v->cell_morton_position = m->cell_position.empty() ? nullptr : m->cell_position.data();
  return B200MF_OK;
}

} // extern "C"
