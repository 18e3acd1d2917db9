// Adaptively refined synthetic mesh + FE_Q numbering with hanging nodes (host): BASELINE
// configs[3] ("3D Q3 Poisson on a distributed (p4est) mesh with hanging nodes").
//
// Domain: coarse[0] x coarse[1] x coarse[2] unit cubes, one per rank, each refined globally
// log2(N) times; then every cell whose centre is closer than ball_radius (in cube edges) to the
// centre of its cube is refined once more -- one level of hanging nodes on the surface of a ball
// inside every cube (tests/matrix_free/matrix_vector_03.cc refines the same way).  All cubes carry
// the same number of cells, so p4est's equal-count cut of the Morton curve falls on the cube
// boundaries, which stay conforming: the ghost exchange is the one of the uniform mesh.
//
// What is restated from the reference (validated bit for bit against the reference itself on one
// rank: tests/golden/ref/c4_*.npz were produced by deal.II on the same mesh):
//   * active cell order: by level, cells of a level in the order of their creation, i.e. Morton
//     (Triangulation::execute_coarsening_and_refinement appends the children of a level's refined
//     cells in order; DoFHandler::active_cell_iterators walks level by level);
//   * DoF numbering: first touch over the active cells in that order, vertices -> lines -> quads ->
//     interior per cell (source/dofs/dof_handler_policy.cc:1676-1719); the dofs of a refined face
//     live on its child objects, those of the unrefined neighbour on the parent object, so both
//     sides of a hanging face carry their own dofs and only the vertices coincide;
//   * hanging-node description of Portable::MatrixFree: ConstraintKinds masks
//     (matrix_free/hanging_nodes_internal.h:40-60) and the index lists of refined cells with the
//     constrained faces / edges redirected to the coarse neighbour's dofs (:520-880);
//   * constrained dofs = Dirichlet boundary dofs + hanging-node dofs (AffineConstraints lines,
//     DoFTools::make_hanging_node_constraints).
// Partition rules as in mesh_gen.cpp: lowest rank owns interface dofs
// (dof_handler_policy.cc:3644-3760), local indices [owned | ghosts sorted by global index]; only
// the dofs the own cells touch are ghosts (B200MF_GHOSTS_TOUCHED).
#include <algorithm>
#include <array>
#include <cmath>
#include <cstring>
#include <map>

#include "internal.h"
#include "mesh_internal.h"

namespace b200mf {
namespace {

struct Cube {
  int dim, p, n, N, levels;
  int coarse[3], n_ranks;
  double radius;
  // bounding box of the refined cells, in coarse cells: [rlo, rhi)
  int rlo[3], rhi[3];
  uint64_t vext[3], cext[3], fext[3];
  uint64_t vstride[3], cstride[3], fstride[3], vsize, csize, fsize;
  uint64_t n_coarse_cells;
  std::vector<uint8_t> refined; // per coarse cell (Morton index)

  void morton_coords(uint64_t m, int ijk[3]) const {
    ijk[0] = ijk[1] = ijk[2] = 0;
    for (int level = 0; level < levels; ++level) {
      const unsigned child = (unsigned)(m >> (dim * level)) & ((1u << dim) - 1);
      for (int k = 0; k < dim; ++k) ijk[k] |= ((child >> k) & 1) << level;
    }
  }
  uint64_t morton_index(const int ijk[3]) const {
    uint64_t m = 0;
    for (int level = 0; level < levels; ++level)
      for (int k = 0; k < dim; ++k) m |= (uint64_t)((ijk[k] >> level) & 1) << (dim * level + k);
    return m;
  }
  bool is_refined(const int ijk[3]) const {
    for (int k = 0; k < dim; ++k)
      if (ijk[k] < 0 || ijk[k] >= N) return false;
    return refined[morton_index(ijk)] != 0;
  }
};

struct Numbering {
  std::vector<int32_t> V, C, F; // vertex / coarse-object / fine-object lattices, -1 = none or foreign
  uint64_t n = 0;
};

// slot of a dof of a COARSE cell ijk at offsets off[] in [0, p]
inline int32_t *coarse_slot(const Cube &q, Numbering &nb, const int ijk[3], const int off[3]) {
  bool vertex = true;
  for (int k = 0; k < q.dim; ++k) vertex &= (off[k] == 0 || off[k] == q.p);
  uint64_t pos = 0;
  if (vertex) {
    for (int k = 0; k < q.dim; ++k) pos += (uint64_t)(ijk[k] + (off[k] ? 1 : 0)) * q.vstride[k];
    return &nb.V[pos];
  }
  for (int k = 0; k < q.dim; ++k) pos += (uint64_t)(ijk[k] * q.p + off[k]) * q.cstride[k];
  return &nb.C[pos];
}
// slot of a dof of a FINE cell f (fine cell coordinates) at offsets off[]
inline int32_t *fine_slot(const Cube &q, Numbering &nb, const int f[3], const int off[3]) {
  bool coarse_vertex = true;
  int64_t P[3] = {0, 0, 0};
  for (int k = 0; k < q.dim; ++k) {
    P[k] = (int64_t)f[k] * q.p + off[k];
    coarse_vertex &= (P[k] % (2 * q.p) == 0);
  }
  uint64_t pos = 0;
  if (coarse_vertex) {
    for (int k = 0; k < q.dim; ++k) pos += (uint64_t)(P[k] / (2 * q.p)) * q.vstride[k];
    return &nb.V[pos];
  }
  for (int k = 0; k < q.dim; ++k) pos += (uint64_t)(P[k] - (int64_t)2 * q.p * q.rlo[k]) * q.fstride[k];
  return &nb.F[pos];
}

// first-touch numbering of one cube whose low faces in the directions of `foreign_low` belong to
// lower ranks
void number_cube(const Cube &q, const bool foreign_low[3], const std::vector<std::array<int, 3>> &hier,
                 Numbering &nb) {
  nb.V.assign(q.vsize, -1);
  nb.C.assign(q.csize, -1);
  nb.F.assign(q.fsize, -1);
  int32_t next = 0;
  const int npc = (int)hier.size();
  for (uint64_t m = 0; m < q.n_coarse_cells; ++m) {
    if (q.refined[m]) continue;
    int ijk[3];
    q.morton_coords(m, ijk);
    bool low_face[3] = {false, false, false};
    for (int k = 0; k < q.dim; ++k) low_face[k] = foreign_low[k] && ijk[k] == 0;
    for (int h = 0; h < npc; ++h) {
      int off[3] = {hier[h][0], hier[h][1], hier[h][2]};
      bool foreign = false;
      for (int k = 0; k < q.dim; ++k) foreign |= (low_face[k] && off[k] == 0);
      if (foreign) continue;
      int32_t *slot = coarse_slot(q, nb, ijk, off);
      if (*slot < 0) *slot = next++;
    }
  }
  for (uint64_t m = 0; m < q.n_coarse_cells; ++m) {
    if (!q.refined[m]) continue;
    int ijk[3];
    q.morton_coords(m, ijk);
    for (int ch = 0; ch < (1 << q.dim); ++ch) {
      int f[3] = {0, 0, 0};
      for (int k = 0; k < q.dim; ++k) f[k] = 2 * ijk[k] + ((ch >> k) & 1);
      for (int h = 0; h < npc; ++h) {
        int off[3] = {hier[h][0], hier[h][1], hier[h][2]};
        int32_t *slot = fine_slot(q, nb, f, off);
        if (*slot < 0) *slot = next++;
      }
    }
  }
  nb.n = (uint64_t)next;
}

struct ActiveCell {
  uint8_t level; // 0 = coarse, 1 = fine
  int c[3];      // coarse or fine cell coordinates inside the cube
  uint32_t active_index;
};

} // namespace
} // namespace b200mf

using namespace b200mf;

extern "C" {

int b200mf_mesh_create_adaptive(const b200mf_adaptive_desc *ad, b200mf_mesh **out) {
  B200MF_REQUIRE(ad && out, "null argument");
  const b200mf_partition_desc *pd = &ad->part;
  const b200mf_mesh_desc *d = &pd->mesh;
  B200MF_REQUIRE(d->dim == 2 || d->dim == 3, "dim must be 2 or 3");
  B200MF_REQUIRE(d->degree >= 1 && d->degree <= 8, "degree must be in 1..8");
  B200MF_REQUIRE(d->cell_order == B200MF_MESH_MORTON, "adaptive meshes use Morton order");
  B200MF_REQUIRE(d->deformation == B200MF_DEFORM_NONE, "adaptive meshes are Cartesian");
  const int N = d->cells_per_direction, dim = d->dim, p = d->degree, n = p + 1;
  B200MF_REQUIRE(N >= 4 && (N & (N - 1)) == 0, "cells_per_direction must be a power of two >= 4");
  Cube q;
  q.dim = dim; q.p = p; q.n = n; q.N = N; q.radius = ad->ball_radius; q.n_ranks = pd->n_ranks;
  q.levels = 0;
  while ((1 << q.levels) < N) ++q.levels;
  uint64_t n_cubes = 1;
  for (int k = 0; k < 3; ++k) {
    q.coarse[k] = (k < dim && pd->coarse[k] > 0) ? pd->coarse[k] : 1;
    n_cubes *= q.coarse[k];
  }
  B200MF_REQUIRE((uint64_t)pd->n_ranks == n_cubes, "adaptive meshes: n_ranks must equal the number of coarse cubes");
  B200MF_REQUIRE(pd->rank >= 0 && pd->rank < pd->n_ranks, "bad rank");
  B200MF_REQUIRE(pd->n_ranks == 1 || pd->ghost_mode == B200MF_GHOSTS_TOUCHED,
                 "adaptive partitioned meshes support ghost_mode = B200MF_GHOSTS_TOUCHED only");
  q.n_coarse_cells = 1;
  for (int k = 0; k < dim; ++k) q.n_coarse_cells *= N;

  // ---- which cells are refined (the same in every cube)
  q.refined.assign(q.n_coarse_cells, 0);
  for (int k = 0; k < 3; ++k) { q.rlo[k] = N; q.rhi[k] = 0; }
  uint64_t n_refined = 0;
  for (uint64_t m = 0; m < q.n_coarse_cells; ++m) {
    int ijk[3];
    q.morton_coords(m, ijk);
    double r2 = 0.0;
    for (int k = 0; k < dim; ++k) {
      const double x = (ijk[k] + 0.5) / N - 0.5;
      r2 += x * x;
    }
    if (std::sqrt(r2) < q.radius) {
      q.refined[m] = 1;
      ++n_refined;
      for (int k = 0; k < dim; ++k) {
        B200MF_REQUIRE(ijk[k] >= 1 && ijk[k] <= N - 2, "ball_radius too large: refined cells must stay inside the cube");
        q.rlo[k] = std::min(q.rlo[k], ijk[k]);
        q.rhi[k] = std::max(q.rhi[k], ijk[k] + 1);
      }
    }
  }
  if (n_refined == 0)
    for (int k = 0; k < 3; ++k) { q.rlo[k] = 0; q.rhi[k] = 0; }
  q.vsize = q.csize = q.fsize = 1;
  for (int k = 0; k < 3; ++k) {
    q.vext[k] = k < dim ? (uint64_t)N + 1 : 1;
    q.cext[k] = k < dim ? (uint64_t)N * p + 1 : 1;
    q.fext[k] = k < dim ? (uint64_t)2 * p * (q.rhi[k] - q.rlo[k]) + 1 : 1;
    q.vstride[k] = q.vsize; q.vsize *= q.vext[k];
    q.cstride[k] = q.csize; q.csize *= q.cext[k];
    q.fstride[k] = q.fsize; q.fsize *= q.fext[k];
  }
  const auto hier = mesh_hierarchic_offsets(dim, p);
  const int npc = (int)hier.size();

  // ---- this rank's cube, its numbering, the offsets of all ranks
  const int me = pd->rank;
  auto cube_coords = [&](int s, int cc[3]) {
    cc[0] = s % q.coarse[0]; cc[1] = (s / q.coarse[0]) % q.coarse[1]; cc[2] = s / (q.coarse[0] * q.coarse[1]);
  };
  int mycc[3];
  cube_coords(me, mycc);
  // owned counts of every rank depend only on which low faces are foreign: 2^dim variants
  std::map<int, uint64_t> count_of_pattern;
  std::vector<uint64_t> offsets(pd->n_ranks + 1, 0);
  Numbering own;
  {
    for (int s = 0; s < pd->n_ranks; ++s) {
      int cc[3];
      cube_coords(s, cc);
      int pat = 0;
      bool fl[3] = {false, false, false};
      for (int k = 0; k < dim; ++k) { fl[k] = cc[k] > 0; pat |= (fl[k] ? 1 : 0) << k; }
      if (!count_of_pattern.count(pat) || s == me) {
        Numbering tmp;
        number_cube(q, fl, hier, s == me ? own : tmp);
        count_of_pattern[pat] = (s == me ? own : tmp).n;
      }
      offsets[s + 1] = offsets[s] + count_of_pattern[pat];
    }
  }
  B200MF_REQUIRE(own.n < 0x3fffffffull, "too many dofs per rank");

  b200mf_mesh *m = new b200mf_mesh();
  m->desc = *d;
  m->partitioned = true;
  m->n_owned = own.n;
  m->n_dofs = own.n;
  m->rank_offsets = offsets;
  m->first_owned = offsets[me];
  m->n_global_dofs = offsets[pd->n_ranks];
  m->dofs_per_cell = npc;

  // ---- ghosts: the dofs of my coarse cells on low faces shared with lower ranks
  // local slot value for a ghost: -(2 + index into ghost list) while collecting
  struct GhostRef { int lattice; uint64_t pos; int owner; int64_t P[3]; bool vertex; };
  std::vector<GhostRef> ghost_refs;
  bool my_foreign[3] = {false, false, false};
  for (int k = 0; k < dim; ++k) my_foreign[k] = mycc[k] > 0;
  if (my_foreign[0] || my_foreign[1] || my_foreign[2]) {
    auto collect = [&](std::vector<int32_t> &lat, const uint64_t ext[3], bool vertex) {
      uint64_t size = 1;
      for (int k = 0; k < dim; ++k) size *= ext[k];
      for (uint64_t idx = 0; idx < size; ++idx) {
        uint64_t r = idx;
        int64_t P[3] = {0, 0, 0};
        bool on_foreign = false;
        for (int k = 0; k < dim; ++k) {
          P[k] = (int64_t)(r % ext[k]);
          r /= ext[k];
          on_foreign |= (my_foreign[k] && P[k] == 0);
        }
        if (!on_foreign) continue;
        if (!vertex) { // coarse-object lattice: positions that are vertices are not used here
          bool is_vertex = true;
          for (int k = 0; k < dim; ++k) is_vertex &= (P[k] % p == 0);
          if (is_vertex) continue;
        }
        // owner cube: step down across every low interface the point lies on
        int occ[3] = {mycc[0], mycc[1], mycc[2]};
        GhostRef g;
        g.vertex = vertex;
        g.pos = idx;
        g.lattice = vertex ? 0 : 1;
        for (int k = 0; k < dim; ++k) {
          g.P[k] = P[k];
          if (my_foreign[k] && P[k] == 0) { occ[k] -= 1; g.P[k] = (int64_t)ext[k] - 1; }
        }
        g.owner = occ[0] + q.coarse[0] * (occ[1] + q.coarse[1] * occ[2]);
        ghost_refs.push_back(g);
      }
    };
    collect(own.V, q.vext, true);
    collect(own.C, q.cext, false);
  }
  std::vector<std::pair<uint64_t, size_t>> ghosts; // (global, index into ghost_refs)
  {
    std::vector<int> owners;
    for (const auto &g : ghost_refs) owners.push_back(g.owner);
    std::sort(owners.begin(), owners.end());
    owners.erase(std::unique(owners.begin(), owners.end()), owners.end());
    for (int s : owners) {
      int cc[3];
      cube_coords(s, cc);
      bool fl[3] = {false, false, false};
      for (int k = 0; k < dim; ++k) fl[k] = cc[k] > 0;
      Numbering other;
      number_cube(q, fl, hier, other);
      for (size_t i = 0; i < ghost_refs.size(); ++i) {
        const GhostRef &g = ghost_refs[i];
        if (g.owner != s) continue;
        uint64_t pos = 0;
        for (int k = 0; k < dim; ++k) pos += (uint64_t)g.P[k] * (g.vertex ? q.vstride[k] : q.cstride[k]);
        const int32_t v = g.vertex ? other.V[pos] : other.C[pos];
        if (v < 0) { delete m; set_error("internal error: ghost dof not numbered by its owner"); return B200MF_ERR_INVALID; }
        ghosts.emplace_back(offsets[s] + (uint64_t)v, i);
      }
    }
  }
  std::sort(ghosts.begin(), ghosts.end());
  m->n_ghost = ghosts.size();
  m->ghost_global.resize(m->n_ghost);
  for (uint64_t i = 0; i < m->n_ghost; ++i) {
    m->ghost_global[i] = ghosts[i].first;
    const GhostRef &g = ghost_refs[ghosts[i].second];
    (g.vertex ? own.V : own.C)[g.pos] = (int32_t)(m->n_owned + i);
  }
  B200MF_REQUIRE(m->n_owned + m->n_ghost < 0x3fffffffull, "too many local dofs");
  const uint64_t n_local = m->n_owned + m->n_ghost;

  // ---- constrained dofs: Dirichlet boundary (coarse objects and vertices on the domain boundary)
  // and hanging nodes (fine-object dofs inside the closure of an unrefined cell)
  std::vector<uint8_t> is_constrained(n_local, 0), is_hanging(m->n_owned, 0);
  if (d->dirichlet_boundary) {
    auto mark = [&](const std::vector<int32_t> &lat, const uint64_t ext[3]) {
      uint64_t size = 1;
      for (int k = 0; k < dim; ++k) size *= ext[k];
      for (uint64_t idx = 0; idx < size; ++idx) {
        if (lat[idx] < 0) continue;
        uint64_t r = idx;
        bool bnd = false;
        for (int k = 0; k < dim; ++k) {
          const uint64_t P = r % ext[k];
          r /= ext[k];
          bnd |= (P == 0 && mycc[k] == 0) || (P == ext[k] - 1 && mycc[k] == q.coarse[k] - 1);
        }
        if (bnd) is_constrained[lat[idx]] = 1;
      }
    };
    mark(own.V, q.vext);
    mark(own.C, q.cext);
  }
  uint64_t n_hanging = 0;
  for (uint64_t idx = 0; idx < q.fsize; ++idx) {
    if (own.F[idx] < 0) continue;
    uint64_t r = idx;
    int lo[3] = {0, 0, 0}, cnt[3] = {1, 1, 1};
    for (int k = 0; k < dim; ++k) {
      const int64_t P = (int64_t)(r % q.fext[k]) + (int64_t)2 * p * q.rlo[k];
      r /= q.fext[k];
      if (P % (2 * p) == 0) { lo[k] = (int)(P / (2 * p)) - 1; cnt[k] = 2; }
      else { lo[k] = (int)(P / (2 * p)); cnt[k] = 1; }
    }
    bool hanging = false;
    for (int a = 0; a < cnt[0] && !hanging; ++a)
      for (int b = 0; b < cnt[1] && !hanging; ++b)
        for (int c = 0; c < cnt[2] && !hanging; ++c) {
          const int ijk[3] = {lo[0] + a, lo[1] + b, lo[2] + c};
          bool inside = true;
          for (int k = 0; k < dim; ++k) inside &= (ijk[k] >= 0 && ijk[k] < N);
          if (inside && !q.is_refined(ijk)) hanging = true;
        }
    if (hanging) {
      is_constrained[own.F[idx]] = 1;
      is_hanging[own.F[idx]] = 1;
      ++n_hanging;
    }
  }
  for (uint64_t i = 0; i < m->n_owned; ++i)
    if (is_constrained[i]) m->boundary.push_back((uint32_t)i);
  m->n_hanging_dofs = n_hanging;

  // ---- active cells in the reference's order
  std::vector<ActiveCell> active;
  active.reserve(q.n_coarse_cells + n_refined * ((1u << dim) - 1));
  for (uint64_t mm = 0; mm < q.n_coarse_cells; ++mm) {
    if (q.refined[mm]) continue;
    ActiveCell a;
    a.level = 0;
    q.morton_coords(mm, a.c);
    a.active_index = (uint32_t)active.size();
    active.push_back(a);
  }
  const uint64_t n_coarse_active = active.size();
  for (uint64_t mm = 0; mm < q.n_coarse_cells; ++mm) {
    if (!q.refined[mm]) continue;
    int ijk[3];
    q.morton_coords(mm, ijk);
    for (int ch = 0; ch < (1 << dim); ++ch) {
      ActiveCell a;
      a.level = 1;
      a.c[0] = a.c[1] = a.c[2] = 0;
      for (int k = 0; k < dim; ++k) a.c[k] = 2 * ijk[k] + ((ch >> k) & 1);
      a.active_index = (uint32_t)active.size();
      active.push_back(a);
    }
  }
  const uint64_t nc = active.size();
  m->n_cells = nc;
  m->n_global_cells = nc * (uint64_t)pd->n_ranks;

  // ---- masks of the fine cells
  std::vector<uint16_t> mask(nc, 0);
  auto fine_mask = [&](const ActiveCell &a, int side[3], bool face[3], bool edge[3]) -> uint16_t {
    int parent[3] = {0, 0, 0}, cp[3] = {0, 0, 0};
    unsigned mk = 0;
    for (int k = 0; k < dim; ++k) {
      parent[k] = a.c[k] >> 1;
      cp[k] = a.c[k] & 1;
      side[k] = cp[k] == 0 ? -1 : 1;
      if (cp[k] == 0) mk |= 1u << k;
    }
    for (int k = 0; k < 3; ++k) { face[k] = false; edge[k] = false; }
    for (int dd = 0; dd < dim; ++dd) {
      int nb[3] = {parent[0], parent[1], parent[2]};
      nb[dd] += side[dd];
      face[dd] = !q.is_refined(nb); // (inside the cube by construction)
      if (face[dd]) mk |= 8u << dd;
    }
    if (dim == 3)
      for (int e = 0; e < 3; ++e) {
        const int d1 = (e + 1) % 3, d2 = (e + 2) % 3;
        if (!face[d1] && !face[d2]) {
          int nb[3] = {parent[0], parent[1], parent[2]};
          nb[d1] += side[d1];
          nb[d2] += side[d2];
          edge[e] = !q.is_refined(nb);
          if (edge[e]) mk |= 64u << e;
        }
      }
    if ((mk >> 3) == 0) mk = 0;
    return (uint16_t)mk;
  };
#pragma omp parallel for schedule(static)
  for (int64_t i = (int64_t)n_coarse_active; i < (int64_t)nc; ++i) {
    int side[3];
    bool face[3], edge[3];
    mask[i] = fine_mask(active[i], side, face, edge);
  }

  // ---- emission order.  brick_friendly: whole unmasked bricks first (coarse, then fine), the
  // other unmasked cells, the masked cells, and -- behind n_cells_interior -- the cells that
  // touch a ghost dof; else the reference's active cell order.
  std::vector<uint32_t> order;
  order.reserve(nc);
  auto touches_ghost = [&](const ActiveCell &a) {
    if (a.level != 0 || m->n_ghost == 0) return false;
    for (int k = 0; k < dim; ++k)
      if (my_foreign[k] && a.c[k] == 0) return true;
    return false;
  };
  if (!ad->brick_friendly_order) {
    std::vector<uint32_t> bnd;
    for (uint32_t i = 0; i < nc; ++i) (touches_ghost(active[i]) ? bnd : order).push_back(i);
    m->n_cells_interior = order.size();
    order.insert(order.end(), bnd.begin(), bnd.end());
  } else {
    const uint64_t b = dim == 3 ? (uint64_t)brick_edge(p) : 1;
    const uint64_t W = dim == 3 ? b * b * b : 1;
    // active index of a coarse cell by Morton index, of the first child of a refined cell
    std::vector<uint32_t> active_of_coarse(q.n_coarse_cells, 0xffffffffu), first_child(q.n_coarse_cells, 0xffffffffu);
    {
      uint32_t ia = 0, ifine = (uint32_t)n_coarse_active;
      for (uint64_t mm = 0; mm < q.n_coarse_cells; ++mm) {
        if (!q.refined[mm]) active_of_coarse[mm] = ia++;
        else { first_child[mm] = ifine; ifine += 1u << dim; }
      }
    }
    std::vector<uint32_t> int_bricks, fine_bricks, int_rest, masked, bnd_bricks, bnd_rest;
    std::vector<uint8_t> placed(nc, 0);
    // coarse bricks: aligned windows of W Morton-consecutive coarse cells without a refined cell
    for (uint64_t blk = 0; blk + 1 <= q.n_coarse_cells / W; ++blk) {
      bool whole = true, ghost = false;
      for (uint64_t mm = blk * W; mm < (blk + 1) * W; ++mm) {
        if (q.refined[mm]) { whole = false; break; }
        ghost |= touches_ghost(active[active_of_coarse[mm]]);
      }
      if (!whole) continue;
      for (uint64_t mm = blk * W; mm < (blk + 1) * W; ++mm) {
        (ghost ? bnd_bricks : int_bricks).push_back(active_of_coarse[mm]);
        placed[active_of_coarse[mm]] = 1;
      }
    }
    // fine bricks: aligned windows of W / 2^dim refined parents whose children carry no mask
    if (dim == 3 && W >= 8) {
      const uint64_t PW = W / 8;
      for (uint64_t blk = 0; blk + 1 <= q.n_coarse_cells / PW; ++blk) {
        bool whole = true;
        for (uint64_t mm = blk * PW; mm < (blk + 1) * PW && whole; ++mm) {
          if (!q.refined[mm]) { whole = false; break; }
          for (uint32_t ch = 0; ch < 8; ++ch)
            if (mask[first_child[mm] + ch]) { whole = false; break; }
        }
        if (!whole) continue;
        for (uint64_t mm = blk * PW; mm < (blk + 1) * PW; ++mm)
          for (uint32_t ch = 0; ch < 8; ++ch) {
            fine_bricks.push_back(first_child[mm] + ch);
            placed[first_child[mm] + ch] = 1;
          }
      }
    }
    for (uint32_t i = 0; i < nc; ++i) {
      if (placed[i]) continue;
      if (mask[i]) masked.push_back(i);
      else (touches_ghost(active[i]) ? bnd_rest : int_rest).push_back(i);
    }
    for (const auto *lst : {&int_bricks, &fine_bricks, &int_rest, &masked}) order.insert(order.end(), lst->begin(), lst->end());
    m->n_cells_interior = order.size();
    for (const auto *lst : {&bnd_bricks, &bnd_rest}) order.insert(order.end(), lst->begin(), lst->end());
  }

  // ---- per emitted cell: lexicographic index list (with redirection), vertices, mask
  m->l2g.resize(nc * (uint64_t)npc);
  const int nv = 1 << dim;
  m->vertices.resize(nc * (uint64_t)nv * dim);
  m->cell_mask.resize(nc);
  m->active_index.resize(nc);
  const double h = (d->right - d->left) / N;
  const bool markc = d->mark_constrained_l2g && d->dirichlet_boundary;
  bool failed = false;
#pragma omp parallel for schedule(static)
  for (int64_t e = 0; e < (int64_t)nc; ++e) {
    const ActiveCell &a = active[order[e]];
    uint32_t *row = m->l2g.data() + (uint64_t)e * npc;
    m->active_index[e] = a.active_index;
    m->cell_mask[e] = mask[order[e]];
    int side[3] = {0, 0, 0};
    bool face[3] = {false, false, false}, edge[3] = {false, false, false};
    if (a.level == 1) fine_mask(a, side, face, edge);
    for (int j = 0; j < npc; ++j) {
      int off[3] = {j % n, (j / n) % n, dim == 3 ? j / (n * n) : 0};
      int32_t v = -1;
      if (a.level == 0) {
        v = *coarse_slot(q, own, a.c, off);
      } else {
        int parent[3] = {a.c[0] >> 1, a.c[1] >> 1, a.c[2] >> 1};
        bool outer[3] = {false, false, false};
        for (int k = 0; k < dim; ++k) outer[k] = off[k] == ((a.c[k] & 1) == 0 ? 0 : p);
        bool done = false;
        for (int dd = 0; dd < dim && !done; ++dd)
          if (face[dd] && outer[dd]) {
            int nb[3] = {parent[0], parent[1], parent[2]};
            nb[dd] += side[dd];
            int qo[3] = {off[0], off[1], off[2]};
            qo[dd] = p - off[dd];
            v = *coarse_slot(q, own, nb, qo);
            done = true;
          }
        if (!done && dim == 3)
          for (int ed = 0; ed < 3 && !done; ++ed) {
            const int d1 = (ed + 1) % 3, d2 = (ed + 2) % 3;
            if (edge[ed] && outer[d1] && outer[d2]) {
              int nb[3] = {parent[0], parent[1], parent[2]};
              nb[d1] += side[d1];
              nb[d2] += side[d2];
              int qo[3] = {off[0], off[1], off[2]};
              qo[d1] = p - off[d1];
              qo[d2] = p - off[d2];
              v = *coarse_slot(q, own, nb, qo);
              done = true;
            }
          }
        if (!done) v = *fine_slot(q, own, a.c, off);
      }
      if (v < 0) { failed = true; v = 0; }
      uint32_t li = (uint32_t)v;
      if (markc && is_constrained[li] && !(li < m->n_owned && is_hanging[li])) li |= B200MF_L2G_CONSTRAINED;
      row[j] = li;
    }
    const double hh = a.level == 0 ? h : 0.5 * h;
    for (int vtx = 0; vtx < nv; ++vtx)
      for (int k = 0; k < dim; ++k) {
        const double x = d->left + (d->right - d->left) * mycc[k] + hh * (a.c[k] + ((vtx >> k) & 1));
        m->vertices[((uint64_t)e * nv + vtx) * dim + k] = x;
      }
  }
  if (failed) { delete m; set_error("internal error: cell touches an unnumbered dof"); return B200MF_ERR_INVALID; }

  // ---- support point of every local dof (tests, analytic vectors): reference-cell node positions
  // are the Gauss-Lobatto points, built by shape.cpp
  if (pd->want_lattice_ids) {
    std::vector<double> sv, sg, qw, qp, sub, gll;
    build_fe_q_support_points(p, gll);
    m->dof_coords.assign(n_local * 3, 0.0);
    for (uint64_t e = 0; e < nc; ++e) {
      const ActiveCell &a = active[order[e]];
      if (m->cell_mask[e]) continue; // redirected entries do not sit on this cell's own nodes
      const double hh = a.level == 0 ? h : 0.5 * h;
      for (int j = 0; j < npc; ++j) {
        const int off[3] = {j % n, (j / n) % n, dim == 3 ? j / (n * n) : 0};
        const uint32_t li = m->l2g[e * npc + j] & ~B200MF_L2G_CONSTRAINED;
        for (int k = 0; k < dim; ++k)
          m->dof_coords[(uint64_t)li * 3 + k] = d->left + (d->right - d->left) * mycc[k] + hh * (a.c[k] + gll[off[k]]);
      }
    }
    // hanging dofs only appear on masked cells: take them from the cell's own (unredirected) nodes
    for (uint64_t e = 0; e < nc; ++e) {
      if (!m->cell_mask[e]) continue;
      const ActiveCell &a = active[order[e]];
      const double hh = 0.5 * h;
      for (int j = 0; j < npc; ++j) {
        int off[3] = {j % n, (j / n) % n, dim == 3 ? j / (n * n) : 0};
        const int32_t v = *fine_slot(q, own, a.c, off);
        if (v < 0) continue;
        for (int k = 0; k < dim; ++k)
          m->dof_coords[(uint64_t)v * 3 + k] = d->left + (d->right - d->left) * mycc[k] + hh * (a.c[k] + gll[off[k]]);
      }
    }
  }
  *out = m;
  return B200MF_OK;
}

int b200mf_mesh_adaptive_view_get(const b200mf_mesh *m, b200mf_adaptive_view *v) {
  B200MF_REQUIRE(m && v, "null argument");
  B200MF_REQUIRE(!m->cell_mask.empty() || m->n_cells == 0, "mesh was not created by b200mf_mesh_create_adaptive");
  v->constraint_mask = m->cell_mask.data();
  v->active_cell_index = m->active_index.data();
  v->n_hanging_dofs = m->n_hanging_dofs;
  uint64_t nm = 0;
  for (uint16_t x : m->cell_mask) nm += x != 0;
  v->n_masked_cells = nm;
  v->dof_coords = m->dof_coords.empty() ? nullptr : m->dof_coords.data();
  return B200MF_OK;
}

} // extern "C"
