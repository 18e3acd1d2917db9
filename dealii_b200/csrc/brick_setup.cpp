// Brick detection (setup side of brick_kernel.cuh): finds the aligned windows of b^3
// consecutive cells whose local_to_global lists fit together as a b x b x b block of a
// Morton-ordered mesh, and builds one index per lattice node for them.
//
// Plays the role of the index compression of the reference's CPU MatrixFree
// (internal::MatrixFreeFunctions::DoFInfo::IndexStorageVariants, matrix_free/dof_info.h:95-180:
// "interleaved_contiguous" & friends compress the per-cell index lists when the numbering
// allows) for the device path: the only input is the index list the reference's setup
// produces (portable_matrix_free.templates.h:292-298), no mesh topology.
#include <algorithm>
#include <cstdlib>
#include <cstring>

#include "internal.h"

namespace b200mf {

namespace {
inline void morton_decode(unsigned c, int &x, int &y, int &z) {
  x = y = z = 0;
  for (int k = 0; k < 4; ++k) { // brick edges up to 16 cells
    x |= ((c >> (3 * k)) & 1u) << k;
    y |= ((c >> (3 * k + 1)) & 1u) << k;
    z |= ((c >> (3 * k + 2)) & 1u) << k;
  }
}
} // namespace

// Colouring of the bricks -- the alternative to "memset + atomics" the reference offers as graph
// colouring (portable_matrix_free.templates.h:1060-1185, portable_fe_evaluation.h:410-435): bricks
// of one colour share no dof and run in one launch, colours run one after the other, so the
// surface dofs need no atomics.  On top of the reference's scheme the FIRST brick (in launch order)
// that touches a dof STORES it -- flagged in the map -- so vmult needs no memset of dst either; later
// touchers do a plain load-add-store.  Results are bit-reproducible.  Launch order = piece of the
// distributed schedule (first half of the interior | cells touching ghosts | rest of the interior),
// then cell shape, then colour.  Dofs that a cell outside the bricks touches, ghost dofs and dofs
// nobody stores are zeroed before the vmult (zero list + memset of the ghost section).
static int build_colouring(const b200mf_setup_desc &d, Setup &s, std::vector<uint32_t> &maps, uint64_t nb,
                           const std::vector<uint64_t> &brick_cell, const std::vector<uint32_t> &brick_geom) {
  constexpr uint32_t CBIT = B200MF_L2G_CONSTRAINED, COMPLETE = 0x40000000u;
  Setup::Colouring &K = s.colouring;
  K = Setup::Colouring();
  if (nb == 0 || std::getenv("B200MF_NO_COLOURING") != nullptr) return B200MF_OK;
  const int p = s.degree, n = s.n, b = s.brick_b ? s.brick_b : brick_edge(s.degree);
  const int L = b * p + 1;
  const uint64_t L3 = (uint64_t)L * L * L, W = (uint64_t)b * b * b, npc = (uint64_t)n * n * n;
  const uint64_t n_total = s.n_owned + s.n_ghost;
  // same convention as dist_vmult_impl (comm.cu): with ghost dofs, cells [0, n_cells_interior) are the
  // interior (possibly none); without, everything is
  const uint64_t ni = s.n_ghost > 0 ? std::min(s.n_cells_interior, s.n_cells) : s.n_cells;
  const bool pieces = s.n_ghost > 0;
  K.half = pieces ? (ni / 2) / W * W : ni;
  auto piece_of_cell = [&](uint64_t c) -> int { return !pieces ? 0 : (c >= ni ? 1 : (c < K.half ? 0 : 2)); };
  // dofs the per-cell kernels accumulate into (atomics on zeroed entries): never stored by a brick
  std::vector<uint8_t> early(n_total, 0);
  for (uint64_t i = s.n_owned; i < n_total; ++i) early[i] = 1;
  {
    std::vector<uint8_t> in_brick(s.n_cells, 0);
    for (uint64_t w = 0; w < nb; ++w)
      for (uint64_t c = 0; c < W; ++c) in_brick[brick_cell[w] + c] = 1;
    for (uint64_t c = 0; c < s.n_cells; ++c) {
      if (in_brick[c]) continue;
      for (uint64_t e = c * npc; e < (c + 1) * npc; ++e) {
        const uint32_t v = d.local_to_global[e];
        if (!(v & CBIT) && v < n_total) early[v] = 1;
      }
    }
  }
  // greedy colouring in brick order: a colour is free if no brick sharing a dof has it
  std::vector<uint32_t> used(n_total, 0); // per dof: colours of the bricks touching it
  std::vector<uint8_t> colour(nb, 0);
  int n_colours = 0;
  for (uint64_t w = 0; w < nb; ++w) {
    const uint32_t *m = maps.data() + w * L3;
    uint32_t forbidden = 0;
    for (uint64_t e = 0; e < L3; ++e)
      if (!(m[e] & (CBIT | COMPLETE))) forbidden |= used[m[e] & B200MF_BRICK_INDEX];
    int c = 0;
    while (c < 32 && (forbidden >> c) & 1u) ++c;
    if (c >= 32) return B200MF_OK; // no colouring: the atomics path serves the setup
    colour[w] = (uint8_t)c;
    n_colours = std::max(n_colours, c + 1);
    for (uint64_t e = 0; e < L3; ++e)
      if (!(m[e] & (CBIT | COMPLETE))) used[m[e] & B200MF_BRICK_INDEX] |= 1u << c;
  }
  used.clear();
  used.shrink_to_fit();
  // launch key of a brick and the smallest key among the touchers of every dof
  auto key_of = [&](uint64_t w) -> uint32_t {
    return ((uint32_t)piece_of_cell(brick_cell[w]) << 16) | (brick_geom[w] << 8) | colour[w];
  };
  std::vector<uint32_t> min_key(n_total, 0xffffffffu);
  for (uint64_t w = 0; w < nb; ++w) {
    const uint32_t *m = maps.data() + w * L3, k = key_of(w);
    for (uint64_t e = 0; e < L3; ++e)
      if (!(m[e] & (CBIT | COMPLETE))) {
        uint32_t &mk = min_key[m[e] & B200MF_BRICK_INDEX];
        mk = std::min(mk, k);
      }
  }
  std::vector<uint8_t> stored(n_total, 0);
#pragma omp parallel for schedule(static)
  for (int64_t w = 0; w < (int64_t)nb; ++w) {
    uint32_t *m = maps.data() + (uint64_t)w * L3;
    const uint32_t k = key_of(w);
    for (uint64_t e = 0; e < L3; ++e) {
      if (m[e] & CBIT) continue;
      const uint32_t v = m[e] & B200MF_BRICK_INDEX;
      if (m[e] & COMPLETE) { stored[v] = 1; continue; }
      if (!early[v] && min_key[v] == k) {
        m[e] |= B200MF_BRICK_FIRST;
        stored[v] = 1;
      }
    }
  }
  // launches: bricks sorted by key, one launch per key
  std::vector<uint32_t> order(nb);
  for (uint64_t w = 0; w < nb; ++w) order[w] = (uint32_t)w;
  std::stable_sort(order.begin(), order.end(), [&](uint32_t a, uint32_t c) { return key_of(a) < key_of(c); });
  for (uint64_t i = 0; i < nb;) {
    uint64_t j = i;
    const uint32_t k = key_of(order[i]);
    while (j < nb && key_of(order[j]) == k) ++j;
    K.launches.push_back({(int)(k >> 16), (k >> 8) & 0xffu, i, j - i});
    i = j;
  }
  std::vector<uint32_t> zero_list;
  for (uint64_t i = 0; i < s.n_owned; ++i)
    if (!stored[i]) zero_list.push_back((uint32_t)i);
  K.n_zero = zero_list.size();
  K.n_colours = n_colours;
  // cells outside the bricks, per piece
  {
    uint64_t pos = 0;
    auto add = [&](uint64_t cb, uint64_t ce) {
      // split at the piece boundaries
      const uint64_t cuts[4] = {0, K.half, ni, s.n_cells};
      for (int q = 0; q < 3; ++q) {
        const uint64_t a = std::max(cb, cuts[q]), e = std::min(ce, cuts[q + 1]);
        if (e > a) K.general.push_back({piece_of_cell(a), a, e});
      }
    };
    for (const Setup::BrickRun &run : s.brick_runs) {
      add(pos, run.cell_begin);
      pos = run.cell_end;
    }
    add(pos, s.n_cells);
  }
  B200MF_CUDA_CHECK(cudaMalloc((void **)&K.d_list, nb * sizeof(uint32_t)));
  B200MF_CUDA_CHECK(cudaMemcpy(K.d_list, order.data(), nb * sizeof(uint32_t), cudaMemcpyHostToDevice));
  B200MF_CUDA_CHECK(cudaMalloc((void **)&K.d_zero, std::max<size_t>(zero_list.size(), 1) * sizeof(uint32_t)));
  B200MF_CUDA_CHECK(cudaMemcpy(K.d_zero, zero_list.data(), zero_list.size() * sizeof(uint32_t), cudaMemcpyHostToDevice));
  s.device_bytes += (nb + zero_list.size()) * sizeof(uint32_t);
  s.index_bytes += (nb + zero_list.size()) * sizeof(uint32_t);
  K.ready = true;
  return B200MF_OK;
}

// Fills s.brick_* ; returns B200MF_OK also when no brick was found.
int build_bricks(const b200mf_setup_desc &d, Setup &s, bool upload, uint64_t *n_complete_out,
                 BulkStats *bulk_stats) {
  s.brick_b = brick_edge(s.degree);
  constexpr uint32_t CBIT = B200MF_L2G_CONSTRAINED, COMPLETE = 0x40000000u, UNSET = 0xffffffffu;
  s.n_bricks = 0;
  s.brick_runs.clear();
  // a window qualifies when its cells carry no hanging-node mask and share one cell shape
  if (s.dim != 3 || s.cell_kind != B200MF_CELLS_CARTESIAN || s.n_geom < 1 || s.n_geom > 64) return B200MF_OK;
  const uint16_t *cmask = s.any_mask ? d.constraint_mask : nullptr;
  const uint32_t *gid = (s.n_geom > 1 && s.h_geom_id.size() == s.n_cells) ? s.h_geom_id.data() : nullptr;
  if (s.n_geom > 1 && gid == nullptr) return B200MF_OK;
  const uint64_t n_total = s.n_owned + s.n_ghost;
  if (n_total >= B200MF_BRICK_FIRST) return B200MF_OK; // bits 29..31 of a map entry are flags
  const int p = s.degree, n = s.n, b = brick_edge(s.degree);
  const int L = b * p + 1, L2 = L * L;
  const uint64_t L3 = (uint64_t)L2 * L, W = (uint64_t)b * b * b, npc = (uint64_t)n * n * n;
  // aligned windows of W consecutive cells: aligned to cell 0 inside the interior class and to
  // n_cells_interior behind it (the generators emit whole blocks first in both classes)
  const uint64_t ni_split = (s.n_cells_interior > 0 && s.n_cells_interior < s.n_cells) ? s.n_cells_interior : s.n_cells;
  std::vector<uint64_t> win_cell; // first cell of every window
  for (uint64_t c = 0; c + W <= ni_split; c += W) win_cell.push_back(c);
  for (uint64_t c = ni_split; c + W <= s.n_cells; c += W) win_cell.push_back(c);
  const uint64_t n_windows = win_cell.size();
  if (n_windows == 0) return B200MF_OK;
  const uint32_t *l2g = d.local_to_global;

  // how many (cell, local dof) pairs reference every dof
  std::vector<uint8_t> count(n_total, 0);
  const int64_t n_entries = (int64_t)(s.n_cells * npc);
#pragma omp parallel for schedule(static)
  for (int64_t e = 0; e < n_entries; ++e) {
    const uint32_t v = l2g[e];
    if (!(v & CBIT) && v < n_total) {
      // saturating at 255 (a dof referenced by > 254 local cells is simply never "complete")
      uint8_t old = __atomic_load_n(&count[v], __ATOMIC_RELAXED);
      while (old != 255 &&
             !__atomic_compare_exchange_n(&count[v], &old, (uint8_t)(old + 1), true, __ATOMIC_RELAXED,
                                          __ATOMIC_RELAXED)) {
      }
    }
  }

  std::vector<uint32_t> maps;
  try {
    maps.resize(n_windows * L3);
  } catch (const std::bad_alloc &) {
    return B200MF_OK; // no bricks: the per-cell kernels serve the whole mesh
  }
  std::vector<uint8_t> ok(n_windows, 0);
  // position of the cells of a window in the block, and multiplicity of a lattice coordinate
  std::vector<int> cx(W), cy(W), cz(W), mult(L);
  for (unsigned c = 0; c < W; ++c) morton_decode(c, cx[c], cy[c], cz[c]);
  for (int X = 0; X < L; ++X) mult[X] = (X % p == 0 && X > 0 && X < L - 1) ? 2 : 1;

#pragma omp parallel for schedule(dynamic, 16)
  for (int64_t w = 0; w < (int64_t)n_windows; ++w) {
    uint32_t *lat = maps.data() + (uint64_t)w * L3;
    std::fill(lat, lat + L3, UNSET);
    bool good = true;
    const uint64_t wc = win_cell[w];
    if (cmask)
      for (unsigned c = 0; c < W && good; ++c) good = cmask[wc + c] == 0;
    if (gid)
      for (unsigned c = 1; c < W && good; ++c) good = gid[wc + c] == gid[wc];
    for (unsigned c = 0; c < W && good; ++c) {
      const uint32_t *cl = l2g + (wc + c) * npc;
      const int ox = cx[c] * p, oy = cy[c] * p, oz = cz[c] * p;
      for (int k = 0; k < n && good; ++k)
        for (int j = 0; j < n && good; ++j)
          for (int i = 0; i < n; ++i) {
            uint32_t v = cl[i + n * (j + n * k)];
            if (v & CBIT) v = CBIT;
            else if (v >= n_total) { good = false; break; }
            uint32_t &slot = lat[(ox + i) + L * ((oy + j) + L * (oz + k))];
            if (slot == UNSET) slot = v;
            else if (slot != v) { good = false; break; }
          }
    }
    if (!good) continue;
    for (int Z = 0; Z < L; ++Z)
      for (int Y = 0; Y < L; ++Y)
        for (int X = 0; X < L; ++X) {
          uint32_t &v = lat[X + L * (Y + L * Z)];
          if (v == CBIT) continue;
          if (count[v] == mult[X] * mult[Y] * mult[Z]) v |= COMPLETE;
        }
    ok[w] = 1;
  }

  // compact the maps of the accepted windows and record the runs of consecutive bricks
  uint64_t nb = 0;
  std::vector<uint64_t> brick_cell;
  std::vector<uint32_t> brick_geom;
  for (uint64_t w = 0; w < n_windows; ++w) {
    if (!ok[w]) continue;
    if (nb != w) std::memmove(maps.data() + nb * L3, maps.data() + w * L3, L3 * sizeof(uint32_t));
    const uint64_t wc = win_cell[w];
    const uint32_t g = gid ? gid[wc] : 0u;
    if (!s.brick_runs.empty() && s.brick_runs.back().cell_end == wc && s.brick_runs.back().geom == g)
      s.brick_runs.back().cell_end = wc + W;
    else
      s.brick_runs.push_back({wc, wc + W, nb, g});
    brick_cell.push_back(wc);
    brick_geom.push_back(g);
    ++nb;
  }
  if (n_complete_out) {
    uint64_t nc = 0;
    for (uint64_t e = 0; e < nb * L3; ++e)
      if (!(maps[e] & CBIT) && (maps[e] & COMPLETE)) ++nc;
    *n_complete_out = nc;
  }
  if (!upload) { // host-only probe (b200mf_brick_probe): counts, no device arrays
    s.n_bricks = nb;
    s.brick_b = b;
    if (bulk_stats) {
      maps.resize(nb * L3);
      return build_bulk(d, s, maps, nb, false, bulk_stats);
    }
    return B200MF_OK;
  }
  if (nb == 0) return B200MF_OK;
  // vmult stores the complete dofs, so only the others have to be zeroed first.  Measured on
  // B200 (Q4 FP64, 135 M dofs): zeroing that list (18 % of the dofs, in runs of 3-9) takes as long
  // as the memset of the whole vector (1.121 vs 1.110 ms per vmult) -- partial-sector writes --
  // so the list is only built on request (B200MF_SELECTIVE_ZERO=1), for A/B runs.
  if (std::getenv("B200MF_SELECTIVE_ZERO") != nullptr) {
    std::vector<uint8_t> complete(n_total, 0);
#pragma omp parallel for schedule(static)
    for (int64_t e = 0; e < (int64_t)(nb * L3); ++e) {
      const uint32_t v = maps[e];
      if (!(v & CBIT) && (v & COMPLETE)) complete[v & B200MF_BRICK_INDEX] = 1;
    }
    std::vector<uint32_t> zero_list;
    for (uint64_t i = 0; i < n_total; ++i)
      if (!complete[i]) zero_list.push_back((uint32_t)i);
    if (zero_list.size() * 3 < n_total) { // worth it: scattered sector writes cost ~2.5x a stream
      B200MF_CUDA_CHECK(cudaMalloc((void **)&s.d_zero_list, std::max<size_t>(zero_list.size(), 1) * sizeof(uint32_t)));
      B200MF_CUDA_CHECK(cudaMemcpy(s.d_zero_list, zero_list.data(), zero_list.size() * sizeof(uint32_t),
                                   cudaMemcpyHostToDevice));
      s.n_zero_list = zero_list.size();
      s.have_zero_list = true;
      s.device_bytes += zero_list.size() * sizeof(uint32_t);
      s.index_bytes += zero_list.size() * sizeof(uint32_t);
    }
  }
  maps.resize(nb * L3);
  // ---- strided bricks: index = base + x + sy y + sz z for every node, flags constant per lattice
  // face (what DoFRenumbering::lexicographic, include/deal.II/dofs/dof_renumbering.h:1327-1342, produces
  // on a tensor-product mesh): then no map has to be streamed at all
  {
    std::vector<uint4> desc(nb);
    bool all = std::getenv("B200MF_NO_STRIDED") == nullptr;
#pragma omp parallel for schedule(static)
    for (int64_t w = 0; w < (int64_t)nb; ++w) {
      if (!all) continue;
      const uint32_t *m = maps.data() + (uint64_t)w * L3;
      // reference node: the lattice centre is never constrained by a face flag... any unconstrained node
      bool ok = true;
      // face flags from the face centres' neighbours: derive per face from all its nodes
      uint32_t shared = 0, cons = 0;
      auto face_of = [&](int X, int Y, int Z, int f) {
        switch (f) { case 0: return X == 0; case 1: return X == L - 1; case 2: return Y == 0; case 3: return Y == L - 1;
                     case 4: return Z == 0; default: return Z == L - 1; }
      };
      // a face is constrained / shared if its centre node is
      for (int f = 0; f < 6; ++f) {
        int X = L / 2, Y = L / 2, Z = L / 2;
        if (f == 0) X = 0; if (f == 1) X = L - 1; if (f == 2) Y = 0; if (f == 3) Y = L - 1; if (f == 4) Z = 0; if (f == 5) Z = L - 1;
        const uint32_t v = m[X + L * (Y + L * Z)];
        if (v & CBIT) cons |= 1u << f;
        else if (!(v & COMPLETE)) shared |= 1u << f;
      }
      // strides from an interior node and its neighbours (interior nodes are never constrained here)
      const int c0 = 1 + L * (1 + L * 1);
      if ((m[c0] | m[c0 + 1] | m[c0 + L] | m[c0 + L2]) & CBIT) ok = false;
      uint32_t sx = 0, sy = 0, sz = 0, base = 0;
      if (ok) {
        const uint32_t i0 = m[c0] & B200MF_BRICK_INDEX;
        sx = (m[c0 + 1] & B200MF_BRICK_INDEX) - i0;
        sy = (m[c0 + L] & B200MF_BRICK_INDEX) - i0;
        sz = (m[c0 + L2] & B200MF_BRICK_INDEX) - i0;
        base = i0 - sx - sy - sz;
        if (sx != 1 || (int32_t)sy <= 0 || (int32_t)sz <= 0) ok = false;
      }
      for (int Z = 0; Z < L && ok; ++Z)
        for (int Y = 0; Y < L && ok; ++Y)
          for (int X = 0; X < L; ++X) {
            const uint32_t v = m[X + L * (Y + L * Z)];
            bool want_cons = false, want_shared = false;
            for (int f = 0; f < 6; ++f)
              if (face_of(X, Y, Z, f)) { want_cons |= (cons >> f) & 1u; want_shared |= (shared >> f) & 1u; }
            if (want_cons) { if (!(v & CBIT)) { ok = false; break; } continue; }
            if (v & CBIT) { ok = false; break; }
            if ((v & B200MF_BRICK_INDEX) != base + X + sy * Y + sz * Z) { ok = false; break; }
            if (((v & COMPLETE) != 0) == want_shared) { ok = false; break; }
          }
      if (!ok) {
#pragma omp atomic write
        all = false;
      }
      desc[w] = make_uint4(base, sy, sz, shared | (cons << 6));
    }
    if (all && nb > 0) {
      B200MF_CUDA_CHECK(cudaMalloc((void **)&s.d_brick_strided, nb * sizeof(uint4)));
      B200MF_CUDA_CHECK(cudaMemcpy(s.d_brick_strided, desc.data(), nb * sizeof(uint4), cudaMemcpyHostToDevice));
      s.device_bytes += nb * sizeof(uint4);
      s.index_bytes += nb * sizeof(uint4);
    }
  }
  int rcc = s.d_brick_strided ? B200MF_OK : build_colouring(d, s, maps, nb, brick_cell, brick_geom);
  if (rcc != B200MF_OK) return rcc;
  B200MF_CUDA_CHECK(cudaMalloc((void **)&s.d_brick_map, nb * L3 * sizeof(uint32_t)));
  B200MF_CUDA_CHECK(cudaMemcpy(s.d_brick_map, maps.data(), nb * L3 * sizeof(uint32_t),
                               cudaMemcpyHostToDevice));
  s.n_bricks = nb;
  s.brick_b = b;
  s.device_bytes += nb * L3 * sizeof(uint32_t);
  s.index_bytes += nb * L3 * sizeof(uint32_t);
  if (s.n_geom != 1) return B200MF_OK; // the bulk kernel applies one cell shape per launch
  return build_bulk(d, s, maps, nb, true, bulk_stats);
}

} // namespace b200mf
