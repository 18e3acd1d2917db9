// Setup of the bulk brick path (bulk_kernel.cuh): turns the per-brick lattice index maps of
// brick_setup.cpp into
//   * a dictionary of relative index PATTERNS shared by many bricks (on a uniformly refined mesh
//     with deal.II's first-touch numbering, dof_handler_policy.cc:1676-1719, a handful of
//     patterns covers every brick), and a 192-byte descriptor per brick (pattern id + group bases);
//   * the "own range" of every brick: the contiguous index range whose dofs the brick touches
//     first in execution order -- read with ONE bulk copy, written with ONE bulk copy;
//   * the write protocol that replaces "memset + atomics everywhere"
//     (portable_matrix_free.templates.h:1060-1185 offers colouring or atomics): the first
//     toucher of a dof STORES it, later touchers add with RED after the first toucher's flag;
//     dofs nobody stores (ghost section, dofs of cells outside bricks) are zeroed beforehand.
// Everything is derived from the index maps alone (no mesh topology, no assumption on the
// numbering): a brick whose numbering is irregular simply gets a private pattern.
#include <algorithm>
#include <climits>
#include <cstdlib>
#include <cstring>
#include <unordered_map>

#include "internal.h"

namespace b200mf {

namespace {
constexpr uint32_t CBIT = B200MF_L2G_CONSTRAINED, IDX = B200MF_BRICK_INDEX;
constexpr int32_t FT_NONE = INT_MAX, FT_EARLY = -1;

inline uint64_t hash_words(const uint32_t *w, size_t n, uint64_t h = 0x9e3779b97f4a7c15ull) {
  for (size_t i = 0; i < n; ++i) {
    h ^= w[i] + 0x9e3779b97f4a7c15ull + (h << 6) + (h >> 2);
    h *= 0xff51afd7ed558ccdull;
    h ^= h >> 33;
  }
  return h;
}
} // namespace

void free_bulk(Setup &s) {
  Setup::Bulk &B = s.bulk;
  cudaFree(B.d_desc); cudaFree(B.d_other); cudaFree(B.d_phdr); cudaFree(B.d_own_pos); cudaFree(B.d_other_pos);
  cudaFree(B.d_flags); cudaFree(B.d_ticket); cudaFree(B.d_zero);
  B = Setup::Bulk();
}

// maps: [nb][L^3] lattice index maps of the accepted bricks (brick_setup.cpp), in brick order.
// Returns B200MF_OK also when the bulk path is not usable (s.bulk.ready stays false).
int build_bulk(const b200mf_setup_desc &d, Setup &s, const std::vector<uint32_t> &maps, uint64_t nb,
               bool upload, BulkStats *stats) {
  Setup::Bulk &B = s.bulk;
  B.ready = false;
  if (nb == 0 || std::getenv("B200MF_NO_BULK") != nullptr) return B200MF_OK;
  const int p = s.degree, n = s.n, b = s.brick_b;
  const int L = b * p + 1, L2 = L * L, TP = ((L2 + 31) / 32) * 32;
  const uint64_t L3 = (uint64_t)L2 * L, W = (uint64_t)b * b * b, npc = (uint64_t)n * n * n;
  const uint64_t n_total = s.n_owned + s.n_ghost;
  const uint32_t A = s.number == B200MF_F64 ? 2u : 4u; // elements per 16 bytes (bulk copy granule)
  if (nb >= (1ull << 31)) return B200MF_OK;

  // ---- execution order of the bricks: [first half of the interior bricks | bricks that touch a
  // ghost dof | second half of the interior bricks] -- the schedule of distributed_cell_loop
  // (portable_matrix_free.templates.h:1602-1656) inside one launch; plain brick order without ghosts
  std::vector<uint32_t> exec_of(nb), brick_at(nb);
  std::vector<uint8_t> is_boundary(nb, 0);
  uint64_t n_boundary = 0;
  if (s.n_ghost) {
#pragma omp parallel for schedule(static) reduction(+ : n_boundary)
    for (int64_t w = 0; w < (int64_t)nb; ++w) {
      const uint32_t *m = maps.data() + (uint64_t)w * L3;
      for (uint64_t e = 0; e < L3; ++e)
        if (!(m[e] & CBIT) && (m[e] & IDX) >= s.n_owned) { is_boundary[w] = 1; ++n_boundary; break; }
    }
  }
  {
    const uint64_t n_int = nb - n_boundary, half = s.n_ghost ? (n_int + 1) / 2 : n_int;
    uint64_t pos_a = 0, pos_b = half, pos_c = half + n_boundary, seen_int = 0;
    for (uint64_t w = 0; w < nb; ++w) {
      uint64_t t;
      if (is_boundary[w]) t = pos_b++;
      else t = (seen_int++ < half) ? pos_a++ : pos_c++;
      exec_of[w] = (uint32_t)t;
      brick_at[t] = (uint32_t)w;
    }
    B.exec_boundary_begin = half;
    B.exec_boundary_end = half + n_boundary;
  }

  // ---- first toucher of every dof in execution order
  std::vector<int32_t> ft(n_total, FT_NONE);
  for (uint64_t i = s.n_owned; i < n_total; ++i) ft[i] = FT_EARLY; // ghost section: zeroed, RED by all
  // cells outside the bricks run in an earlier launch (atomics into zeroed entries)
  B.general_ranges.clear();
  {
    uint64_t pos = 0;
    auto mark = [&](uint64_t cb, uint64_t ce) {
      if (ce <= cb) return;
      B.general_ranges.push_back({cb, ce});
      for (uint64_t e = cb * npc; e < ce * npc; ++e) {
        const uint32_t v = d.local_to_global[e];
        if (!(v & CBIT) && v < n_total) ft[v] = FT_EARLY;
      }
    };
    for (const Setup::BrickRun &run : s.brick_runs) {
      mark(pos, run.cell_begin);
      pos = run.cell_end;
    }
    mark(pos, s.n_cells);
  }
#pragma omp parallel for schedule(static)
  for (int64_t w = 0; w < (int64_t)nb; ++w) {
    const uint32_t *m = maps.data() + (uint64_t)w * L3;
    const int32_t t = (int32_t)exec_of[w];
    for (uint64_t e = 0; e < L3; ++e) {
      if (m[e] & CBIT) continue;
      int32_t *slot = &ft[m[e] & IDX];
      int32_t old = __atomic_load_n(slot, __ATOMIC_RELAXED);
      while (t < old && !__atomic_compare_exchange_n(slot, &old, t, true, __ATOMIC_RELAXED, __ATOMIC_RELAXED)) {
      }
    }
  }

  // ---- per brick: own range, relative table, dependencies; patterns deduplicated by content
  constexpr int kMaxDeps = 26, kMaxHoles = 1 << 20, kDescWords = 48;
  struct Pattern { std::vector<uint32_t> entries; std::vector<uint32_t> holes; };
  std::vector<Pattern> patterns;
  std::unordered_multimap<uint64_t, uint32_t> dict;
  std::vector<uint32_t> desc(nb * (uint64_t)kDescWords, 0);
  std::vector<uint8_t> stored(n_total, 0); // dofs some brick stores (own range incl. holes, FIRST nodes)
  bool failed = false;
  uint64_t n_own_total = 0, n_first_scalar = 0, n_later = 0;
  // the bricks are processed in parallel into private tables, then merged serially
  struct Work { std::vector<uint32_t> entries, holes; uint32_t lo = 0, R = 0, base[16] = {0}; std::vector<uint32_t> deps; bool ok = true; };
  const int64_t chunk = 2048;
  for (int64_t w0 = 0; w0 < (int64_t)nb && !failed; w0 += chunk) {
    const int64_t w1 = std::min<int64_t>(nb, w0 + chunk);
    std::vector<Work> work(w1 - w0);
#pragma omp parallel for schedule(dynamic, 16)
    for (int64_t w = w0; w < w1; ++w) {
      Work &k = work[w - w0];
      const uint32_t *m = maps.data() + (uint64_t)w * L3;
      const int32_t t = (int32_t)exec_of[w];
      // own range: contiguous, every index first touched by this brick or by nobody, aligned to
      // the 16-byte granule of the bulk copies; what falls outside stays a FIRST (scalar) node
      uint32_t lo = UINT32_MAX, hi = 0;
      for (uint64_t e = 0; e < L3; ++e) {
        if (m[e] & CBIT) continue;
        const uint32_t v = m[e] & IDX;
        if (ft[v] == t) { lo = std::min(lo, v); hi = std::max(hi, v); }
      }
      uint32_t R = 0;
      if (lo != UINT32_MAX) {
        lo = (lo + A - 1) / A * A;
        uint32_t end = (hi + 1) / A * A;
        if (end > lo && end - lo <= L3 / A * A && end <= s.n_owned) {
          bool clean = true;
          for (uint32_t i = lo; i < end; ++i)
            if (ft[i] != t && ft[i] != FT_NONE) { clean = false; break; }
          if (clean) R = end - lo;
        }
      }
      if (R == 0) lo = 0;
      k.lo = lo;
      k.R = R;
      // groups of the nodes outside the own range: by position class (each coordinate low / mid /
      // high), a base per group so that the offsets are small and shared between bricks
      int slot_of_class[27];
      uint32_t base_of_class[27];
      for (int c = 0; c < 27; ++c) { slot_of_class[c] = -1; base_of_class[c] = UINT32_MAX; }
      auto cls = [&](int X, int Y, int Z) {
        auto c1 = [&](int q) { return q == 0 ? 0 : (q == L - 1 ? 2 : 1); };
        return c1(X) + 3 * c1(Y) + 9 * c1(Z);
      };
      for (int Z = 0, e = 0; Z < L; ++Z)
        for (int Y = 0; Y < L; ++Y)
          for (int X = 0; X < L; ++X, ++e) {
            if (m[e] & CBIT) continue;
            const uint32_t v = m[e] & IDX;
            if (R && v >= lo && v < lo + R && ft[v] == t) continue;
            const int c = cls(X, Y, Z);
            base_of_class[c] = std::min(base_of_class[c], v);
          }
      int n_used = 0; // slots 1..14; more groups than that share slot 14
      for (int i = 0; i < 16; ++i) k.base[i] = UINT32_MAX;
      for (int c = 0; c < 27; ++c)
        if (base_of_class[c] != UINT32_MAX) {
          const int sl = n_used < 14 ? ++n_used : 14;
          slot_of_class[c] = sl;
          k.base[sl] = std::min(k.base[sl], base_of_class[c]);
        }
      for (int i = 0; i < 16; ++i)
        if (k.base[i] == UINT32_MAX) k.base[i] = 0;
      k.entries.resize(L3);
      for (int Z = 0, e = 0; Z < L; ++Z)
        for (int Y = 0; Y < L; ++Y)
          for (int X = 0; X < L; ++X, ++e) {
            if (m[e] & CBIT) { k.entries[e] = 15u << 28; continue; }
            const uint32_t v = m[e] & IDX;
            if (R && v >= lo && v < lo + R && ft[v] == t) { k.entries[e] = v - lo; continue; }
            const int sl = slot_of_class[cls(X, Y, Z)];
            const uint32_t off = v - k.base[sl];
            if (off >= (1u << 27)) { k.ok = false; continue; }
            const bool first = ft[v] == t;
            k.entries[e] = ((uint32_t)sl << 28) | (first ? (1u << 27) : 0u) | off;
            if (!first && ft[v] >= 0) {
              const uint32_t dep = (uint32_t)ft[v];
              if (std::find(k.deps.begin(), k.deps.end(), dep) == k.deps.end()) k.deps.push_back(dep);
            }
          }
      for (uint32_t i = 0; i < R; ++i)
        if (ft[lo + i] == FT_NONE) k.holes.push_back(i);
      if ((int)k.deps.size() > kMaxDeps || (int)k.holes.size() > kMaxHoles) k.ok = false;
      std::sort(k.deps.begin(), k.deps.end());
    }
    for (int64_t w = w0; w < w1; ++w) {
      Work &k = work[w - w0];
      if (!k.ok) { failed = true; break; }
      // dictionary lookup
      uint64_t h = hash_words(k.entries.data(), k.entries.size());
      h = hash_words(k.holes.data(), k.holes.size(), h ^ k.holes.size());
      uint32_t pid = UINT32_MAX;
      auto range = dict.equal_range(h);
      for (auto it = range.first; it != range.second; ++it)
        if (patterns[it->second].entries == k.entries && patterns[it->second].holes == k.holes) { pid = it->second; break; }
      if (pid == UINT32_MAX) {
        pid = (uint32_t)patterns.size();
        patterns.push_back({std::move(k.entries), k.holes});
        dict.emplace(h, pid);
        // private patterns cost 2 L^3 words each: give up when the numbering has no regularity
        if (patterns.size() > 4096 && patterns.size() * 4 > nb) { failed = true; break; }
      }
      uint32_t *D = desc.data() + (uint64_t)exec_of[w] * kDescWords;
      D[0] = pid; D[1] = k.lo; D[2] = k.R; D[3] = (uint32_t)k.deps.size();
      for (int i = 1; i < 16; ++i) D[4 + i] = k.base[i];
      D[4] = k.lo;
      for (size_t i = 0; i < k.deps.size(); ++i) D[20 + i] = k.deps[i];
      // bookkeeping of what gets stored
      const Pattern &P = patterns[pid];
      for (uint32_t i = 0; i < k.R; ++i) stored[k.lo + i] = 1;
      n_own_total += k.R - P.holes.size();
      uint32_t later_here = 0;
      for (uint64_t e = 0; e < L3; ++e) {
        const uint32_t en = P.entries[e], sl = en >> 28;
        if (sl == 0 || sl == 15) continue;
        if (en & (1u << 27)) { stored[k.base[sl] + (en & 0x7ffffffu)] = 1; ++n_first_scalar; }
        else ++later_here;
      }
      D[46] = later_here;
      n_later += later_here;
    }
  }
  if (failed) return B200MF_OK;

  // self-check: the tables reproduce the maps
  for (uint64_t w = 0; w < nb; w += std::max<uint64_t>(1, nb / 64)) {
    const uint32_t *m = maps.data() + w * L3;
    const uint32_t *D = desc.data() + (uint64_t)exec_of[w] * kDescWords;
    const Pattern &P = patterns[D[0]];
    for (uint64_t e = 0; e < L3; ++e) {
      const uint32_t en = P.entries[e], sl = en >> 28;
      const uint32_t want = (m[e] & CBIT) ? UINT32_MAX : (m[e] & IDX);
      const uint32_t got = sl == 15 ? UINT32_MAX : D[4 + sl] + (en & 0x7ffffffu);
      if (want != got) {
        set_error("bulk brick tables do not reproduce the index map (brick %llu node %llu)",
                  (unsigned long long)w, (unsigned long long)e);
        return B200MF_ERR_INVALID;
      }
    }
  }

  // ---- dofs nobody stores: zero them before the launch (ghost section by memset)
  std::vector<uint32_t> zero_list;
  for (uint64_t i = 0; i < s.n_owned; ++i)
    if (!stored[i]) zero_list.push_back((uint32_t)i);

  BulkStats local_stats;
  if (!stats) stats = &local_stats;
  {
    stats->n_bricks = nb;
    stats->n_patterns = patterns.size();
    stats->n_own = n_own_total;
    stats->n_first_scalar = n_first_scalar;
    stats->n_later = n_later;
    stats->n_zero = zero_list.size();
    stats->n_general_cells = 0;
    for (auto &r : B.general_ranges) stats->n_general_cells += r.second - r.first;
    stats->n_boundary_bricks = n_boundary;
  }
  B.stats = *stats;
  B.n_exec = (uint32_t)nb;
  B.n_patterns = (uint32_t)patterns.size();
  B.L = L;
  B.TP = TP;
  B.n_zero = zero_list.size();
  if (!upload) { B.ready = true; return B200MF_OK; }

  // ---- device tables per pattern (stride PS = L^3 rounded up to 16):
  //   own_pos[PS]   uint16: lattice position of the i-th dof of the own range (0xffff: hole)
  //   other[PS]     uint32: entries of the nodes outside the own range, later touchers first,
  //                 then first touchers, then constrained nodes;  other_pos[PS] their positions
  //   phdr[4]       n_other, n_later, n_first, 0
  const size_t PS = (L3 + 15) / 16 * 16;
  std::vector<uint16_t> own_pos(patterns.size() * PS, 0xffffu), other_pos(patterns.size() * PS, 0);
  std::vector<uint32_t> other(patterns.size() * PS, 15u << 28), phdr(patterns.size() * 4, 0);
  for (size_t q = 0; q < patterns.size(); ++q) {
    const Pattern &P = patterns[q];
    std::vector<uint32_t> later, first, cons;
    for (uint32_t e = 0; e < L3; ++e) {
      const uint32_t en = P.entries[e], sl = en >> 28;
      if (sl == 0) own_pos[q * PS + (en & 0x7ffffffu)] = (uint16_t)e;
      else if (sl == 15) cons.push_back(e);
      else if (en & (1u << 27)) first.push_back(e);
      else later.push_back(e);
    }
    size_t k = 0;
    for (const auto *lst : {&later, &first, &cons})
      for (uint32_t e : *lst) {
        other[q * PS + k] = P.entries[e];
        other_pos[q * PS + k] = (uint16_t)e;
        ++k;
      }
    phdr[q * 4 + 0] = (uint32_t)k;
    phdr[q * 4 + 1] = (uint32_t)later.size();
    phdr[q * 4 + 2] = (uint32_t)first.size();
  }
  auto up16 = [&](uint16_t **dp, const std::vector<uint16_t> &v) -> int {
    B200MF_CUDA_CHECK(cudaMalloc((void **)dp, std::max<size_t>(v.size(), 1) * sizeof(uint16_t)));
    B200MF_CUDA_CHECK(cudaMemcpy(*dp, v.data(), v.size() * sizeof(uint16_t), cudaMemcpyHostToDevice));
    s.device_bytes += v.size() * sizeof(uint16_t);
    s.index_bytes += v.size() * sizeof(uint16_t);
    return B200MF_OK;
  };
  auto up = [&](uint32_t **dp, const std::vector<uint32_t> &v) -> int {
    B200MF_CUDA_CHECK(cudaMalloc((void **)dp, std::max<size_t>(v.size(), 1) * sizeof(uint32_t)));
    B200MF_CUDA_CHECK(cudaMemcpy(*dp, v.data(), v.size() * sizeof(uint32_t), cudaMemcpyHostToDevice));
    s.device_bytes += v.size() * sizeof(uint32_t);
    s.index_bytes += v.size() * sizeof(uint32_t);
    return B200MF_OK;
  };
  int rc;
  if ((rc = up(&B.d_desc, desc)) != B200MF_OK) return rc;
  if ((rc = up(&B.d_other, other)) != B200MF_OK) return rc;
  if ((rc = up(&B.d_phdr, phdr)) != B200MF_OK) return rc;
  if ((rc = up16(&B.d_own_pos, own_pos)) != B200MF_OK) return rc;
  if ((rc = up16(&B.d_other_pos, other_pos)) != B200MF_OK) return rc;
  if ((rc = up(&B.d_zero, zero_list)) != B200MF_OK) return rc;
  B200MF_CUDA_CHECK(cudaMalloc((void **)&B.d_flags, nb * sizeof(uint32_t)));
  B200MF_CUDA_CHECK(cudaMemset(B.d_flags, 0, nb * sizeof(uint32_t)));
  B200MF_CUDA_CHECK(cudaMalloc((void **)&B.d_ticket, 64));
  B200MF_CUDA_CHECK(cudaMemset(B.d_ticket, 0, 64));
  s.device_bytes += nb * sizeof(uint32_t) + 64;
  B.ready = true;
  return B200MF_OK;
}

} // namespace b200mf
