// Plane-per-thread fused cell kernel (3D, n = p+1 <= 6): the fast path of the operator kernel.
//
// Same algebra as cell_kernels.cuh (and as ApplyKernel + Portable::FEEvaluation of the
// reference, matrix_free/portable_matrix_free.templates.h:498-528,
// matrix_free/portable_fe_evaluation.h:363-535, portable_evaluation_kernels.h:504-650) but a
// different work decomposition, chosen from the ncu profile of the line-per-thread kernel
// (profiles/r01_v1_cell_loop_q4_f64_ncu.txt: shared-memory wavefronts at 81 % of peak, FP64
// pipe at 21 %):
//   * n threads work on one cell; a thread owns one n x n PLANE of the cell in registers and
//     applies the 1D matrices along BOTH in-plane directions without touching shared memory
//     (arithmetic intensity per shared-memory byte grows from n/16 to n/8 flop/B);
//   * the planes are re-cut through shared memory: "Y-planes" (thread <-> y, holds [z][x])
//     for the x and z sweeps, "Z-planes" (thread <-> z, holds [y][x]) for the y sweeps and the
//     in-plane derivatives;
//   * all n threads of a cell sit in one warp (in one 16-lane bank group for FP64), so the
//     only synchronisation is __syncwarp(): no CTA barrier in the loop, warps run free;
//   * the shared-memory layout idx = y + n*x + sz*z with sz = 1 (mod bank group) and cell
//     stride n*sz makes every access of both plane cuts bank-conflict free;
//   * Cartesian cells (diagonal metric) apply D^T c D direction by direction, which needs
//     8 plane transfers per cell; affine/general cells (full symmetric metric) need 12;
//   * persistent warps with a 3-stage software pipeline of asynchronous copies (cp.async):
//     while group g is computed, the dof values of group g+1 are gathered from src straight
//     into the padded shared-memory layout and the index list of group g+2 is copied in;
//     no global-memory latency is exposed and no registers are held by loads in flight;
//   * the quadrature weights are folded into the transposed derivative matrix (DtW) so the
//     Cartesian quadrature-point operation is one scalar per line.
#pragma once
#include "cell_kernels.cuh"

namespace b200mf {

template <int n, typename Number>
struct PlaneCfg {
  static constexpr int group = sizeof(Number) == 8 ? 16 : 32; // lanes served by one smem wavefront
  static constexpr int cpg = group / n;                       // cells per bank group
  static constexpr int groups = 32 / group;
  static constexpr int cpw = cpg * groups;                    // cells per warp
  static constexpr int nact = cpw * n;                        // working lanes per warp
  static constexpr int npc = n * n * n;
  static constexpr int sz = ((n * n - 1 + group) / group) * group + 1; // plane stride
  static constexpr int cs = n * sz;                                    // cell stride
  static constexpr int buf = cpw * cs;                                 // one data buffer of one warp
  static constexpr int gdofs = cpw * npc;                              // dofs of one warp's cell group
  static constexpr int pairs = (gdofs + 1) / 2;
  static constexpr int gather_iters = (pairs + nact - 1) / nact;       // index pairs per lane
  // per warp: three rotating data buffers (B0, B1, gather target of the next group) and three
  // rotating index buffers (this group: scatter; next: gather in flight; after next: loading)
  static constexpr size_t ibuf_bytes = ((sizeof(uint32_t) * (gdofs + 1) + 15) / 16) * 16;
  static constexpr size_t warp_bytes = 3 * (sizeof(Number) * buf + ibuf_bytes);
  static constexpr size_t table_bytes = ((sizeof(uint16_t) * (gdofs + 2) + 15) / 16) * 16;
  static constexpr int warps_fit = (int)((226 * 1024 - table_bytes) / warp_bytes);
  static constexpr int warps = warps_fit > 8 ? 8 : warps_fit;          // one CTA per SM
  static constexpr int threads = 32 * warps;
  static constexpr int cells = cpw * warps;                            // cells per CTA and pass
  static constexpr size_t smem_bytes = warp_bytes * warps + table_bytes;
  static constexpr unsigned lane_mask() {
    unsigned m = 0;
    for (int l = 0; l < 32; ++l)
      if ((l % group) / n < cpg) m |= 1u << l;
    return m;
  }
  static_assert(cpw <= kL2gPadCells, "index list padding too small");
};

__device__ __forceinline__ void cp_async_zfill(void *smem_dst, const void *gsrc, int bytes, bool zero) {
  const unsigned d = (unsigned)__cvta_generic_to_shared(smem_dst);
  const int src_size = zero ? 0 : bytes;
  if (bytes == 8)
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8, %2;" ::"r"(d), "l"(gsrc), "r"(src_size) : "memory");
  else
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4, %2;" ::"r"(d), "l"(gsrc), "r"(src_size) : "memory");
}
__device__ __forceinline__ void cp_async_copy(void *smem_dst, const void *gsrc, int bytes) {
  const unsigned d = (unsigned)__cvta_generic_to_shared(smem_dst);
  if (bytes == 8)
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(d), "l"(gsrc) : "memory");
  else
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(d), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_group 0;" ::: "memory"); }
__device__ __forceinline__ void prefetch_l2(const void *ptr) {
  asm volatile("prefetch.global.L2 [%0];" ::"l"(ptr));
}

template <typename Number> struct Vec2;
template <> struct Vec2<double> { using type = double2; };
template <> struct Vec2<float> { using type = float2; };

// in-register sweep over a plane a[i][j]: along j (ROWS) or along i (!ROWS)
template <int n, int sym, bool ROWS, typename Number>
__device__ __forceinline__ void plane_sweep(const EoMatrix<Number, n> &M, Number (&a)[n][n]) {
#pragma unroll
  for (int l = 0; l < n; ++l) {
    Number in[n], out[n];
#pragma unroll
    for (int k = 0; k < n; ++k) in[k] = ROWS ? a[l][k] : a[k][l];
    apply_eo<Number, n, sym>(M, in, out);
#pragma unroll
    for (int k = 0; k < n; ++k) {
      if (ROWS) a[l][k] = out[k];
      else      a[k][l] = out[k];
    }
  }
}

// DOT: also accumulate src . (A src) of the processed cells into *p.dot_accum (the p.Ap of CG,
// lac/solver_cg.h:739), evaluated at the quadrature points as sum_q grad u . (M grad u).
template <int n, typename Number, int KIND, bool DOT>
__global__ void __launch_bounds__(PlaneCfg<n, Number>::threads, 1)
cell_loop_plane_kernel(const __grid_constant__ CellKernelParams<3, n, Number, KIND> p) {
  using Cfg = PlaneCfg<n, Number>;
  constexpr int npc = Cfg::npc, sz = Cfg::sz, n2 = n * n;
  constexpr unsigned MASK = Cfg::lane_mask();
  constexpr uint32_t CBIT = B200MF_L2G_CONSTRAINED;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int grp = lane / Cfg::group, r = lane - grp * Cfg::group;
  const int cw = r / n, t = r - cw * n; // cell within the bank group, plane index
  const int ciw = grp * Cfg::cpg + cw;  // cell within the warp
  const int rank = grp * (Cfg::cpg * n) + r; // rank among the working lanes

  // shared memory: [offset table | warp 0: 3 data buffers, 3 index buffers | warp 1: ...]
  uint16_t *table = reinterpret_cast<uint16_t *>(smem_raw);
  unsigned char *wbase = smem_raw + Cfg::table_bytes + (size_t)warp * Cfg::warp_bytes;
  Number *wbuf = reinterpret_cast<Number *>(wbase);
  unsigned char *ibase = wbase + 3 * sizeof(Number) * Cfg::buf;
  // table[e] = padded offset of dof e = cell_in_warp * npc + local index of the warp's cell group
  for (int e = threadIdx.x; e < Cfg::gdofs + 2; e += Cfg::threads) {
    const int ee = e < Cfg::gdofs ? e : Cfg::gdofs - 1;
    const int c = ee / npc, loc = ee - c * npc;
    const int x = loc % n, y = (loc / n) % n, z = loc / n2;
    table[e] = (uint16_t)(c * Cfg::cs + y + n * x + sz * z);
  }
  __syncthreads();
  if (cw >= Cfg::cpg) return; // idle lanes (32 is not a multiple of n)

  const ShapeData<Number, n> &sh = p.shape;
  const bool has_mass = p.op.has_mass;
  const Number wt = sh.w[t]; // quadrature weight of this thread's plane coordinate
  constexpr int NS = 6;
  double dot = 0.0;

  const unsigned long long n_cells = p.cell_end - p.cell_begin;
  const unsigned long long n_groups = (n_cells + Cfg::cpw - 1) / Cfg::cpw;
  const unsigned long long g_stride = (unsigned long long)gridDim.x * Cfg::warps;
  unsigned long long g = (unsigned long long)blockIdx.x * Cfg::warps + warp;

  // stage 1 of the software pipeline: the contiguous index list of cell group gg -> shared
  // memory (coalesced asynchronous copies, no registers held).  The list is padded past the
  // last cell (kL2gPadCells) so whole groups can always be copied.
  auto issue_indices = [&](unsigned long long gg, uint32_t *I) {
    const uint32_t *base = p.l2g + (p.cell_begin + gg * Cfg::cpw) * npc;
    if ((reinterpret_cast<uintptr_t>(base) & 7) == 0) {
#pragma unroll
      for (int k = 0; k < Cfg::gather_iters; ++k) {
        const int e = 2 * (rank + Cfg::nact * k);
        if (k + 1 < Cfg::gather_iters || e < Cfg::gdofs) cp_async_copy(I + e, base + e, 8);
      }
    } else {
#pragma unroll
      for (int k = 0; k < Cfg::gather_iters; ++k) {
        const int e = 2 * (rank + Cfg::nact * k);
        if (k + 1 < Cfg::gather_iters || e < Cfg::gdofs) {
          cp_async_copy(I + e, base + e, 4);
          cp_async_copy(I + e + 1, base + e + 1, 4);
        }
      }
    }
  };
  // stage 2: read_dof_values of a group -- one asynchronous copy per dof from src into the
  // padded plane layout of data buffer U (constrained / padding entries are zero-filled)
  auto issue_gather = [&](Number *U, const uint32_t *I) {
    uint2 i2[Cfg::gather_iters];
    unsigned tb[Cfg::gather_iters];
#pragma unroll
    for (int k = 0; k < Cfg::gather_iters; ++k) { // all shared-memory reads first ...
      const int e = 2 * (rank + Cfg::nact * k);
      if (k + 1 < Cfg::gather_iters || e < Cfg::gdofs) {
        i2[k] = *reinterpret_cast<const uint2 *>(I + e);
        tb[k] = *reinterpret_cast<const unsigned *>(table + e);
      }
    }
#pragma unroll
    for (int k = 0; k < Cfg::gather_iters; ++k) { // ... then the asynchronous copies
      const int e = 2 * (rank + Cfg::nact * k);
      if (k + 1 < Cfg::gather_iters || e < Cfg::gdofs) {
        const bool z0 = (i2[k].x & CBIT) != 0, z1 = (i2[k].y & CBIT) != 0;
        cp_async_zfill(U + (tb[k] & 0xffffu), p.src + (i2[k].x & ~CBIT), sizeof(Number), z0);
        if (Cfg::gdofs % 2 == 0 || e + 1 < Cfg::gdofs)
          cp_async_zfill(U + (tb[k] >> 16), p.src + (i2[k].y & ~CBIT), sizeof(Number), z1);
      }
    }
  };
  auto dbuf = [&](int i) { return wbuf + i * Cfg::buf; };
  auto ibuf = [&](int i) { return reinterpret_cast<uint32_t *>(ibase + i * Cfg::ibuf_bytes); };

  // metric of the (very common) single-geometry mesh: loaded once
  Number m0[NS], det0 = Number(1);
#pragma unroll
  for (int s = 0; s < NS; ++s) m0[s] = Number(0);
  if (KIND != B200MF_CELLS_GENERAL && p.geom_id == nullptr) {
    if (KIND == B200MF_CELLS_CARTESIAN) {
      m0[0] = p.geom_table[0]; m0[1] = p.geom_table[1]; m0[2] = p.geom_table[2]; det0 = p.geom_table[3];
    } else {
#pragma unroll
      for (int s = 0; s < NS; ++s) m0[s] = p.geom_table[s];
      det0 = p.geom_table[NS];
    }
  }

  // pipeline prologue
  int rot = 0; // B0 = dbuf(rot), B1 = dbuf(rot+1), next gather -> dbuf(rot+2); same for ibuf
  if (g < n_groups) {
    issue_indices(g, ibuf(0));
    cp_async_commit();
    cp_async_wait_all();
    __syncwarp(MASK);
    issue_gather(dbuf(0), ibuf(0));
    if (g + g_stride < n_groups) issue_indices(g + g_stride, ibuf(2));
    cp_async_commit();
  }

  for (; g < n_groups; g += g_stride) {
    const unsigned long long cell_raw = p.cell_begin + g * Cfg::cpw + ciw;
    const bool valid = cell_raw < p.cell_end;
    double dit = 0.0; // this group's contribution to src . A src
    const unsigned long long cell = valid ? cell_raw : p.cell_end - 1; // geometry reads stay in range
    const int r0 = rot, r1 = (rot + 1) % 3, r2 = (rot + 2) % 3;
    rot = r2; // the next group's data lands in r2 and becomes its B0
    Number *B0 = dbuf(r0) + ciw * Cfg::cs;
    Number *B1 = dbuf(r1) + ciw * Cfg::cs;
    // plane cuts: Y-plane element (z,x) of thread t -> t + n*x + sz*z
    //             Z-plane element (y,x) of thread t -> y + n*x + sz*t
    Number *By0 = B0 + t, *By1 = B1 + t;
    Number *Bz0 = B0 + sz * t, *Bz1 = B1 + sz * t;
    const uint32_t *I0 = ibuf(r0) + ciw * npc + t * n2; // this thread's Z-plane indices

    // data of this group and indices of the next were issued one iteration ago
    cp_async_wait_all();
    __syncwarp(MASK);
    {
      const unsigned long long g1 = g + g_stride, g2 = g1 + g_stride;
      if (g1 < n_groups) {
        issue_gather(dbuf(r2), ibuf(r2));
        if (g2 < n_groups) issue_indices(g2, ibuf(r1));
        if (KIND == B200MF_CELLS_GENERAL) {
          // pull the metric of the next group towards L2
          const char *pm = reinterpret_cast<const char *>(p.metric + (p.cell_begin + g1 * Cfg::cpw) * (NS * npc));
          const unsigned long long rem = p.cell_end - (p.cell_begin + g1 * Cfg::cpw);
          const int nc = rem < (unsigned long long)Cfg::cpw ? (int)rem : Cfg::cpw;
          const int lines = (int)((nc * NS * npc * sizeof(Number) + 127) / 128);
#pragma unroll 1
          for (int l = rank; l < lines; l += Cfg::nact) prefetch_l2(pm + (size_t)l * 128);
        }
      }
      cp_async_commit();
    }

    // per-cell metric (Cartesian: diagonal of JxW J^-1 J^-T / w; affine: upper triangle)
    Number m[NS];
    Number det = det0;
#pragma unroll
    for (int s = 0; s < NS; ++s) m[s] = m0[s];
    if (KIND != B200MF_CELLS_GENERAL && p.geom_id != nullptr) {
      const unsigned gi = p.geom_id[cell];
      if (KIND == B200MF_CELLS_CARTESIAN) {
        const Number *tb = p.geom_table + gi * 4;
        m[0] = tb[0]; m[1] = tb[1]; m[2] = tb[2]; det = tb[3];
      } else {
        const Number *tb = p.geom_table + gi * (NS + 1);
#pragma unroll
        for (int s = 0; s < NS; ++s) m[s] = tb[s];
        det = tb[NS];
      }
    }

    // ================= phase Y1: S along x and z on the gathered values ===================
    {
      Number a[n][n]; // [z][x] at y = t
#pragma unroll
      for (int z = 0; z < n; ++z)
#pragma unroll
        for (int x = 0; x < n; ++x) a[z][x] = By0[n * x + sz * z];
      plane_sweep<n, 1, true>(sh.S, a);
      plane_sweep<n, 1, false>(sh.S, a);
#pragma unroll
      for (int z = 0; z < n; ++z)
#pragma unroll
        for (int x = 0; x < n; ++x) By0[n * x + sz * z] = a[z][x];
    }
    __syncwarp(MASK);

    if constexpr (KIND == B200MF_CELLS_CARTESIAN) {
      // =============== phase Z1: S along y, D^T c D along x and y =========================
      {
        Number u[n][n], v[n][n]; // [y][x] at z = t
#pragma unroll
        for (int y = 0; y < n; ++y)
#pragma unroll
          for (int x = 0; x < n; ++x) u[y][x] = Bz0[y + n * x];
        plane_sweep<n, 1, false>(sh.S, u);
        const Number *gc = p.op.grad_coef ? p.op.grad_coef + cell * npc + t * n2 : nullptr;
        const Number ax = p.op.grad_const * m[0] * wt, ay = p.op.grad_const * m[1] * wt;
        // x direction, row by row: v = s_y (D^T W)(c .* D u)
#pragma unroll
        for (int y = 0; y < n; ++y) {
          Number in[n], gq[n], o[n];
#pragma unroll
          for (int x = 0; x < n; ++x) in[x] = u[y][x];
          apply_eo<Number, n, -1>(sh.D, in, gq);
          const Number s = ax * sh.w[y];
          if (gc) {
#pragma unroll
            for (int x = 0; x < n; ++x) {
              const Number h = gq[x] * gc[y * n + x];
              if (DOT) dit += double(s * sh.w[x]) * double(gq[x]) * double(h);
              gq[x] = h;
            }
          } else if (DOT) {
#pragma unroll
            for (int x = 0; x < n; ++x) dit += double(s * sh.w[x]) * double(gq[x]) * double(gq[x]);
          }
          apply_eo<Number, n, -1>(sh.DtW, gq, o);
#pragma unroll
          for (int x = 0; x < n; ++x) v[y][x] = s * o[x];
        }
        // y direction, column by column
#pragma unroll
        for (int x = 0; x < n; ++x) {
          Number in[n], gq[n], o[n];
#pragma unroll
          for (int y = 0; y < n; ++y) in[y] = u[y][x];
          apply_eo<Number, n, -1>(sh.D, in, gq);
          const Number s = ay * sh.w[x];
          if (gc) {
#pragma unroll
            for (int y = 0; y < n; ++y) {
              const Number h = gq[y] * gc[y * n + x];
              if (DOT) dit += double(s * sh.w[y]) * double(gq[y]) * double(h);
              gq[y] = h;
            }
          } else if (DOT) {
#pragma unroll
            for (int y = 0; y < n; ++y) dit += double(s * sh.w[y]) * double(gq[y]) * double(gq[y]);
          }
          apply_eo<Number, n, -1>(sh.DtW, gq, o);
#pragma unroll
          for (int y = 0; y < n; ++y) v[y][x] += s * o[y];
        }
        if (has_mass) {
          const Number *mc = p.op.mass_coef ? p.op.mass_coef + cell * npc + t * n2 : nullptr;
          const Number dw = det * wt;
#pragma unroll
          for (int y = 0; y < n; ++y)
#pragma unroll
            for (int x = 0; x < n; ++x) {
              Number cm = p.op.mass_const;
              if (mc) cm += mc[y * n + x];
              const Number h = cm * dw * sh.w2[y * n + x] * u[y][x];
              if (DOT) dit += double(u[y][x]) * double(h);
              v[y][x] += h;
            }
        }
#pragma unroll
        for (int y = 0; y < n; ++y)
#pragma unroll
          for (int x = 0; x < n; ++x) {
            Bz0[y + n * x] = u[y][x];
            Bz1[y + n * x] = v[y][x];
          }
      }
      __syncwarp(MASK);
      // =============== phase Y2: D^T c D along z, S^T along z and x =======================
      {
        Number a[n][n]; // [z][x] at y = t
        const Number *gc = p.op.grad_coef ? p.op.grad_coef + cell * npc + t * n : nullptr;
        const Number az = p.op.grad_const * m[2] * wt;
#pragma unroll
        for (int x = 0; x < n; ++x) {
          Number in[n], gq[n], o[n];
#pragma unroll
          for (int z = 0; z < n; ++z) in[z] = By0[n * x + sz * z];
          apply_eo<Number, n, -1>(sh.D, in, gq);
          const Number s = az * sh.w[x];
          if (gc) {
#pragma unroll
            for (int z = 0; z < n; ++z) {
              const Number h = gq[z] * gc[z * n2 + x];
              if (DOT) dit += double(s * sh.w[z]) * double(gq[z]) * double(h);
              gq[z] = h;
            }
          } else if (DOT) {
#pragma unroll
            for (int z = 0; z < n; ++z) dit += double(s * sh.w[z]) * double(gq[z]) * double(gq[z]);
          }
          apply_eo<Number, n, -1>(sh.DtW, gq, o);
#pragma unroll
          for (int z = 0; z < n; ++z) a[z][x] = s * o[z] + By1[n * x + sz * z];
        }
        plane_sweep<n, 1, false>(sh.St, a);
        plane_sweep<n, 1, true>(sh.St, a);
#pragma unroll
        for (int z = 0; z < n; ++z)
#pragma unroll
          for (int x = 0; x < n; ++x) By0[n * x + sz * z] = a[z][x];
      }
      __syncwarp(MASK);
    } else {
      // =============== full symmetric metric (affine / general cells) =====================
      Number gx[n][n], gy[n][n]; // [y][x] at z = t, live across phase Y2
      {
        Number u[n][n];
#pragma unroll
        for (int y = 0; y < n; ++y)
#pragma unroll
          for (int x = 0; x < n; ++x) u[y][x] = Bz0[y + n * x];
        plane_sweep<n, 1, false>(sh.S, u);
#pragma unroll
        for (int y = 0; y < n; ++y) {
          Number in[n], gq[n];
#pragma unroll
          for (int x = 0; x < n; ++x) in[x] = u[y][x];
          apply_eo<Number, n, -1>(sh.D, in, gq);
#pragma unroll
          for (int x = 0; x < n; ++x) gx[y][x] = gq[x];
        }
#pragma unroll
        for (int x = 0; x < n; ++x) {
          Number in[n], gq[n];
#pragma unroll
          for (int y = 0; y < n; ++y) in[y] = u[y][x];
          apply_eo<Number, n, -1>(sh.D, in, gq);
#pragma unroll
          for (int y = 0; y < n; ++y) gy[y][x] = gq[y];
        }
#pragma unroll
        for (int y = 0; y < n; ++y)
#pragma unroll
          for (int x = 0; x < n; ++x) Bz0[y + n * x] = u[y][x];
      }
      __syncwarp(MASK);
      // metric rows of this thread's z-plane (layout metric_offset: [cell][z][y][s][x]) are
      // streamed through registers one row ahead of their use, with wide loads
      using V2 = typename Vec2<Number>::type;
      constexpr int RV = 3 * n; // 6n numbers per row = 3n two-element vectors
      V2 mrow[2][RV];
      const V2 *mbase = (KIND == B200MF_CELLS_GENERAL)
                            ? reinterpret_cast<const V2 *>(p.metric + ((cell * n + t) * n) * (NS * n))
                            : nullptr;
      auto load_row = [&](int y, V2 (&dst)[RV]) {
#pragma unroll
        for (int i = 0; i < RV; ++i) dst[i] = __ldg(mbase + y * RV + i);
      };
      if (KIND == B200MF_CELLS_GENERAL) load_row(0, mrow[0]);
      // ---- phase Y2: reference z-derivative -> B1
#pragma unroll
      for (int x = 0; x < n; ++x) {
        Number in[n], gq[n];
#pragma unroll
        for (int z = 0; z < n; ++z) in[z] = By0[n * x + sz * z];
        apply_eo<Number, n, -1>(sh.D, in, gq);
#pragma unroll
        for (int z = 0; z < n; ++z) By1[n * x + sz * z] = gq[z];
      }
      __syncwarp(MASK);
      // ---- phase Z2: quadrature-point operator, D^T along x and y
      {
        Number v[n][n];
        const unsigned long long q0 = cell * npc + t * n2;
        const Number *gc = p.op.grad_coef ? p.op.grad_coef + q0 : nullptr;
        const Number *mc = p.op.mass_coef ? p.op.mass_coef + q0 : nullptr;
#pragma unroll
        for (int y = 0; y < n; ++y) {
          Number hx[n], o[n];
          if (KIND == B200MF_CELLS_GENERAL && y + 1 < n) load_row(y + 1, mrow[(y + 1) & 1]);
          const Number *mr = reinterpret_cast<const Number *>(mrow[y & 1]); // [s][x]
#pragma unroll
          for (int x = 0; x < n; ++x) {
            const Number gz = Bz1[y + n * x];
            Number cg = p.op.grad_const;
            if (gc) cg *= gc[y * n + x];
            Number jxw;
            if (KIND == B200MF_CELLS_GENERAL) {
#pragma unroll
              for (int s = 0; s < NS; ++s) m[s] = mr[s * n + x];
              jxw = has_mass ? p.jxw[q0 + y * n + x] : Number(0);
            } else {
              const Number wq = wt * sh.w2[y * n + x];
              cg *= wq;
              jxw = det * wq;
            }
            const Number a0 = gx[y][x], a1 = gy[y][x];
            const Number h0 = cg * (m[0] * a0 + m[1] * a1 + m[2] * gz);
            const Number h1 = cg * (m[1] * a0 + m[3] * a1 + m[4] * gz);
            const Number h2 = cg * (m[2] * a0 + m[4] * a1 + m[5] * gz);
            if (DOT) dit += double(a0) * double(h0) + double(a1) * double(h1) + double(gz) * double(h2);
            hx[x] = h0;
            gy[y][x] = h1;
            Bz1[y + n * x] = h2;
            Number vm = Number(0);
            if (has_mass) {
              Number cm = p.op.mass_const;
              if (mc) cm += mc[y * n + x];
              const Number uq = Bz0[y + n * x];
              vm = cm * jxw * uq;
              if (DOT) dit += double(uq) * double(vm);
            }
            gx[y][x] = vm; // reuse as the value-term accumulator
          }
          apply_eo<Number, n, -1>(sh.Dt, hx, o);
#pragma unroll
          for (int x = 0; x < n; ++x) v[y][x] = o[x] + gx[y][x];
        }
#pragma unroll
        for (int x = 0; x < n; ++x) {
          Number in[n], o[n];
#pragma unroll
          for (int y = 0; y < n; ++y) in[y] = gy[y][x];
          apply_eo<Number, n, -1>(sh.Dt, in, o);
#pragma unroll
          for (int y = 0; y < n; ++y) v[y][x] += o[y];
        }
#pragma unroll
        for (int y = 0; y < n; ++y)
#pragma unroll
          for (int x = 0; x < n; ++x) Bz0[y + n * x] = v[y][x];
      }
      __syncwarp(MASK);
      // ---- phase Y3: D^T along z, S^T along z and x
      {
        Number a[n][n];
#pragma unroll
        for (int x = 0; x < n; ++x) {
          Number in[n], o[n];
#pragma unroll
          for (int z = 0; z < n; ++z) in[z] = By1[n * x + sz * z];
          apply_eo<Number, n, -1>(sh.Dt, in, o);
#pragma unroll
          for (int z = 0; z < n; ++z) a[z][x] = o[z] + By0[n * x + sz * z];
        }
        plane_sweep<n, 1, false>(sh.St, a);
        plane_sweep<n, 1, true>(sh.St, a);
#pragma unroll
        for (int z = 0; z < n; ++z)
#pragma unroll
          for (int x = 0; x < n; ++x) By0[n * x + sz * z] = a[z][x];
      }
      __syncwarp(MASK);
    }

    // ================= phase Z-last: S^T along y, scatter =================================
    {
      Number u[n][n];
#pragma unroll
      for (int y = 0; y < n; ++y)
#pragma unroll
        for (int x = 0; x < n; ++x) u[y][x] = Bz0[y + n * x];
      plane_sweep<n, 1, false>(sh.St, u);
      // branch-free: constrained / padding / out-of-range entries add 0 to a valid address
      uint32_t idx[n][n];
#pragma unroll
      for (int y = 0; y < n; ++y)
#pragma unroll
        for (int x = 0; x < n; ++x) idx[y][x] = I0[y * n + x];
#pragma unroll
      for (int y = 0; y < n; ++y)
#pragma unroll
        for (int x = 0; x < n; ++x) {
          const bool live = valid && !(idx[y][x] & CBIT);
          atomicAdd(p.dst + (idx[y][x] & ~CBIT), live ? u[y][x] : Number(0));
        }
    }
    if (DOT && valid) dot += dit;
    __syncwarp(MASK); // all reads of this iteration done before the buffers rotate
  } // group loop
  if (DOT && p.dot_accum != nullptr) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      // working lanes only; a partner that has exited contributes 0
      const double other = __shfl_xor_sync(MASK, dot, o);
      if ((MASK >> ((lane ^ o) & 31)) & 1u) dot += other;
    }
    if (lane == 0) atomicAdd(p.dot_accum, dot);
  }
}

} // namespace b200mf
