// Brick kernel: the fast path of the operator kernel for Cartesian cells with constant
// coefficients (the headline case: 3D Laplace / Helmholtz on a uniformly refined hyper_cube).
//
// Same operator as cell_loop_kernel / cell_loop_plane_kernel (ApplyKernel + FEEvaluation of
// matrix_free/portable_matrix_free.templates.h:498-528, portable_fe_evaluation.h:363-745),
// restructured around what bounds it on B200 (profiles/r01_plane_cell_loop_q4_f64_ncu.txt:
// FP64 pipe 36 %, 2.35x the algorithmic DRAM traffic from index lists and atomics):
//
//  * A "brick" is an aligned window of b^3 cells that are consecutive in the (Morton-ordered)
//    cell array and form a b x b x b block; its dofs form an L^3 lattice, L = b p + 1.  The
//    setup (brick_setup.cpp) detects bricks from local_to_global alone and stores ONE index per
//    lattice node (4 L^3 bytes per brick instead of 4 b^3 (p+1)^3), flagged "complete" when no
//    cell outside the brick touches the dof.
//  * One CTA gathers the lattice once (shared dofs are read once, not once per cell), and
//    applies the operator to the whole brick as one macro element: on a Cartesian cell with a
//    constant coefficient the cell matrix is  Kx (x) M (x) M + M (x) Ky (x) M + M (x) M (x) Kz'
//    with the 1D mass matrix M = S W S^T and stiffness matrix K = (S D) W (S D)^T scaled by the
//    cell's metric (and Kz' = Kz + c_mass det M), i.e. evaluate/quadrature/integrate collapse
//    to 7 1D sweeps in nodal space instead of 12 sweeps through the quadrature points:
//        A = Mx u, B = Kx u;   C = My A, D = Ky A + My B;   v = Kz' C + Mz D.
//    One thread owns one lattice line per sweep (L^2 lines), walks its b cell blocks with the
//    (p+1)x(p+1) matrices in even-odd form (both are symmetric and centro-symmetric) taken from
//    the constant bank, and sums the shared end points of neighbouring blocks in registers: no
//    scatter conflicts inside the brick, two lattice-sized shared-memory arrays, three CTA
//    barriers per brick, all shared-memory accesses conflict free (lines are unit-stride across
//    the threads, or L-strided with L odd).
//  * Results of complete dofs (all of the lattice interior) are written with plain stores;
//    only the brick surface uses atomics (RED.ADD): 1 - ((L-2)/(L-1))^3 of the dofs.
#pragma once
#include "cell_kernels.cuh"

namespace b200mf {

#define B200MF_MAP_COMPLETE 0x40000000u
#define B200MF_MAP_FIRST 0x20000000u
#define B200MF_MAP_INDEX 0x1fffffffu

template <int p, int b, typename Number>
struct BrickCfg {
  static constexpr int n = p + 1;
  static constexpr int L = b * p + 1;
  static constexpr int L2 = L * L, L3 = L * L * L;
  static constexpr int W = b * b * b; // cells per brick
  static constexpr int threads = ((L2 + 31) / 32) * 32;
  static constexpr size_t smem_bytes = 2 * sizeof(Number) * L3;
  static constexpr int gather_iters = (L3 + threads - 1) / threads;
  // resident CTAs per SM the kernel is compiled for: what shared memory allows, as long as that
  // leaves ~96 registers per thread (degrees whose sweeps need more run one CTA per SM)
  static constexpr int by_smem = (int)((227 * 1024) / (smem_bytes + 1024));
  static constexpr int by_regs = n <= 5 ? 65536 / (threads * 96) : 1;
  static constexpr int ctas_min = by_smem < by_regs ? by_smem : by_regs;
  static constexpr int ctas_per_sm = ctas_min < 1 ? 1 : ctas_min;
};

template <typename Number, int n>
struct EoHalf { // even / odd parts of one input line
  static constexpr int h = n / 2;
  Number e[h > 0 ? h : 1], o[h > 0 ? h : 1], c;
};

template <typename Number, int n>
__device__ __forceinline__ void eo_split(const Number (&in)[n], EoHalf<Number, n> &x) {
  constexpr int h = n / 2;
#pragma unroll
  for (int i = 0; i < h; ++i) {
    x.e[i] = in[i] + in[n - 1 - i];
    x.o[i] = in[i] - in[n - 1 - i];
  }
  x.c = (n % 2) ? in[h] : Number(0);
}

template <typename Number, int n>
struct EoAcc { // even / odd parts of one output line
  static constexpr int h = n / 2, hq = (n + 1) / 2;
  Number X[hq], Y[h > 0 ? h : 1];
};

// acc (=, +=) Mat * x
template <bool first, typename Number, int n>
__device__ __forceinline__ void eo_mac(const EoMatrix<Number, n> &M, const EoHalf<Number, n> &x,
                                       EoAcc<Number, n> &acc) {
  constexpr int h = n / 2, hq = (n + 1) / 2;
#pragma unroll
  for (int q = 0; q < hq; ++q) {
    Number X = first ? M.E[q] * x.e[0] : acc.X[q] + M.E[q] * x.e[0];
#pragma unroll
    for (int i = 1; i < h; ++i) X += M.E[i * hq + q] * x.e[i];
    if (n % 2) X += M.mid[q] * x.c;
    acc.X[q] = X;
  }
#pragma unroll
  for (int q = 0; q < h; ++q) {
    Number Y = first ? M.O[q] * x.o[0] : acc.Y[q] + M.O[q] * x.o[0];
#pragma unroll
    for (int i = 1; i < h; ++i) Y += M.O[i * h + q] * x.o[i];
    acc.Y[q] = Y;
  }
}

template <typename Number, int n>
__device__ __forceinline__ void eo_join(const EoAcc<Number, n> &acc, Number (&out)[n]) {
  constexpr int h = n / 2;
#pragma unroll
  for (int q = 0; q < h; ++q) {
    out[q] = acc.X[q] + acc.Y[q];
    out[n - 1 - q] = acc.X[q] - acc.Y[q];
  }
  if (n % 2) out[h] = acc.X[h];
}

template <int p, typename Number>
struct BrickKernelParams {
  BrickMatrices<Number, p + 1> mat;
  const uint32_t *map; // [n_bricks][L^3]: local dof index | B200MF_MAP_COMPLETE | B200MF_L2G_CONSTRAINED
  const Number *src;
  Number *dst;
  double *dot_accum; // optional: += src . (A src) over these bricks
  unsigned long long brick_begin;
  const uint32_t *list; // optional: brick ids of this launch (coloured launches)
  // STRIDED kernels: per brick {base, stride_y, stride_z, flags}: the lattice is an affine image of the
  // numbering (index = base + x + stride_y y + stride_z z, e.g. after DoFRenumbering::lexicographic), so
  // no index map is read; flags: bit f = lattice face f (x-,x+,y-,y+,z-,z+) is shared with cells outside
  // the brick, bit 6+f = the dofs of face f are constrained
  const uint4 *strided;
  // 0: add into dst; 1: dst is known to be zero (vmult) -> complete dofs are stored, the others use
  // RED; 2: coloured launch: complete / first-toucher dofs are stored, the others load-add-store
  int overwrite;
};

// lattice face flags of a node
template <int L>
__device__ __forceinline__ bool on_flagged_face(uint32_t bits, int x, int y, int z) {
  return ((bits & 1u) && x == 0) || ((bits & 2u) && x == L - 1) || ((bits & 4u) && y == 0) ||
         ((bits & 8u) && y == L - 1) || ((bits & 16u) && z == 0) || ((bits & 32u) && z == L - 1);
}

template <int p, int b, typename Number, bool DOT, bool STRIDED>
// (the register bound is only needed by the DOT variant; without it ptxas picks 64 registers for
// the plain one, which measures 4 % faster than the 96 it takes when allowed to)
__global__ void __launch_bounds__(BrickCfg<p, b, Number>::threads, DOT ? BrickCfg<p, b, Number>::ctas_per_sm : 0)
brick_cartesian_kernel(const __grid_constant__ BrickKernelParams<p, Number> prm) {
  using Cfg = BrickCfg<p, b, Number>;
  constexpr int n = p + 1, L = Cfg::L, L2 = Cfg::L2, L3 = Cfg::L3, T = Cfg::threads;
  constexpr uint32_t CBIT = B200MF_L2G_CONSTRAINED;
  extern __shared__ __align__(16) unsigned char brick_smem[];
  Number *P0 = reinterpret_cast<Number *>(brick_smem);
  Number *P1 = P0 + L3;
  const int tid = threadIdx.x;
  const unsigned long long brick = prm.list ? (unsigned long long)__ldg(prm.list + prm.brick_begin + blockIdx.x)
                                            : prm.brick_begin + blockIdx.x;
  const uint32_t *__restrict__ map = prm.map + brick * (unsigned long long)L3;
  const Number *__restrict__ src = prm.src;
  uint4 sd = make_uint4(0u, 0u, 0u, 0u);
  if (STRIDED) sd = __ldg(prm.strided + brick);

  // ---- read_dof_values of the whole brick: every lattice node once
  if (STRIDED) {
    // computed indices: thread <-> (x, y) walks its z column, consecutive lanes read consecutive dofs
    // of a lattice x-line; one multiply-add per node, no index map
    if (tid < L2) {
      const int gx = tid % L, gy = tid / L;
      const uint32_t line = sd.x + gx + sd.y * gy;
      const uint32_t cbits = sd.w >> 6;
      const bool cons_xy = cbits != 0u && on_flagged_face<L>(cbits & 15u, gx, gy, 1);
      Number val[L];
#pragma unroll
      for (int z = 0; z < L; ++z) {
        const bool cons = cons_xy || ((cbits & 16u) && z == 0) || ((cbits & 32u) && z == L - 1);
        val[z] = cons ? Number(0) : __ldg(src + line + sd.z * z);
      }
#pragma unroll
      for (int z = 0; z < L; ++z) P0[tid + z * L2] = val[z];
    }
  } else {
    uint32_t idx[Cfg::gather_iters];
#pragma unroll
    for (int k = 0; k < Cfg::gather_iters; ++k) {
      const int e = tid + k * T;
      idx[k] = (e < L3) ? __ldg(map + e) : CBIT;
    }
    Number val[Cfg::gather_iters];
#pragma unroll
    for (int k = 0; k < Cfg::gather_iters; ++k)
      val[k] = (idx[k] & CBIT) ? Number(0) : __ldg(src + (idx[k] & B200MF_MAP_INDEX));
#pragma unroll
    for (int k = 0; k < Cfg::gather_iters; ++k) {
      const int e = tid + k * T;
      if (e < L3) P0[e] = val[k];
    }
  }
  __syncthreads();

  const bool active = tid < L2;
  const int la = tid % L, lb = tid / L;

  // ---- x sweep: A = Mx u -> P0 (in place), B = Kx u -> P1; thread <-> (y, z)
  if (active) {
    Number *l0 = P0 + L * tid, *l1 = P1 + L * tid;
    Number in[n], cA = Number(0), cB = Number(0);
    in[0] = l0[0];
#pragma unroll
    for (int c = 0; c < b; ++c) {
#pragma unroll
      for (int k = 1; k < n; ++k) in[k] = l0[c * p + k];
      EoHalf<Number, n> x;
      eo_split<Number, n>(in, x);
      EoAcc<Number, n> a;
      Number oA[n], oB[n];
      eo_mac<true, Number, n>(prm.mat.M, x, a);
      eo_join<Number, n>(a, oA);
      eo_mac<true, Number, n>(prm.mat.Kx, x, a);
      eo_join<Number, n>(a, oB);
      if (c > 0) { oA[0] += cA; oB[0] += cB; }
#pragma unroll
      for (int k = 0; k < p; ++k) {
        l0[c * p + k] = oA[k];
        l1[c * p + k] = oB[k];
      }
      cA = oA[p];
      cB = oB[p];
      in[0] = in[p];
    }
    l0[L - 1] = cA;
    l1[L - 1] = cB;
  }
  __syncthreads();

  // the indices of this thread's z line (the scatter targets): requested now, used after the
  // y sweep, so the z sweep never waits for them
  uint32_t mz[STRIDED ? 1 : L];
  // strided mode: the word of node z of this thread's line is computed where it is used (z is a
  // compile-time constant there)
  uint32_t s_line = 0u, s_shared_xy = 0u, s_cons_xy = 0u;
  if (STRIDED) {
    s_line = sd.x + la + sd.y * lb; // la = x, lb = y of this thread's z line
    s_shared_xy = on_flagged_face<L>(sd.w & 15u, la, lb, 1) ? 1u : 0u;
    s_cons_xy = ((sd.w >> 6) != 0u && on_flagged_face<L>((sd.w >> 6) & 15u, la, lb, 1)) ? 1u : 0u;
  }
  auto word = [&](int z) -> uint32_t {
    if (STRIDED) {
      const bool shared = s_shared_xy || ((sd.w & 16u) && z == 0) || ((sd.w & 32u) && z == L - 1);
      const bool cons = s_cons_xy || ((sd.w & (16u << 6)) && z == 0) || ((sd.w & (32u << 6)) && z == L - 1);
      return (s_line + sd.z * z) | (shared ? 0u : B200MF_MAP_COMPLETE) | (cons ? CBIT : 0u);
    }
    return mz[STRIDED ? 0 : z];
  };
  if (active && !STRIDED) {
#pragma unroll
    for (int z = 0; z < L; ++z) mz[STRIDED ? 0 : z] = __ldg(map + tid + z * L2);
    if (prm.overwrite == 2) {
      // coloured launch: the dofs an earlier launch stored are read back at scatter time; ask L2
      // for them now so that the reads do not sit on the critical path of the z sweep
#pragma unroll
      for (int z = 0; z < L; ++z)
        if ((word(z) & (CBIT | B200MF_MAP_COMPLETE | B200MF_MAP_FIRST)) == 0)
          asm volatile("prefetch.global.L2 [%0];" ::"l"(prm.dst + (word(z) & B200MF_MAP_INDEX)));
    }
  }

  // ---- y sweep: C = My A -> P0, D = Ky A + My B -> P1 (both in place); thread <-> (x, z)
  if (active) {
    Number *l0 = P0 + la + L2 * lb, *l1 = P1 + la + L2 * lb;
    Number inA[n], inB[n], cC = Number(0), cD = Number(0);
    inA[0] = l0[0];
    inB[0] = l1[0];
#pragma unroll
    for (int c = 0; c < b; ++c) {
#pragma unroll
      for (int k = 1; k < n; ++k) {
        inA[k] = l0[(c * p + k) * L];
        inB[k] = l1[(c * p + k) * L];
      }
      EoHalf<Number, n> xa, xb;
      eo_split<Number, n>(inA, xa);
      eo_split<Number, n>(inB, xb);
      EoAcc<Number, n> a;
      Number oC[n], oD[n];
      eo_mac<true, Number, n>(prm.mat.M, xa, a);
      eo_join<Number, n>(a, oC);
      eo_mac<true, Number, n>(prm.mat.Ky, xa, a);
      eo_mac<false, Number, n>(prm.mat.M, xb, a);
      eo_join<Number, n>(a, oD);
      if (c > 0) { oC[0] += cC; oD[0] += cD; }
#pragma unroll
      for (int k = 0; k < p; ++k) {
        l0[(c * p + k) * L] = oC[k];
        l1[(c * p + k) * L] = oD[k];
      }
      cC = oC[p];
      cD = oD[p];
      inA[0] = inA[p];
      inB[0] = inB[p];
    }
    l0[(L - 1) * L] = cC;
    l1[(L - 1) * L] = cD;
  }
  __syncthreads();

  // ---- z sweep: v = Kz' C + Mz D, written straight to dst; thread <-> (x, y)
  double dot = 0.0;
  if (active) {
    const Number *l0 = P0 + tid, *l1 = P1 + tid;
    Number *__restrict__ dst = prm.dst;
    // emit(): complete dofs are stored (vmult: dst is zero) or stored after adding the old value
    // (cell_loop adds into dst), everything else goes through RED.ADD
    auto emit = [&](uint32_t m, Number v, Number u, Number o) {
      if (!(m & CBIT)) {
        Number *d = dst + (m & B200MF_MAP_INDEX);
        if ((m & B200MF_MAP_COMPLETE) || prm.overwrite == 2) *d = v + o;
        else atomicAdd(d, v);
        if (DOT) dot += double(u) * double(v);
      }
    };
    Number inC[n], inD[n], cV = Number(0);
    inC[0] = l0[0];
    inD[0] = l1[0];
#pragma unroll
    for (int c = 0; c < b; ++c) {
      Number ui[p], old[p];
      if (DOT) {
#pragma unroll
        for (int k = 0; k < p; ++k)
          ui[k] = (word(c * p + k) & CBIT) ? Number(0) : __ldg(src + (word(c * p + k) & B200MF_MAP_INDEX));
      }
      // values already in dst that this brick adds to, requested before the block's arithmetic:
      // cell_loop mode adds into the complete dofs (the others go through RED), a coloured launch
      // adds into the dofs an earlier launch stored (no atomics: bricks of a colour share no dof)
      if (prm.overwrite != 1) {
        const uint32_t want = prm.overwrite == 2 ? 0u : B200MF_MAP_COMPLETE;
        const uint32_t bits = prm.overwrite == 2 ? (CBIT | B200MF_MAP_COMPLETE | B200MF_MAP_FIRST) : (CBIT | B200MF_MAP_COMPLETE);
#pragma unroll
        for (int k = 0; k < p; ++k)
          old[k] = (word(c * p + k) & bits) == want ? dst[word(c * p + k) & B200MF_MAP_INDEX] : Number(0);
      }
#pragma unroll
      for (int k = 1; k < n; ++k) {
        inC[k] = l0[(c * p + k) * L2];
        inD[k] = l1[(c * p + k) * L2];
      }
      EoHalf<Number, n> xc, xd;
      eo_split<Number, n>(inC, xc);
      eo_split<Number, n>(inD, xd);
      EoAcc<Number, n> a;
      Number oV[n];
      eo_mac<true, Number, n>(prm.mat.Kz, xc, a);
      eo_mac<false, Number, n>(prm.mat.M, xd, a);
      eo_join<Number, n>(a, oV);
      if (c > 0) oV[0] += cV;
#pragma unroll
      for (int k = 0; k < p; ++k)
        emit(word(c * p + k), oV[k], DOT ? ui[k] : Number(0), prm.overwrite == 1 ? Number(0) : old[k]);
      cV = oV[p];
      inC[0] = inC[p];
      inD[0] = inD[p];
    }
    {
      const uint32_t m = word(L - 1);
      Number u = Number(0);
      if (DOT && !(m & CBIT)) u = __ldg(src + (m & B200MF_MAP_INDEX));
      Number o = Number(0);
      if ((prm.overwrite == 0 && (m & (CBIT | B200MF_MAP_COMPLETE)) == B200MF_MAP_COMPLETE) ||
          (prm.overwrite == 2 && (m & (CBIT | B200MF_MAP_COMPLETE | B200MF_MAP_FIRST)) == 0))
        o = dst[m & B200MF_MAP_INDEX];
      emit(m, cV, u, o);
    }
  }
  if (DOT && prm.dot_accum != nullptr) {
    dot = block_sum(dot);
    if (tid == 0) atomicAdd(prm.dot_accum, dot);
  }
}

// ---------------------------------------------------------------------------------------------
// Strided bricks in z slabs: the same brick operator for numberings in which the lattice of a brick
// is an affine image of the dof indices (STRIDED above), with the brick processed in NS slabs of
// b / NS cell layers.  The x and y sweeps act inside one z plane, so only the planes of the current
// slab have to be resident: two arrays of L x L x (b p / NS + 1) values instead of L^3 -- 41.6 KB
// instead of 78.6 KB for Q4 FP64, i.e. four resident CTAs per SM instead of two for a kernel that is
// latency bound once the index traffic is gone (profiles/r02_brick_strided_q4_f64_ncu.txt:
// warps_active 30 %).  The z sweep carries its partial sums and the last plane's inputs across the
// slabs in registers, exactly as it carries them across the cell blocks inside a slab.
template <int p, int b, int NS, typename Number>
struct SlabCfg {
  static constexpr int n = p + 1;
  static constexpr int L = b * p + 1, L2 = L * L;
  static constexpr int CS = b / NS;           // cell layers per slab
  static constexpr int NZ = CS * p + 1;       // planes of the first slab (the others hold CS p)
  static constexpr int threads = ((L2 + 31) / 32) * 32;
  static constexpr size_t smem_bytes = 2 * sizeof(Number) * L2 * NZ;
};

template <int p, int b, int NS, typename Number, bool DOT>
__global__ void __launch_bounds__(SlabCfg<p, b, NS, Number>::threads, DOT ? 2 : 3)
brick_strided_slab_kernel(const __grid_constant__ BrickKernelParams<p, Number> prm) {
  using Cfg = SlabCfg<p, b, NS, Number>;
  constexpr int n = p + 1, L = Cfg::L, L2 = Cfg::L2, CS = Cfg::CS;
  constexpr uint32_t CBIT = B200MF_L2G_CONSTRAINED;
  extern __shared__ __align__(16) unsigned char brick_smem[];
  Number *P0 = reinterpret_cast<Number *>(brick_smem);
  Number *P1 = P0 + L2 * Cfg::NZ;
  const int tid = threadIdx.x;
  const unsigned long long brick = prm.list ? (unsigned long long)__ldg(prm.list + prm.brick_begin + blockIdx.x)
                                            : prm.brick_begin + blockIdx.x;
  const Number *__restrict__ src = prm.src;
  Number *__restrict__ dst = prm.dst;
  const uint4 sd = __ldg(prm.strided + brick);
  const bool col = tid < L2;                 // thread <-> lattice column (x, y)
  const int la = tid % L, lb = tid / L;
  const uint32_t line = sd.x + la + sd.y * lb;
  const uint32_t cbits = sd.w >> 6;
  const bool cons_xy = cbits != 0u && on_flagged_face<L>(cbits & 15u, la, lb, 1);
  const bool shared_xy = on_flagged_face<L>(sd.w & 15u, la, lb, 1);
  auto word = [&](int z) -> uint32_t {
    const bool shared = shared_xy || ((sd.w & 16u) && z == 0) || ((sd.w & 32u) && z == L - 1);
    const bool cons = cons_xy || ((cbits & 16u) && z == 0) || ((cbits & 32u) && z == L - 1);
    return (line + sd.z * z) | (shared ? 0u : B200MF_MAP_COMPLETE) | (cons ? CBIT : 0u);
  };
  double dot = 0.0;
  auto emit = [&](uint32_t m, Number v, Number u, Number o) {
    if (!(m & CBIT)) {
      Number *d = dst + (m & B200MF_MAP_INDEX);
      if (m & B200MF_MAP_COMPLETE) *d = v + o;
      else atomicAdd(d, v);
      if (DOT) dot += double(u) * double(v);
    }
  };
  // carried across the slabs by the z sweep of this thread's column
  Number inC[n], inD[n], cV = Number(0);

#pragma unroll
  for (int sl = 0; sl < NS; ++sl) {
    constexpr int dummy = 0;
    (void)dummy;
    const int zb = sl * CS * p + (sl > 0 ? 1 : 0); // first plane this slab loads
    const int nz = CS * p + (sl == 0 ? 1 : 0);     // number of planes it loads
    if (sl > 0) __syncthreads();                   // the previous slab's z sweep has read P0 / P1
    // ---- read_dof_values of the slab: thread (x, y) walks its column, lanes on consecutive dofs
    if (col) {
      Number val[Cfg::NZ];
#pragma unroll
      for (int zl = 0; zl < Cfg::NZ; ++zl)
        if (zl < nz) {
          const int z = zb + zl;
          const bool cons = cons_xy || ((cbits & 16u) && z == 0) || ((cbits & 32u) && z == L - 1);
          val[zl] = cons ? Number(0) : __ldg(src + line + sd.z * z);
        }
#pragma unroll
      for (int zl = 0; zl < Cfg::NZ; ++zl)
        if (zl < nz) P0[tid + zl * L2] = val[zl];
    }
    __syncthreads();
    const bool rowt = tid < L * nz;                // thread <-> lattice row inside the slab
    // ---- x sweep: A = Mx u -> P0 (in place), B = Kx u -> P1; thread <-> (y, zl)
    if (rowt) {
      Number *l0 = P0 + L * tid, *l1 = P1 + L * tid;
      Number in[n], cA = Number(0), cB = Number(0);
      in[0] = l0[0];
#pragma unroll
      for (int c = 0; c < b; ++c) {
#pragma unroll
        for (int k = 1; k < n; ++k) in[k] = l0[c * p + k];
        EoHalf<Number, n> x;
        eo_split<Number, n>(in, x);
        EoAcc<Number, n> a;
        Number oA[n], oB[n];
        eo_mac<true, Number, n>(prm.mat.M, x, a);
        eo_join<Number, n>(a, oA);
        eo_mac<true, Number, n>(prm.mat.Kx, x, a);
        eo_join<Number, n>(a, oB);
        if (c > 0) { oA[0] += cA; oB[0] += cB; }
#pragma unroll
        for (int k = 0; k < p; ++k) {
          l0[c * p + k] = oA[k];
          l1[c * p + k] = oB[k];
        }
        cA = oA[p];
        cB = oB[p];
        in[0] = in[p];
      }
      l0[L - 1] = cA;
      l1[L - 1] = cB;
    }
    __syncthreads();
    // ---- y sweep: C = My A -> P0, D = Ky A + My B -> P1 (in place); thread <-> (x, zl)
    if (rowt) {
      Number *l0 = P0 + la + L2 * lb, *l1 = P1 + la + L2 * lb;
      Number inA[n], inB[n], cC = Number(0), cD = Number(0);
      inA[0] = l0[0];
      inB[0] = l1[0];
#pragma unroll
      for (int c = 0; c < b; ++c) {
#pragma unroll
        for (int k = 1; k < n; ++k) {
          inA[k] = l0[(c * p + k) * L];
          inB[k] = l1[(c * p + k) * L];
        }
        EoHalf<Number, n> xa, xb;
        eo_split<Number, n>(inA, xa);
        eo_split<Number, n>(inB, xb);
        EoAcc<Number, n> a;
        Number oC[n], oD[n];
        eo_mac<true, Number, n>(prm.mat.M, xa, a);
        eo_join<Number, n>(a, oC);
        eo_mac<true, Number, n>(prm.mat.Ky, xa, a);
        eo_mac<false, Number, n>(prm.mat.M, xb, a);
        eo_join<Number, n>(a, oD);
        if (c > 0) { oC[0] += cC; oD[0] += cD; }
#pragma unroll
        for (int k = 0; k < p; ++k) {
          l0[(c * p + k) * L] = oC[k];
          l1[(c * p + k) * L] = oD[k];
        }
        cC = oC[p];
        cD = oD[p];
        inA[0] = inA[p];
        inB[0] = inB[p];
      }
      l0[(L - 1) * L] = cC;
      l1[(L - 1) * L] = cD;
    }
    __syncthreads();
    // ---- z sweep over the cell layers of this slab: v = Kz' C + Mz D, written straight to dst
    if (col) {
      const Number *l0 = P0 + tid, *l1 = P1 + tid;
      if (sl == 0) {
        inC[0] = l0[0];
        inD[0] = l1[0];
      }
#pragma unroll
      for (int cc = 0; cc < CS; ++cc) {
        const int c = sl * CS + cc; // cell layer of the brick
        Number ui[p], old[p];
        if (DOT) {
#pragma unroll
          for (int k = 0; k < p; ++k)
            ui[k] = (word(c * p + k) & CBIT) ? Number(0) : __ldg(src + (word(c * p + k) & B200MF_MAP_INDEX));
        }
        if (!prm.overwrite) {
#pragma unroll
          for (int k = 0; k < p; ++k)
            old[k] = (word(c * p + k) & (CBIT | B200MF_MAP_COMPLETE)) == B200MF_MAP_COMPLETE
                         ? dst[word(c * p + k) & B200MF_MAP_INDEX]
                         : Number(0);
        }
#pragma unroll
        for (int k = 1; k < n; ++k) {
          inC[k] = l0[(c * p + k - zb) * L2];
          inD[k] = l1[(c * p + k - zb) * L2];
        }
        EoHalf<Number, n> xc, xd;
        eo_split<Number, n>(inC, xc);
        eo_split<Number, n>(inD, xd);
        EoAcc<Number, n> a;
        Number oV[n];
        eo_mac<true, Number, n>(prm.mat.Kz, xc, a);
        eo_mac<false, Number, n>(prm.mat.M, xd, a);
        eo_join<Number, n>(a, oV);
        if (c > 0) oV[0] += cV;
#pragma unroll
        for (int k = 0; k < p; ++k)
          emit(word(c * p + k), oV[k], DOT ? ui[k] : Number(0), prm.overwrite ? Number(0) : old[k]);
        cV = oV[p];
        inC[0] = inC[p];
        inD[0] = inD[p];
      }
      if (sl == NS - 1) {
        const uint32_t m = word(L - 1);
        Number u = Number(0);
        if (DOT && !(m & CBIT)) u = __ldg(src + (m & B200MF_MAP_INDEX));
        Number o = Number(0);
        if (!prm.overwrite && (m & (CBIT | B200MF_MAP_COMPLETE)) == B200MF_MAP_COMPLETE) o = dst[m & B200MF_MAP_INDEX];
        emit(m, cV, u, o);
      }
    }
  }
  if (DOT && prm.dot_accum != nullptr) {
    dot = block_sum(dot);
    if (tid == 0) atomicAdd(prm.dot_accum, dot);
  }
}

} // namespace b200mf
