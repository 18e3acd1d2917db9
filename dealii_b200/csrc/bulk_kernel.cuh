// Bulk brick kernel: the vmult hot path for Cartesian cells with constant coefficients.
//
// Same macro-element operator as brick_kernel.cuh (7 nodal-space sweeps over the L^3 lattice of
// a brick of b^3 cells: ApplyKernel + FEEvaluation::evaluate/integrate of
// matrix_free/portable_matrix_free.templates.h:498-528, portable_fe_evaluation.h:444-535,
// collapsed for a constant Jacobian), but the data movement is rebuilt around what bounded that
// kernel (profiles/r01_brick_cell_loop_q4_f64_ncu.txt: L1 data pipe 91 %, 4 B/node index map,
// uncoalesced lattice-order gather/scatter, memset + RED.ADD on every shared dof):
//
//  * read_dof_values: the brick's OWN RANGE -- the contiguous run of dofs it touches first, 32 KB
//    for a Q4 brick under deal.II's first-touch numbering -- arrives with ONE bulk async copy
//    (cp.async.bulk, TMA engine, mbarrier completion) into shared memory in dof order; the
//    permutation to lattice order happens on the way into the registers of the x sweep through a
//    PATTERN table shared by all bricks with the same relative numbering (bulk_setup.cpp).  Only
//    the lattice nodes other bricks own (the low faces: 17 % for Q4) are gathered with indexed
//    loads, from a per-brick base + the pattern's offset: no per-node index map is streamed.
//  * distribute_local_to_global: results of the own range are permuted back into dof order in
//    shared memory and leave with ONE bulk async store.  No memset of dst, no atomics on them.
//  * conflicts (portable_matrix_free.templates.h:1060-1185 offers colouring or atomics): the
//    brick that touches a dof first STORES it, every later toucher waits for that brick's flag and
//    adds with RED.  Bricks are handed out through a ticket counter in execution order, so a
//    brick only ever waits for bricks that are already running: no deadlock, one launch, and
//    the boundary bricks of a partitioned mesh can be placed between two halves of the interior.
//  * persistent CTAs: the next ticket/descriptor is fetched while the current brick is swept and
//    its bulk load is issued as soon as the staging buffer is free.
#pragma once
#include "brick_kernel.cuh"

namespace b200mf {

constexpr int kBulkDescWords = 48;

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t *bar, int count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "BULK_WAIT:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra BULK_DONE;\n"
      "bra BULK_WAIT;\n"
      "BULK_DONE:\n"
      "}\n" ::"r"(smem_u32(bar)),
      "r"(parity)
      : "memory");
}
// global -> shared bulk copy, completion on the mbarrier (bytes: multiple of 16, both 16-B aligned)
__device__ __forceinline__ void bulk_load(void *smem_dst, const void *gsrc, uint32_t bytes, uint64_t *bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   smem_u32(smem_dst)),
               "l"(gsrc), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}
__device__ __forceinline__ void bulk_store(void *gdst, const void *smem_src, uint32_t bytes) {
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(gdst), "r"(smem_u32(smem_src)),
               "r"(bytes)
               : "memory");
  asm volatile("cp.async.bulk.commit_group;" ::: "memory");
}
__device__ __forceinline__ void bulk_store_wait() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ uint32_t ld_acquire(const uint32_t *p) {
  uint32_t v;
  asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_release(uint32_t *p, uint32_t v) {
  asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}

template <int p, int b, typename Number>
struct BulkCfg {
  static constexpr int n = p + 1;
  static constexpr int L = b * p + 1;
  static constexpr int L2 = L * L, L3 = L * L * L;
  static constexpr int threads = ((L2 + 31) / 32) * 32;
  static constexpr int plane_elems = ((L3 + 15) / 16) * 16; // 128-byte multiple for FP64
  static constexpr size_t smem_bytes = 2 * sizeof(Number) * plane_elems + 2 * kBulkDescWords * 4 + 64;
  static constexpr int by_smem = (int)((227 * 1024) / (smem_bytes + 1024));
  static constexpr int ctas_per_sm = by_smem < 1 ? 1 : (by_smem > 2 ? 2 : by_smem);
};

template <int p, typename Number>
struct BulkKernelParams {
  BrickMatrices<Number, p + 1> mat;
  const uint32_t *desc;  // [n_exec][48]: pattern, own_lo, own_count, n_deps, base[16], dep[26]
  const uint32_t *tx;    // [pattern][L][TP]  entry of node (x, tid % L, tid / L)
  const uint32_t *tz;    // [pattern][L][TP]  entry of node (tid % L, tid / L, z)
  const uint32_t *holes; // [pattern][1 + max_holes]
  uint32_t *flags;       // [n_exec] epoch of the launch in which the brick stored its dofs
  uint32_t *ticket;
  const Number *src;
  Number *dst;
  double *dot_accum;
  uint32_t n_exec, epoch, max_holes;
  // boundary bricks (tickets in [boundary_begin, boundary_end)) wait for *ghost_ready == epoch and
  // count themselves into *boundary_done when their contributions are out (distributed vmult)
  uint32_t boundary_begin, boundary_end;
  const uint32_t *ghost_ready;
  uint32_t *boundary_done;
};

// table entry: [31:28] slot (0 = own range, 1..14 = group with a per-brick base, 15 = constrained:
// reads 0, never written), [27] first toucher (plain store), [26:0] offset
#define BULK_SLOT(e) ((e) >> 28)
#define BULK_FIRST(e) (((e) >> 27) & 1u)
#define BULK_OFF(e) ((e) & 0x7ffffffu)

template <int p, int b, typename Number, bool DOT>
__global__ void __launch_bounds__(BulkCfg<p, b, Number>::threads, BulkCfg<p, b, Number>::ctas_per_sm)
bulk_brick_kernel(const __grid_constant__ BulkKernelParams<p, Number> prm) {
  using Cfg = BulkCfg<p, b, Number>;
  constexpr int n = p + 1, L = Cfg::L, L2 = Cfg::L2, T = Cfg::threads;
  extern __shared__ __align__(128) unsigned char bulk_smem[];
  Number *P0 = reinterpret_cast<Number *>(bulk_smem);
  Number *P1 = P0 + Cfg::plane_elems; // also the staging buffer of the bulk load
  uint32_t *s_desc = reinterpret_cast<uint32_t *>(P1 + Cfg::plane_elems); // [2][48]
  uint64_t *mbar = reinterpret_cast<uint64_t *>(s_desc + 2 * kBulkDescWords);
  volatile uint32_t *s_ticket = reinterpret_cast<volatile uint32_t *>(mbar + 1); // [2]
  const int tid = threadIdx.x;
  const Number *__restrict__ src = prm.src;
  Number *__restrict__ dst = prm.dst;
  const bool active = tid < L2;
  const int la = tid % L, lb = tid / L;

  if (tid == 0) {
    mbar_init(mbar, 1);
    s_ticket[0] = atomicAdd(prm.ticket, 1u);
  }
  __syncthreads();
  uint32_t t = s_ticket[0];
  if (t < prm.n_exec && tid < kBulkDescWords) s_desc[tid] = __ldg(prm.desc + (size_t)t * kBulkDescWords + tid);
  __syncthreads();
  if (tid == 0 && t < prm.n_exec) {
    if (t >= prm.boundary_begin && t < prm.boundary_end && prm.ghost_ready != nullptr)
      while (ld_acquire(prm.ghost_ready) != prm.epoch) {
      }
    const uint32_t R = s_desc[2];
    if (R) {
      mbar_expect_tx(mbar, R * (uint32_t)sizeof(Number));
      bulk_load(P1, src + s_desc[1], R * (uint32_t)sizeof(Number), mbar);
    }
  }
  int cur = 0;
  uint32_t phase = 0;
  double dot = 0.0;

  while (t < prm.n_exec) {
    const uint32_t *D = s_desc + cur * kBulkDescWords;
    const uint32_t pattern = D[0], own_lo = D[1], R = D[2], n_deps = D[3];
    if (tid == 0) s_ticket[cur ^ 1] = atomicAdd(prm.ticket, 1u);

    // ---- read_dof_values: own range from the staging buffer, other nodes by indexed loads
    Number in[L];
    {
      uint32_t ex[L];
      const uint32_t *tx = prm.tx + (size_t)pattern * L * T + tid;
#pragma unroll
      for (int x = 0; x < L; ++x) ex[x] = active ? __ldg(tx + x * T) : (15u << 28);
#pragma unroll
      for (int x = 0; x < L; ++x) {
        const uint32_t sl = BULK_SLOT(ex[x]);
        in[x] = (sl != 0 && sl != 15) ? __ldg(src + D[4 + sl] + BULK_OFF(ex[x])) : Number(0);
      }
      if (R) mbar_wait(mbar, phase);
#pragma unroll
      for (int x = 0; x < L; ++x)
        if (BULK_SLOT(ex[x]) == 0) in[x] = P1[BULK_OFF(ex[x])];
    }
    if (R) phase ^= 1;
    __syncthreads(); // staging buffer consumed: P1 is free; s_ticket[cur ^ 1] is visible
    const uint32_t tn = s_ticket[cur ^ 1];
    if (tn < prm.n_exec && tid < kBulkDescWords)
      s_desc[(cur ^ 1) * kBulkDescWords + tid] = __ldg(prm.desc + (size_t)tn * kBulkDescWords + tid);

    // ---- x sweep: A = Mx u -> P0, B = Kx u -> P1; thread <-> (y, z)
    if (active) {
      Number *l0 = P0 + L * tid, *l1 = P1 + L * tid;
      Number cA = Number(0), cB = Number(0);
#pragma unroll
      for (int c = 0; c < b; ++c) {
        Number blk[n];
#pragma unroll
        for (int k = 0; k < n; ++k) blk[k] = in[c * p + k];
        EoHalf<Number, n> x;
        eo_split<Number, n>(blk, x);
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
      }
      l0[L - 1] = cA;
      l1[L - 1] = cB;
    }
    __syncthreads();

    // table entries of this thread's z line (the scatter targets): requested now, used after the
    // y and z sweeps
    uint32_t ez[L];
    {
      const uint32_t *tz = prm.tz + (size_t)pattern * L * T + tid;
#pragma unroll
      for (int z = 0; z < L; ++z) ez[z] = active ? __ldg(tz + z * T) : (15u << 28);
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

    // ---- z sweep: v = Kz' C + Mz D into registers; thread <-> (x, y)
    Number out[L];
    if (active) {
      const Number *l0 = P0 + tid, *l1 = P1 + tid;
      Number inC[n], inD[n], cV = Number(0);
      inC[0] = l0[0];
      inD[0] = l1[0];
#pragma unroll
      for (int c = 0; c < b; ++c) {
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
        for (int k = 0; k < p; ++k) out[c * p + k] = oV[k];
        cV = oV[p];
        inC[0] = inC[p];
        inD[0] = inD[p];
      }
      out[L - 1] = cV;
    }
    fence_async_smem(); // generic-proxy accesses of P1 before the async-proxy write of the next load
    __syncthreads();    // P0 and P1 are free

    // ---- the next brick's own range starts to arrive while this one is written out
    if (tid == 0 && tn < prm.n_exec) {
      const uint32_t *Dn = s_desc + (cur ^ 1) * kBulkDescWords;
      if (tn >= prm.boundary_begin && tn < prm.boundary_end && prm.ghost_ready != nullptr)
        while (ld_acquire(prm.ghost_ready) != prm.epoch) {
        }
      if (Dn[2]) {
        mbar_expect_tx(mbar, Dn[2] * (uint32_t)sizeof(Number));
        bulk_load(P1, src + Dn[1], Dn[2] * (uint32_t)sizeof(Number), mbar);
      }
    }

    // ---- distribute_local_to_global: own range through P0 (dof order) + one bulk store; nodes this
    // brick touches first are stored, the others wait for their first toucher's flag
    {
      const uint32_t *hl = prm.holes + (size_t)pattern * (1 + prm.max_holes);
      const uint32_t nh = __ldg(hl);
      for (uint32_t i = tid; i < nh; i += T) P0[__ldg(hl + 1 + i)] = Number(0);
    }
    bool stored_any = false;
#pragma unroll
    for (int z = 0; z < L; ++z) {
      const uint32_t e = ez[z], sl = BULK_SLOT(e);
      if (sl == 0) P0[BULK_OFF(e)] = out[z];
      else if (sl != 15 && BULK_FIRST(e)) {
        dst[D[4 + sl] + BULK_OFF(e)] = out[z];
        stored_any = true;
      }
    }
    if (stored_any) __threadfence();
    fence_async_smem();
    __syncthreads();
    if (tid == 0) {
      if (R) {
        bulk_store(dst + own_lo, P0, R * (uint32_t)sizeof(Number));
        bulk_store_wait();
      }
      __threadfence();
      st_release(prm.flags + t, prm.epoch);
    }
    if (DOT) {
      // src . (A src) of this brick: own range (dof order, coalesced) + the other nodes
      for (uint32_t i = tid; i < R; i += T) dot += double(__ldg(src + own_lo + i)) * double(P0[i]);
#pragma unroll
      for (int z = 0; z < L; ++z) {
        const uint32_t e = ez[z], sl = BULK_SLOT(e);
        if (sl != 0 && sl != 15) dot += double(__ldg(src + D[4 + sl] + BULK_OFF(e))) * double(out[z]);
      }
    }
    if ((uint32_t)tid < n_deps) {
      const uint32_t *f = prm.flags + D[20 + tid];
      while (ld_acquire(f) != prm.epoch) {
      }
    }
    __syncthreads(); // dependencies met; the bulk store has read P0
#pragma unroll
    for (int z = 0; z < L; ++z) {
      const uint32_t e = ez[z], sl = BULK_SLOT(e);
      if (sl != 0 && sl != 15 && !BULK_FIRST(e)) atomicAdd(dst + D[4 + sl] + BULK_OFF(e), out[z]);
    }
    if (prm.boundary_done != nullptr && t >= prm.boundary_begin && t < prm.boundary_end) {
      __threadfence();
      __syncthreads();
      if (tid == 0) atomicAdd(prm.boundary_done, 1u);
    }
    cur ^= 1;
    t = tn;
  }
  if (tid == 0 && t == prm.n_exec + gridDim.x - 1) *prm.ticket = 0u; // last ticket drawn: rearm
  if (DOT && prm.dot_accum != nullptr) {
    dot = block_sum(dot);
    if (tid == 0) atomicAdd(prm.dot_accum, dot);
  }
}

} // namespace b200mf
