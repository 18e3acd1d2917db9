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
//    (cp.async.bulk, TMA engine, mbarrier completion) into shared memory in dof order; a fill pass
//    permutes it to lattice order through a 2-byte PATTERN table shared by all bricks with the
//    same relative numbering (bulk_setup.cpp): linear reads, scattered writes into the odd-stride
//    lattice (few bank conflicts).  Only the lattice nodes other bricks own (the low faces: 17 %
//    for Q4) are gathered with indexed loads, from a per-brick base + the pattern's offset, spread
//    evenly over the threads: no per-node index map is streamed.
//  * distribute_local_to_global: the z sweep leaves the results in lattice order, a second pass
//    permutes the own range back into dof order and it leaves with ONE bulk async store.  No
//    memset of dst, no atomics on these dofs.
//  * conflicts (portable_matrix_free.templates.h:1060-1185 offers colouring or atomics): the
//    brick that touches a dof first STORES it, every later toucher waits for that brick's flag and
//    adds with RED.  Bricks are handed out through a ticket counter in execution order, so a
//    brick only ever waits for bricks that are already running: no deadlock, one launch, and
//    the boundary bricks of a partitioned mesh can be placed between two halves of the interior.
//  * persistent CTAs, software-pipelined: the next ticket/descriptor is fetched while the current
//    brick is swept, its bulk load is issued as soon as the staging buffer is free, and the tail of
//    a brick -- wait for the bulk store to land, raise the flag, wait for the first touchers, issue
//    the REDs from a small shared-memory stash -- runs inside the sweeps of the NEXT brick, so no
//    warp ever sits at a barrier waiting for a store or a flag.
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
// the shared-memory source of every committed bulk store has been read (the buffer may be reused)
__device__ __forceinline__ void bulk_store_wait_read() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ uint32_t ld_acquire(const uint32_t *p) {
  uint32_t v;
  asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_release(uint32_t *p, uint32_t v) {
  asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}

// streaming load of a value used once (the gathered nodes of other bricks): do not allocate it in
// L1, whose capacity is better spent on the pattern tables every brick re-reads
__device__ __forceinline__ double ld_stream(const double *p) {
  double v;
  asm volatile("ld.global.nc.L1::no_allocate.f64 %0, [%1];" : "=d"(v) : "l"(p));
  return v;
}
__device__ __forceinline__ float ld_stream(const float *p) {
  float v;
  asm volatile("ld.global.nc.L1::no_allocate.f32 %0, [%1];" : "=f"(v) : "l"(p));
  return v;
}

template <int p, int b, typename Number>
struct BulkCfg {
  static constexpr int n = p + 1;
  static constexpr int L = b * p + 1;
  static constexpr int L2 = L * L, L3 = L * L * L;
  static constexpr int threads = ((L2 + 31) / 32) * 32;
  static constexpr int plane_elems = ((L3 + 15) / 16) * 16; // 128-byte multiple for FP64
  // stash of the later-toucher contributions of one brick (its low faces), double buffered
  static constexpr int stash = ((3 * L2 + 63) / 64) * 64;
  static constexpr size_t smem_bytes =
      2 * sizeof(Number) * plane_elems + 2 * stash * sizeof(Number) + 3 * kBulkDescWords * 4 + 64;
  static constexpr int by_smem = (int)((227 * 1024) / (smem_bytes + 1024));
  static constexpr int ctas_per_sm = by_smem < 1 ? 1 : (by_smem > 2 ? 2 : by_smem);
};

template <int p, typename Number>
struct BulkKernelParams {
  BrickMatrices<Number, p + 1> mat;
  const uint32_t *desc;      // [n_exec][48]: pattern, own_lo, own_count, n_deps, base[16], dep[26], n_later
  const uint32_t *other;     // [pattern][PS] entries of the nodes outside the own range (later touchers first)
  const uint32_t *phdr;      // [pattern][4]  n_other, n_later, n_first
  const uint16_t *own_pos;   // [pattern][PS] lattice position of the i-th dof of the own range, 0xffff = hole
  const uint16_t *other_pos; // [pattern][PS] lattice position of other[k]
  uint32_t *flags;           // [n_exec] epoch of the launch in which the brick stored its dofs
  uint32_t *ticket;
  const Number *src;
  Number *dst;
  double *dot_accum;
  uint32_t n_exec, epoch;
  // boundary bricks (tickets in [boundary_begin, boundary_end)) wait for *ghost_ready == epoch and
  // count themselves into *boundary_done when their contributions are out (distributed vmult)
  uint32_t boundary_begin, boundary_end;
  const uint32_t *ghost_ready;
  uint32_t *boundary_done;
  uint32_t debug; // experiments only (B200MF_BULK_DEBUG): 2 no flag polling, 4 no REDs, 16 no bulk store
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
  constexpr uint32_t CAP = Cfg::stash, PS = Cfg::plane_elems;
  extern __shared__ __align__(128) unsigned char bulk_smem[];
  Number *P0 = reinterpret_cast<Number *>(bulk_smem);
  Number *P1 = P0 + PS; // also the staging buffer of the bulk load
  Number *H = P1 + PS;  // [2][CAP]
  uint32_t *s_desc = reinterpret_cast<uint32_t *>(H + 2 * CAP); // [3][48]
  uint64_t *mbar = reinterpret_cast<uint64_t *>(s_desc + 3 * kBulkDescWords);
  volatile uint32_t *s_ticket = reinterpret_cast<volatile uint32_t *>(mbar + 1); // [2]
  const int tid = threadIdx.x;
  const Number *__restrict__ src = prm.src;
  Number *__restrict__ dst = prm.dst;
  const bool active = tid < L2;
  const int la = tid % L, lb = tid / L;

  // thread 0: the brick's own range starts to arrive in P1 (every brick completes one phase of the
  // mbarrier, also those without an own range, so that "phase done" also means "P0/P1 are free")
  auto issue_load = [&](uint32_t tk, const uint32_t *Dk) {
    if (tk >= prm.boundary_begin && tk < prm.boundary_end && prm.ghost_ready != nullptr)
      while (ld_acquire(prm.ghost_ready) != prm.epoch) {
      }
    const uint32_t bytes = Dk[2] * (uint32_t)sizeof(Number);
    mbar_expect_tx(mbar, bytes);
    if (bytes) bulk_load(P1, src + Dk[1], bytes, mbar);
  };

  if (tid == 0) {
    mbar_init(mbar, 1);
    s_ticket[0] = atomicAdd(prm.ticket, 1u);
  }
  __syncthreads();
  uint32_t t = s_ticket[0];
  if (t < prm.n_exec && tid < kBulkDescWords) s_desc[tid] = __ldg(prm.desc + (size_t)t * kBulkDescWords + tid);
  __syncthreads();
  if (tid == 0 && t < prm.n_exec) issue_load(t, s_desc);
  uint32_t it = 0, phase = 0;
  double dot = 0.0;
  // the deferred tail of the previous brick of this CTA: raise its flag once its bulk store has
  // landed (tail_flag), then, once its first touchers have raised theirs, RED its later-toucher
  // contributions from the stash (tail_red)
  bool tail_flag = false, tail_red = false;
  uint32_t t_prev = 0;

  auto raise_flag = [&]() {
    if (tid == 0) {
      bulk_store_wait();
      __threadfence();
      st_release(prm.flags + t_prev, prm.epoch);
    }
  };
  auto deps_ready = [&](const uint32_t *Dp) -> bool { // one non-blocking look at the flags
    if ((uint32_t)tid < Dp[3] && !(prm.debug & 2)) return ld_acquire(prm.flags + Dp[20 + tid]) == prm.epoch;
    return true;
  };
  auto red_stash = [&](const uint32_t *Dp, const Number *Hp) {
    const uint32_t pat = Dp[0], nl = __ldg(prm.phdr + pat * 4 + 1);
    const uint32_t *oe = prm.other + (size_t)pat * PS;
    if (!(prm.debug & 4))
      for (uint32_t k = tid; k < nl; k += T) {
        const uint32_t e = __ldg(oe + k);
        atomicAdd(dst + Dp[4 + BULK_SLOT(e)] + BULK_OFF(e), Hp[k]);
      }
    if (prm.boundary_done != nullptr && t_prev >= prm.boundary_begin && t_prev < prm.boundary_end) {
      __threadfence();
      __syncthreads();
      if (tid == 0) atomicAdd(prm.boundary_done, 1u);
    }
  };

  while (t < prm.n_exec) {
    const uint32_t *D = s_desc + (it % 3) * kBulkDescWords;
    const uint32_t *Dprev = s_desc + ((it + 2) % 3) * kBulkDescWords;
    const uint32_t pattern = D[0], own_lo = D[1], R = D[2];
    const uint32_t n_other = __ldg(prm.phdr + pattern * 4), n_later = __ldg(prm.phdr + pattern * 4 + 1);
    const bool deferred = n_later <= CAP; // the later-toucher nodes fit the stash
    const uint32_t *oe = prm.other + (size_t)pattern * PS;
    const uint16_t *op = prm.other_pos + (size_t)pattern * PS;
    const uint16_t *wp = prm.own_pos + (size_t)pattern * PS;
    Number *Hcur = H + (it & 1) * CAP;
    const Number *Hprev = H + ((it + 1) & 1) * CAP;
    if (tid == 0) s_ticket[(it + 1) & 1] = atomicAdd(prm.ticket, 1u);

    // ---- read_dof_values, fill pass: nodes outside the own range by indexed loads (issued first),
    // own range from the staging buffer in dof order -> lattice order in P0
    {
      constexpr int CH = 4;
      for (uint32_t k0 = 0; k0 < n_other; k0 += CH * T) {
        Number v[CH];
        uint32_t pos[CH];
#pragma unroll
        for (int j = 0; j < CH; ++j) {
          const uint32_t k = k0 + j * T + tid;
          v[j] = Number(0);
          pos[j] = 0xffffu;
          if (k < n_other) {
            const uint32_t e = __ldg(oe + k), sl = BULK_SLOT(e);
            pos[j] = __ldg(op + k);
            if (sl != 15) v[j] = ld_stream(src + D[4 + sl] + BULK_OFF(e));
          }
        }
        if (k0 == 0) mbar_wait(mbar, phase); // P0 is free (and the own range has arrived in P1)
#pragma unroll
        for (int j = 0; j < CH; ++j)
          if (pos[j] != 0xffffu) P0[pos[j]] = v[j];
      }
      if (n_other == 0) mbar_wait(mbar, phase);
      phase ^= 1;
      constexpr int UN = 4;
      for (uint32_t i0 = 0; i0 < R; i0 += UN * T) {
        uint32_t pos[UN];
        Number v[UN];
#pragma unroll
        for (int j = 0; j < UN; ++j) {
          const uint32_t i = i0 + j * T + tid;
          pos[j] = i < R ? (uint32_t)__ldg(wp + i) : 0xffffu;
          v[j] = i < R ? P1[i] : Number(0);
        }
#pragma unroll
        for (int j = 0; j < UN; ++j)
          if (pos[j] != 0xffffu) P0[pos[j]] = v[j];
      }
    }
    __syncthreads(); // A: lattice filled, staging buffer consumed; s_ticket of the next brick visible
    const uint32_t tn = s_ticket[(it + 1) & 1];
    if (tn < prm.n_exec && tid < kBulkDescWords)
      s_desc[((it + 1) % 3) * kBulkDescWords + tid] = __ldg(prm.desc + (size_t)tn * kBulkDescWords + tid);
    if (tail_flag) raise_flag();
    tail_flag = false;

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
    __syncthreads(); // B
    bool ready = tail_red ? deps_ready(Dprev) : true;

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
    // C (with a vote: have the first touchers of the previous brick raised their flags?)
    if (__syncthreads_and(ready ? 1 : 0)) {
      if (tail_red) red_stash(Dprev, Hprev);
      tail_red = false;
    }
    if (tail_red) ready = deps_ready(Dprev);

    // ---- z sweep: v = Kz' C + Mz D -> P1 (in place); thread <-> (x, y)
    if (active) {
      const Number *l0 = P0 + tid;
      Number *l1 = P1 + tid;
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
        for (int k = 0; k < p; ++k) l1[(c * p + k) * L2] = oV[k];
        cV = oV[p];
        inC[0] = inC[p];
        inD[0] = inD[p];
      }
      l1[(L - 1) * L2] = cV;
    }
    // D
    if (__syncthreads_and((!tail_red || ready) ? 1 : 0)) {
      if (tail_red) red_stash(Dprev, Hprev);
      tail_red = false;
    }

    // ---- distribute_local_to_global: own range P1 (lattice order) -> P0 (dof order) -> one bulk
    // store; nodes this brick touches first are stored, the others go to the stash
    {
      constexpr int UN = 4;
      for (uint32_t i0 = 0; i0 < R; i0 += UN * T) {
        uint32_t pos[UN];
#pragma unroll
        for (int j = 0; j < UN; ++j) {
          const uint32_t i = i0 + j * T + tid;
          pos[j] = i < R ? (uint32_t)__ldg(wp + i) : 0xfffeu;
        }
#pragma unroll
        for (int j = 0; j < UN; ++j) {
          const uint32_t i = i0 + j * T + tid;
          if (pos[j] != 0xfffeu) {
            const Number v = pos[j] != 0xffffu ? P1[pos[j]] : Number(0);
            P0[i] = v;
            if (DOT) dot += double(__ldg(src + own_lo + i)) * double(v);
          }
        }
      }
      bool stored_any = false;
      for (uint32_t k = tid; k < n_other; k += T) {
        const uint32_t e = __ldg(oe + k), sl = BULK_SLOT(e);
        if (sl == 15) break; // constrained nodes come last
        const Number v = P1[__ldg(op + k)];
        const uint32_t gi = D[4 + sl] + BULK_OFF(e);
        if (DOT) dot += double(__ldg(src + gi)) * double(v);
        if (k >= n_later) {
          dst[gi] = v;
          stored_any = true;
        } else if (deferred) {
          Hcur[k] = v;
        }
      }
      if (stored_any) __threadfence();
    }
    fence_async_smem();
    __syncthreads(); // E: the own range is staged in P0
    if (deferred && tid == 0) {
      if (R && !(prm.debug & 16)) bulk_store(dst + own_lo, P0, R * (uint32_t)sizeof(Number));
      bulk_store_wait_read(); // P0 may be overwritten: the next brick's arrival says so to everybody
      if (tn < prm.n_exec) issue_load(tn, s_desc + ((it + 1) % 3) * kBulkDescWords);
    }
    if (tail_red) { // the first touchers of the previous brick are late: wait for them now
      if ((uint32_t)tid < Dprev[3])
        while (ld_acquire(prm.flags + Dprev[20 + tid]) != prm.epoch) {
        }
      __syncthreads();
      red_stash(Dprev, Hprev);
      tail_red = false;
    }
    t_prev = t;
    if (!deferred) {
      // irregular brick (more later-toucher nodes than the stash holds): finish it now, with the
      // results still in P1, before the next brick's own range may arrive there
      if (tid == 0 && R) bulk_store(dst + own_lo, P0, R * (uint32_t)sizeof(Number));
      raise_flag();
      if ((uint32_t)tid < D[3])
        while (ld_acquire(prm.flags + D[20 + tid]) != prm.epoch) {
        }
      __syncthreads();
      for (uint32_t k = tid; k < n_later; k += T) {
        const uint32_t e = __ldg(oe + k);
        atomicAdd(dst + D[4 + BULK_SLOT(e)] + BULK_OFF(e), P1[__ldg(op + k)]);
      }
      const bool is_boundary = prm.boundary_done != nullptr && t >= prm.boundary_begin && t < prm.boundary_end;
      if (is_boundary) __threadfence();
      __syncthreads();
      if (tid == 0) {
        if (is_boundary) atomicAdd(prm.boundary_done, 1u);
        if (tn < prm.n_exec) issue_load(tn, s_desc + ((it + 1) % 3) * kBulkDescWords);
      }
    }
    tail_flag = deferred;
    tail_red = deferred && n_later > 0;
    ++it;
    t = tn;
  }
  if (tail_flag) raise_flag();
  if (tail_red) {
    const uint32_t *Dprev = s_desc + ((it + 2) % 3) * kBulkDescWords;
    if ((uint32_t)tid < Dprev[3])
      while (ld_acquire(prm.flags + Dprev[20 + tid]) != prm.epoch) {
      }
    __syncthreads();
    red_stash(Dprev, H + ((it + 1) & 1) * CAP);
  }
  if (tid == 0 && t == prm.n_exec + gridDim.x - 1) *prm.ticket = 0u; // last ticket drawn: rearm
  if (DOT && prm.dot_accum != nullptr) {
    dot = block_sum(dot);
    if (tid == 0) atomicAdd(prm.dot_accum, dot);
  }
}

} // namespace b200mf
