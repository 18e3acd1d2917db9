// Vector kernels shared by vector_ops.cu and solver.cu: block reductions and the fused
// update kernels of CG / Chebyshev.  They replace the one-pass-per-operation Kokkos
// dispatches of LinearAlgebra::distributed::Vector<Number, MemorySpace::Default>
// (lac/vector_operations_internal.h:2140-2660) and the three Chebyshev kernels of
// lac/precondition.h:3358-3430.
#pragma once
#include "internal.h"

namespace b200mf {

constexpr int kVecThreads = 256;

inline unsigned vec_grid(uint64_t n, int per_thread = 4) {
  // enough CTAs to fill 148 SMs several times over, capped so that tiny vectors stay cheap
  uint64_t blocks = (n + (uint64_t)kVecThreads * per_thread - 1) / ((uint64_t)kVecThreads * per_thread);
  const uint64_t cap = 148ull * 16;
  if (blocks > cap) blocks = cap;
  if (blocks == 0) blocks = 1;
  return (unsigned)blocks;
}

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// Sum `v` over the CTA; the result is valid in thread 0.
__device__ __forceinline__ double block_sum(double v) {
  __shared__ double part[32];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  v = warp_sum(v);
  __syncthreads(); // protect `part` against a previous use
  if (lane == 0) part[warp] = v;
  __syncthreads();
  if (warp == 0) {
    const int nw = (blockDim.x + 31) >> 5;
    v = lane < nw ? part[lane] : 0.0;
    v = warp_sum(v);
  }
  return v;
}

} // namespace b200mf
