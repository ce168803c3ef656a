// common.cuh — shared device/host helpers for libeals_b200 (sm_100a only).
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

#include <cstdio>
#include <string>

namespace eals {

constexpr unsigned kFullMask = 0xffffffffu;
constexpr int kFB = 16;        // factor-block width in doubles = one 128-byte line of a factor row
constexpr int kTilePad = 17;   // smem row stride of a staged factor block (odd => conflict-free column reads)

// Leading dimension of factor rows: next power of two >= K, at least one 128 B line.
inline int leading_dim_for(int K) {
  int ld = kFB;
  while (ld < K) ld <<= 1;
  return ld;
}

__device__ __forceinline__ int lane_id() { return threadIdx.x & 31; }

// Sum a and b over the warp with 6 fp64 shuffles instead of 10: after round one, even lanes carry
// the a-partials and odd lanes the b-partials; xor-rounds 2..16 stay within a parity class; a last
// exchange hands every lane both totals.  The order of additions is fixed => run-to-run identical.
__device__ __forceinline__ void warp_sum_pair(double& a, double& b) {
  const bool odd = lane_id() & 1;
  double keep = odd ? b : a;
  const double send = odd ? a : b;
  keep += __shfl_xor_sync(kFullMask, send, 1);
#pragma unroll
  for (int o = 2; o < 32; o <<= 1) keep += __shfl_xor_sync(kFullMask, keep, o);
  const double other = __shfl_xor_sync(kFullMask, keep, 1);
  a = odd ? other : keep;
  b = odd ? keep : other;
}

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(kFullMask, v, o);
  return v;
}

// 16-byte read-only load of two doubles.
__device__ __forceinline__ double2 ldg2(const double* p) {
  return __ldg(reinterpret_cast<const double2*>(p));
}

}  // namespace eals
