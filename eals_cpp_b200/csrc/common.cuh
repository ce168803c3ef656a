// common.cuh — shared device/host helpers for libeals_b200 (sm_100a only).
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

#include <cstdio>
#include <string>

namespace eals {

constexpr unsigned kFullMask = 0xffffffffu;
constexpr int kFB = 16;        // factor-block width in doubles = one 128-byte line of a factor row

// Leading dimension of factor rows: next power of two >= K, at least one 128 B line.
inline int leading_dim_for(int K) {
  int ld = kFB;
  while (ld < K) ld <<= 1;
  return ld;
}

__device__ __forceinline__ int lane_id() { return threadIdx.x & 31; }

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
