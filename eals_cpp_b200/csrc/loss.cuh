// loss.cuh — K4: the eALS objective (MF_fastALS::loss, MF_fastALS.cpp:184-206).
//
//   L = reg (|U|^2 + |V|^2) + sum_u [ sum_{i in R_u} ( w_ui (r_ui - p_ui)^2 - Wi[i] p_ui^2 ) + u^T SV u ]
//
// with p_ui = <u, v_i> (predict, :208-221).  The per-nonzero part is one gather of v_i per nonzero
// and no recurrence; sum_u u^T SV u equals <SU, SV>_F with SU = U^T U, which the S cache already
// holds, so the reference's O(M K^2) loop (:200) collapses to K^2 multiply-adds.
// Partial sums go through per-CTA slots and are added in slot order: deterministic.
#pragma once

#include "common.cuh"

namespace eals {

struct LossSide {
  const int64_t* ptr;
  const int32_t* idx;
  const double* val;
  const double* X;   // row factors (U)
  const double* Y;   // column factors (V)
  const double* Wi;
  int row_base;
};

constexpr int kLossThreads = 256;

// TEAM = 32: one warp per row (short rows); TEAM = 256: one CTA per row (long rows).
// Eight lanes cooperate on one nonzero, each reading 16 B of every 128 B line of v_i.
template <int LD, int TEAM>
__global__ void __launch_bounds__(kLossThreads)
loss_rows_kernel(LossSide a, const int32_t* __restrict__ order, int first, int count,
                 double* __restrict__ partials) {
  __shared__ double red[kLossThreads / 32];
  const int tid = threadIdx.x, lane = lane_id(), warp = tid >> 5;
  constexpr int kTeams = kLossThreads / TEAM;
  const int team = tid / TEAM, tt = tid % TEAM;
  const int g = tt >> 3, gl = tt & 7;
  constexpr int kGroups = TEAM / 8;

  double acc = 0.0;
  for (int slot = blockIdx.x * kTeams + team; slot < count; slot += gridDim.x * kTeams) {
    const int row = order[first + slot];
    const int64_t p0 = a.ptr[row];
    const int n = (int)(a.ptr[row + 1] - p0);
    const double* xrow = a.X + (size_t)(a.row_base + row) * LD;
    double ux[LD / kFB], uy[LD / kFB];
#pragma unroll
    for (int c = 0; c < LD / kFB; c++) {
      const double2 d = ldg2(xrow + c * kFB + gl * 2);
      ux[c] = d.x;
      uy[c] = d.y;
    }
    for (int j0 = 0; j0 < n; j0 += kGroups) {
      const int j = j0 + g;
      double p = 0.0;
      int id = 0;
      if (j < n) {
        id = a.idx[p0 + j];
        const double* yrow = a.Y + (size_t)id * LD;
#pragma unroll
        for (int c = 0; c < LD / kFB; c++) {
          const double2 d = ldg2(yrow + c * kFB + gl * 2);
          p += ux[c] * d.x;
          p += uy[c] * d.y;
        }
      }
      p += __shfl_xor_sync(kFullMask, p, 1);
      p += __shfl_xor_sync(kFullMask, p, 2);
      p += __shfl_xor_sync(kFullMask, p, 4);
      if (j < n && gl == 0) {
        const double r = a.val ? a.val[p0 + j] : 1.0;
        const double d = r - p;
        acc += r * (d * d) - a.Wi[id] * (p * p);
      }
    }
  }
  acc = warp_sum(acc);
  if (lane == 0) red[warp] = acc;
  __syncthreads();
  if (tid == 0) {
    double s = 0.0;
#pragma unroll
    for (int w = 0; w < kLossThreads / 32; w++) s += red[w];
    partials[blockIdx.x] = s;
  }
}

// Per-nonzero part from the symmetric prediction cache (CSR order): p = <u, v> is already there, so
// the loss is one streaming pass over (index, cached prediction) plus the Wi gather.
__global__ void __launch_bounds__(kLossThreads)
loss_cached_kernel(const int32_t* __restrict__ idx, const double* __restrict__ val, const double* __restrict__ pcache,
                   const double* __restrict__ Wi, int64_t nnz, double* __restrict__ partials) {
  __shared__ double red[kLossThreads / 32];
  double acc = 0.0;
  for (int64_t q = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; q < nnz; q += (int64_t)gridDim.x * blockDim.x) {
    const double p = pcache[q];
    const double r = val ? val[q] : 1.0;
    const double d = r - p;
    acc += r * (d * d) - __ldg(Wi + idx[q]) * (p * p);
  }
  acc = warp_sum(acc);
  if (lane_id() == 0) red[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    double s = 0.0;
#pragma unroll
    for (int w = 0; w < kLossThreads / 32; w++) s += red[w];
    partials[blockIdx.x] = s;
  }
}

// Sum of squares of rows [r0, r1) (DenseMat::squaredSum, DenseMat.cpp:86-92); padding is zero.
__global__ void __launch_bounds__(256)
sumsq_kernel(const double* __restrict__ X, size_t begin, size_t end, double* __restrict__ partials) {
  __shared__ double red[8];
  double acc = 0.0;
  for (size_t t = begin + (size_t)blockIdx.x * blockDim.x + threadIdx.x; t < end;
       t += (size_t)gridDim.x * blockDim.x) {
    const double x = X[t];
    acc += x * x;
  }
  acc = warp_sum(acc);
  if (lane_id() == 0) red[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    double s = 0.0;
    for (int w = 0; w < 8; w++) s += red[w];
    partials[blockIdx.x] = s;
  }
}

// One CTA: terms[slot] = sum of n partials in index order (fixed tree), optionally += .
__global__ void __launch_bounds__(256)
sum_partials_kernel(const double* __restrict__ partials, int n, double* __restrict__ out, int accumulate) {
  __shared__ double red[8];
  double acc = 0.0;
  for (int t = threadIdx.x; t < n; t += blockDim.x) acc += partials[t];
  acc = warp_sum(acc);
  if (lane_id() == 0) red[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    double s = 0.0;
    for (int w = 0; w < 8; w++) s += red[w];
    *out = accumulate ? *out + s : s;
  }
}

// <SU, SV>_F over the K x K live part.
__global__ void __launch_bounds__(256)
frob_inner_kernel(const double* __restrict__ A, const double* __restrict__ B, int K, int LD,
                  double* __restrict__ out) {
  __shared__ double red[8];
  double acc = 0.0;
  for (int t = threadIdx.x; t < K * K; t += blockDim.x) {
    const int f = t / K, k = t % K;
    acc += A[(size_t)f * LD + k] * B[(size_t)f * LD + k];
  }
  acc = warp_sum(acc);
  if (lane_id() == 0) red[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    double s = 0.0;
    for (int w = 0; w < 8; w++) s += red[w];
    *out = s;
  }
}

}  // namespace eals
