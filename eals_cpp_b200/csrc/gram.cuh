// gram.cuh — K2: the S caches as a weighted Gram  S = X^T diag(w) X  (fp64).
//
// Replaces MF_fastALS::initS (MF_fastALS.cpp:583-595) and, by recomputation after each sweep, the
// per-row rank-1 patches update_user_SU / update_item_SV (:324-335, :409-422): SU = U^T U (w = 1),
// SV = V^T diag(Wi) V.  Recomputing instead of patching changes the loss by <= 2e-15 relative
//  and removes the old-factor clone the reference keeps per iteration (:119,137).
//
// Two stages so the result is run-to-run deterministic: every CTA reduces a contiguous slab of
// rows into its own TB x TB partial (register-tiled, operands staged in shared memory), then a
// second kernel adds the partials in slab order and mirrors the off-diagonal blocks.
#pragma once

#include "common.cuh"

namespace eals {

constexpr int kGramThreads = 256;
constexpr int kGramChunk = 16;  // rows staged per step

template <int LD>
struct GramCfg {
  static constexpr int TB = LD < 128 ? LD : 128;  // output block edge handled by one CTA
  static constexpr int TT = TB / 16;              // per-thread tile edge (threads form a 16 x 16 grid)
  static constexpr int NB = LD / TB;              // output blocks per dimension (2 only for LD = 256)
  static constexpr int NPAIR = NB * (NB + 1) / 2; // lower-triangular block pairs
};

// partials layout: [pair][slab][TB*TB]
template <int LD>
__global__ void __launch_bounds__(kGramThreads)
gram_partial_kernel(const double* __restrict__ X, const double* __restrict__ w, int r0, int r1,
                    double* __restrict__ partials) {
  using C = GramCfg<LD>;
  __shared__ double As[kGramChunk][C::TB];
  __shared__ double Bs[kGramChunk][C::TB];
  const int tid = threadIdx.x, ty = tid >> 4, tx = tid & 15;
  int bi = 0, bj = 0;
  if (C::NB == 2) {  // pair 0 -> (0,0), 1 -> (1,0), 2 -> (1,1)
    bi = blockIdx.y >= 1;
    bj = blockIdx.y == 2;
  }
  const int rows = r1 - r0;
  const int per = (rows + gridDim.x - 1) / gridDim.x;
  const int s0 = r0 + blockIdx.x * per;
  const int s1 = min(s0 + per, r1);

  double acc[C::TT][C::TT];
#pragma unroll
  for (int i = 0; i < C::TT; i++)
#pragma unroll
    for (int j = 0; j < C::TT; j++) acc[i][j] = 0.0;

  for (int c0 = s0; c0 < s1; c0 += kGramChunk) {
    for (int t = tid; t < kGramChunk * C::TB; t += kGramThreads) {
      const int r = t / C::TB, c = t % C::TB;
      const int row = c0 + r;
      double a = 0.0, b = 0.0;
      if (row < s1) {
        a = X[(size_t)row * LD + bi * C::TB + c];
        b = X[(size_t)row * LD + bj * C::TB + c];
        if (w) b *= w[row];
      }
      As[r][c] = a;
      Bs[r][c] = b;
    }
    __syncthreads();
#pragma unroll 4
    for (int r = 0; r < kGramChunk; r++) {
      double av[C::TT], bv[C::TT];
#pragma unroll
      for (int i = 0; i < C::TT; i++) av[i] = As[r][ty + 16 * i];
#pragma unroll
      for (int j = 0; j < C::TT; j++) bv[j] = Bs[r][tx + 16 * j];
#pragma unroll
      for (int i = 0; i < C::TT; i++)
#pragma unroll
        for (int j = 0; j < C::TT; j++) acc[i][j] += av[i] * bv[j];
    }
    __syncthreads();
  }
  double* out = partials + ((size_t)blockIdx.y * gridDim.x + blockIdx.x) * C::TB * C::TB;
#pragma unroll
  for (int i = 0; i < C::TT; i++)
#pragma unroll
    for (int j = 0; j < C::TT; j++) out[(ty + 16 * i) * C::TB + tx + 16 * j] = acc[i][j];
}

// S[f][k] = sum over slabs (ascending) of the partials; rows f >= K are not written.
template <int LD>
__global__ void gram_reduce_kernel(const double* __restrict__ partials, int nslabs, int K,
                                   double* __restrict__ S) {
  using C = GramCfg<LD>;
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= C::NPAIR * C::TB * C::TB) return;
  const int pair = t / (C::TB * C::TB), e = t % (C::TB * C::TB);
  const int i = e / C::TB, j = e % C::TB;
  int bi = 0, bj = 0;
  if (C::NB == 2) {
    bi = pair >= 1;
    bj = pair == 2;
  }
  const double* p = partials + (size_t)pair * nslabs * C::TB * C::TB + e;
  double s = 0.0;
  for (int sl = 0; sl < nslabs; sl++) s += p[(size_t)sl * C::TB * C::TB];
  const int f = bi * C::TB + i, k = bj * C::TB + j;
  if (f < K) S[(size_t)f * LD + k] = s;
  if (bi != bj && k < K) S[(size_t)k * LD + f] = s;
}

// Rank-1 patch used by the single-row API (update_user_SU / update_item_SV,
// MF_fastALS.cpp:324-335, 409-422):  S += scale * (new new^T - old old^T).
__global__ void gram_patch_kernel(double* __restrict__ S, const double* __restrict__ oldv,
                                  const double* __restrict__ newv, double scale, int K, int LD) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= K * K) return;
  const int f = t / K, k = t % K;
  S[(size_t)f * LD + k] = S[(size_t)f * LD + k] - oldv[f] * oldv[k] * scale + newv[f] * newv[k] * scale;
}

}  // namespace eals
