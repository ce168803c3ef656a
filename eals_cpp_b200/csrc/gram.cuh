// gram.cuh — K2: the S caches as a weighted Gram  S = X^T diag(w) X  (fp64).
//
// Replaces MF_fastALS::initS (MF_fastALS.cpp:583-595) and, by recomputation after each sweep, the
// per-row rank-1 patches update_user_SU / update_item_SV (:324-335, :409-422): SU = U^T U (w = 1),
// SV = V^T diag(Wi) V.  Recomputing instead of patching changes the loss by <= 2e-15 relative
//  and removes the old-factor clone the reference keeps per iteration (:119,137).
//
// Two stages so the result is run-to-run deterministic: every CTA reduces a contiguous slab of
// rows into its own TB x TB partial on the fp64 tensor cores (operands staged in shared memory), then
// a second kernel adds the partials in slab order and mirrors the off-diagonal blocks.
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
//
// Tensor-core version (fp64 mma.sync.m8n8k4): the CTA's 8 warps share the 8 x 8 DMMA tiles of the TB x TB
// output block; rows of X are staged 16 at a time through shared memory (A = X^T chunk, B = diag(w) X chunk).
// TRI (diagonal block pairs, i.e. everything for K <= 128): S is symmetric, so only the tiles on and below
// the tile diagonal are computed — 136 instead of 256 for TB = 128, 17 per warp — and the reduction mirrors
// them, exactly like the reference's loop over k <= f (MF_fastALS.cpp:586-593).  Round 1 computed the full
// block: 2x the DMMA work of a kernel that runs at the DMMA rate.
template <int LD, bool TRI>
__global__ void __launch_bounds__(kGramThreads)
gram_partial_kernel(const double* __restrict__ X, const double* __restrict__ w, int r0, int r1,
                    double* __restrict__ partials, int pair_base, int pair_step) {
  using C = GramCfg<LD>;
  constexpr int TB = C::TB;
  constexpr int NT = TB / 8;                 // 8 x 8 tiles per dimension of the output block
  constexpr int NTILES = TRI ? NT * (NT + 1) / 2 : NT * NT;
  constexpr int TPW = (NTILES + 7) / 8;      // tiles per warp (TB = 128: 17 / 32, 64: 5 / 8, 32: 2, 16: 1)
  // staged rows: As[r][c] = X[row][bi*TB + c], Bs[r][c] = w[row] * X[row][bj*TB + c]; stride TB + 4
  // doubles keeps the 4-row x 8-column fragment reads on distinct banks
  constexpr int ST = TB + 4;
  __shared__ double As[kGramChunk][ST];
  __shared__ double Bs[kGramChunk][ST];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int pair = pair_base + blockIdx.y * pair_step;
  int bi = 0, bj = 0;
  if (C::NB == 2) {  // pair 0 -> (0,0), 1 -> (1,0), 2 -> (1,1)
    bi = pair >= 1;
    bj = pair == 2;
  }
  const int rows = r1 - r0;
  const int per = (rows + gridDim.x - 1) / gridDim.x;
  const int s0 = r0 + blockIdx.x * per;
  const int s1 = min(s0 + per, r1);

  // this warp's tiles: tile index -> (ti, tj); triangular enumeration row by row: idx = ti (ti + 1) / 2 + tj
  int ti_[TPW], tj_[TPW];
#pragma unroll
  for (int t = 0; t < TPW; t++) {
    const int tile = warp * TPW + t;
    int ti = 0, tj = 0;
    if (TRI) {
      while ((ti + 1) * (ti + 2) / 2 <= tile) ti++;
      tj = tile - ti * (ti + 1) / 2;
    } else {
      ti = tile / NT; tj = tile % NT;
    }
    ti_[t] = ti; tj_[t] = tj;
  }

  double acc[TPW][2];
#pragma unroll
  for (int t = 0; t < TPW; t++) acc[t][0] = acc[t][1] = 0.0;

  for (int c0 = s0; c0 < s1; c0 += kGramChunk) {
    for (int t = tid; t < kGramChunk * TB; t += kGramThreads) {
      const int r = t / TB, c = t % TB;
      const int row = c0 + r;
      double a = 0.0, b = 0.0;
      if (row < s1) {
        a = X[(size_t)row * LD + bi * TB + c];
        b = X[(size_t)row * LD + bj * TB + c];
        if (w) b *= w[row];
      }
      As[r][c] = a;
      Bs[r][c] = b;
    }
    __syncthreads();
#pragma unroll
    for (int k0 = 0; k0 < kGramChunk; k0 += 4) {
#pragma unroll
      for (int t = 0; t < TPW; t++) {
        if (warp * TPW + t < NTILES) {
          const double a = As[k0 + (lane & 3)][ti_[t] * 8 + (lane >> 2)];   // A[row = f][col = k]
          const double b = Bs[k0 + (lane & 3)][tj_[t] * 8 + (lane >> 2)];   // B[row = k][col = f']
          asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                       : "+d"(acc[t][0]), "+d"(acc[t][1])
                       : "d"(a), "d"(b));
        }
      }
    }
    __syncthreads();
  }
  double* out = partials + ((size_t)pair * gridDim.x + blockIdx.x) * TB * TB;
#pragma unroll
  for (int t = 0; t < TPW; t++) {
    if (warp * TPW + t < NTILES) {
      const int rr = ti_[t] * 8 + (lane >> 2), cc = tj_[t] * 8 + 2 * (lane & 3);
      out[rr * TB + cc] = acc[t][0];
      out[rr * TB + cc + 1] = acc[t][1];
    }
  }
}

// S[f][k] = sum over slabs (ascending) of the partials; rows f >= K are not written.  Diagonal block pairs
// hold the lower triangle only (elements with column <= row): those are summed and mirrored, so S comes out
// exactly symmetric, as the reference's does.
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
  if (bi == bj && j > i) return;             // upper triangle of a diagonal block: written by its mirror image
  const double* p = partials + (size_t)pair * nslabs * C::TB * C::TB + e;
  double s = 0.0;
  for (int sl = 0; sl < nslabs; sl++) s += p[(size_t)sl * C::TB * C::TB];
  const int f = bi * C::TB + i, k = bj * C::TB + j;
  if (f < K) S[(size_t)f * LD + k] = s;
  if (f != k && k < K) S[(size_t)k * LD + f] = s;
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
