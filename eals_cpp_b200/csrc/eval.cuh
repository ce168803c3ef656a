// eval.cuh — K3: leave-one-out evaluation, device part (MF_fastALS::evaluate_for_user,
// MF_fastALS.cpp:620-662, driven over all users by evaluate_model, main.cpp:37-65).
//
// Per user u with held-out item g the reference scores every item, counts the items whose score
// is STRICTLY larger than score(u,g) (:628-634) and gives (0,0,0) as soon as that count exceeds
// topK.  Whether a user survives therefore depends on exact fp64 comparisons, and one flipped
// user moves HR by 1/M.  This first, exact version reproduces the reference's scores BIT FOR BIT:
// predict() is a sequential-k sum of separately rounded products (:216-218; the reference build
// has no FMA contraction), so the kernels use __dmul_rn/__dadd_rn in k order — never an FMA.
//
// Outputs: count_larger[u] (exact), and for the surviving users a sparse stream of
// (slot, item, (int)score) triples for the items whose truncated score is non-zero — all the host
// needs to replay the reference's int-truncating partial_sort_copy (:643-651).
#pragma once

#include "common.cuh"

namespace eals {

__device__ __forceinline__ double seq_dot(const double* __restrict__ a, const double* __restrict__ b, int K) {
  double acc = 0.0;
  for (int k = 0; k < K; k++) acc = __dadd_rn(acc, __dmul_rn(a[k], b[k]));
  return acc;
}

// 32 reference-order dot products per warp: lane l owns the pair of rows (a, b) it passes in (nullptr: none)
// and gets back the sequential-k sum of separately rounded products, exactly seq_dot — but the rows are read
// COOPERATIVELY, a 128-byte line of a row by 8 lanes, 16 factors of all 32 pairs at a time through shared
// memory (one thread walking its own 1 KB rows touches 32 different lines per load instruction: the gt scores
// of 10M users took 10 ms, the exact re-score of 27M candidate pairs 33 ms that way).  `sm` = this warp's
// scratch of 2 * 32 * 17 doubles.
__device__ __forceinline__ double seq_dot_warp32(const double* a, const double* b, int K, double* sm) {
  const int lane = threadIdx.x & 31;
  double* sa = sm;
  double* sb = sm + 32 * 17;
  double acc = 0.0;
  const unsigned long long pa = (unsigned long long)a, pb = (unsigned long long)b;
  for (int k0 = 0; k0 < K; k0 += 16) {
#pragma unroll
    for (int i = 0; i < 8; i++) {
      const int r = i * 4 + (lane >> 3), c = 2 * (lane & 7);
      const double* ra = (const double*)__shfl_sync(kFullMask, pa, r);
      const double* rb = (const double*)__shfl_sync(kFullMask, pb, r);
      double2 va = make_double2(0.0, 0.0), vb = va;
      if (ra && k0 + c < K) {      // rows are padded to a multiple of 16 doubles (LD), so a 16-byte read is in bounds
        va = ldg2(ra + k0 + c);
        vb = ldg2(rb + k0 + c);
      }
      sa[r * 17 + c] = va.x; sa[r * 17 + c + 1] = va.y;
      sb[r * 17 + c] = vb.x; sb[r * 17 + c + 1] = vb.y;
    }
    __syncwarp();
    const int kend = min(16, K - k0);
    for (int kk = 0; kk < kend; kk++) acc = __dadd_rn(acc, __dmul_rn(sa[lane * 17 + kk], sb[lane * 17 + kk]));
    __syncwarp();
  }
  return acc;
}

// gt_score[s] = predict(user(s), gt[s]).  `users` == nullptr: user(s) = u_begin + s.  128 threads per block.
__global__ void __launch_bounds__(128)
eval_gt_score_kernel(const double* __restrict__ U, const double* __restrict__ V,
                     const int32_t* __restrict__ gt, const int32_t* __restrict__ users,
                     int u_begin, int count, int K, int LD, double* __restrict__ out) {
  __shared__ double sm[4][2 * 32 * 17];
  const int s = blockIdx.x * blockDim.x + threadIdx.x;
  const double *a = nullptr, *b = nullptr;
  if (s < count) {
    const int u = users ? users[s] : u_begin + s;
    a = U + (size_t)u * LD;
    b = V + (size_t)gt[s] * LD;
  }
  const double acc = seq_dot_warp32(a, b, K, sm[threadIdx.x >> 5]);
  if (s < count) out[s] = acc;
}

// out[a] = src[index[a]]
__global__ void gather_i32_kernel(const int32_t* __restrict__ src, const int32_t* __restrict__ index, int n,
                                  int32_t* __restrict__ out) {
  const int a = blockIdx.x * blockDim.x + threadIdx.x;
  if (a < n) out[a] = src[index[a]];
}

constexpr int kEvalTile = 64;     // users x items per CTA tile
constexpr int kEvalThreads = 256; // 16 x 16 threads, 4 x 4 scores each
constexpr int kEvalKC = 16;       // factor chunk staged per step

struct EvalTriple {
  int32_t slot;
  int32_t item;
  int32_t key;
};

// MODE 0: count_larger[s] += #{ i in [item_begin, n_items) : score(s,i) > gt_score[s] }.
// MODE 1: append (s, i, (int)score) for every (s,i) with (int)score != 0 to `triples`.
// `active` (nullable): the a-th row of the launch is slot active[a]; n_slots = number of rows.
template <int MODE>
__global__ void __launch_bounds__(kEvalThreads)
eval_tile_kernel(const double* __restrict__ U, const double* __restrict__ V,
                 const int32_t* __restrict__ users, const int32_t* __restrict__ active, int u_begin, int n_slots,
                 int item_begin, int n_items, int K, int LD,
                 const double* __restrict__ gt_score, int32_t* __restrict__ count_larger,
                 EvalTriple* __restrict__ triples, unsigned long long* __restrict__ n_triples,
                 unsigned long long cap_triples) {
  __shared__ double Us[kEvalKC][kEvalTile + 1];
  __shared__ double Vs[kEvalKC][kEvalTile + 1];
  const int tid = threadIdx.x, ty = tid >> 4, tx = tid & 15;
  const int item_tiles = (n_items - item_begin + kEvalTile - 1) / kEvalTile;
  const int ut = blockIdx.x / item_tiles, it = blockIdx.x % item_tiles;
  const int s0 = ut * kEvalTile, i0 = item_begin + it * kEvalTile;

  double acc[4][4];
#pragma unroll
  for (int a = 0; a < 4; a++)
#pragma unroll
    for (int b = 0; b < 4; b++) acc[a][b] = 0.0;

  for (int k0 = 0; k0 < K; k0 += kEvalKC) {
    // stage: 64 rows x 16 factors per side, transposed into [k][row]
    for (int t = tid; t < kEvalTile * kEvalKC; t += kEvalThreads) {
      const int r = t / kEvalKC, kk = t % kEvalKC;
      double uv = 0.0, vv = 0.0;
      if (k0 + kk < K) {
        const int sa = s0 + r;
        if (sa < n_slots) {
          const int s = active ? active[sa] : sa;
          const int u = users ? users[s] : u_begin + s;
          uv = U[(size_t)u * LD + k0 + kk];
        }
        const int i = i0 + r;
        if (i < n_items) vv = V[(size_t)i * LD + k0 + kk];
      }
      Us[kk][r] = uv;
      Vs[kk][r] = vv;
    }
    __syncthreads();
    const int kend = min(kEvalKC, K - k0);
    for (int kk = 0; kk < kend; kk++) {
      double av[4], bv[4];
#pragma unroll
      for (int a = 0; a < 4; a++) av[a] = Us[kk][ty + 16 * a];
#pragma unroll
      for (int b = 0; b < 4; b++) bv[b] = Vs[kk][tx + 16 * b];
#pragma unroll
      for (int a = 0; a < 4; a++)
#pragma unroll
        for (int b = 0; b < 4; b++) acc[a][b] = __dadd_rn(acc[a][b], __dmul_rn(av[a], bv[b]));
    }
    __syncthreads();
  }

#pragma unroll
  for (int a = 0; a < 4; a++) {
    const int sa = s0 + ty + 16 * a;
    const bool live = sa < n_slots;
    const int s = live ? (active ? active[sa] : sa) : 0;
    if (MODE == 0) {
      const double g = live ? gt_score[s] : 0.0;
      int c = 0;
#pragma unroll
      for (int b = 0; b < 4; b++) {
        const int i = i0 + tx + 16 * b;
        if (live && i < n_items && acc[a][b] > g) c++;
      }
      // 16 threads (same ty) share a user: reduce inside the half-warp before the atomic
      c += __shfl_xor_sync(kFullMask, c, 1);
      c += __shfl_xor_sync(kFullMask, c, 2);
      c += __shfl_xor_sync(kFullMask, c, 4);
      c += __shfl_xor_sync(kFullMask, c, 8);
      if (live && tx == 0 && c) atomicAdd(count_larger + s, c);
    } else {
#pragma unroll
      for (int b = 0; b < 4; b++) {
        const int i = i0 + tx + 16 * b;
        if (!live || i >= n_items) continue;
        const int key = __double2int_rz(acc[a][b]);
        if (key != 0) {
          const unsigned long long pos = atomicAdd(n_triples, 1ull);
          if (pos < cap_triples) triples[pos] = EvalTriple{s, i, key};
        }
      }
    }
  }
}

}  // namespace eals
