// cd_sweep.cuh — K1: the per-row coordinate-descent sweep of eALS.
//
// Replaces MF_fastALS::update_user_thread (MF_fastALS.cpp:243-322) and update_item_thread
// (:338-407).  For one row (a user, or an item on the transposed side) with nonzeros j = 1..n over
// neighbour rows y_j of the OTHER factor matrix, and the frozen S cache of the other side:
//
//   pred_j = <x, y_j>                                                      (:261-270 / :352-363)
//   for f = 0..K-1 (sequential — each step sees the factors updated before it):
//     numer = -g * sum_{k != f} x_k S[f][k]          g = 1 (user) | Wi[row] (item)   (:284-287 / :375-380)
//     pred_j -= x_f y_jf ; numer += (w_j r_j - c_j pred_j) y_jf ; denom += c_j y_jf^2     (:297-305 / :385-390)
//     denom += g S[f][f] + reg ;  x_f = numer / denom ; pred_j += x_f y_jf            (:307-315 / :391-400)
//   with c_j = w_j - Wi[item] (item = neighbour j on the user side, the row itself on the item side).
//
// Rows are independent inside a half-epoch (they only read the other side's factors, its S cache
// and Wi), so they are spread over warps / CTAs in any order; only the floating-point summation
// order over j differs from the reference's sequential loop.
//
// Data movement.  The n x K tile of gathered neighbour rows is what the K steps walk column by
// column.  It is staged through shared memory one 128-byte factor block (16 doubles of every
// gathered row = one full cache line each) at a time, so global memory sees whole lines exactly
// once per pass; the per-nonzero prediction cache lives in registers (warp kernel) or in a
// device-resident fp64 array (CTA kernel, rows of any length).
#pragma once

#include "common.cuh"

namespace eals {

// Replicas of the factor matrix being updated on the other GPUs of the box (CUDA IPC mappings of
// their X buffers): a finished row is stored straight into every replica over NVLink, which fuses
// the all-gather of the updated rows into the sweep.
constexpr int kMaxPeers = 7;
struct PeerSet {
  double* x[kMaxPeers];
  int n;
};

// The other orientation's prediction caches: rank r holds global positions [bound[r], bound[r+1]).
struct PcOut {
  double* base[8];
  uint32_t bound[9];
  int n;                   // ranks (0: no cache)
};

struct CdSide {
  const int64_t* ptr;   // [rows+1] offsets of the owned rows (first owned row at 0)
  const int32_t* idx;   // neighbour ids, ascending inside a row
  const double* val;    // rating (= confidence weight) per nonzero, or nullptr for all-ones
  double* X;            // factors being updated, full replica [n][LD]
  const double* Y;      // the other side's factors [n'][LD]
  const double* S;      // the other side's S cache [K][LD]
  const double* Wi;     // item popularity weights [n_items]
  int row_base;         // global id of owned row 0
  int K;
  double reg;
  // Symmetric prediction cache.  The prediction <u, v> a user sweep leaves behind for a nonzero is
  // exactly the value the item sweep starts from (and vice versa), so the gather pass that rebuilds
  // it (MF_fastALS.cpp:261-270 / 352-363) is skipped: each side READS its own cache sequentially
  // (pc_in, this side's nonzero order) and WRITES the other side's cache at the position of the
  // same nonzero in the other orientation (pc_map), on whichever rank owns that row (pc_out).
  const double* pc_in;     // nullptr: no cache
  const uint32_t* pc_map;  // per local nonzero: its global position in the other orientation
  PcOut pc_out;
  // Multi-rank: final predictions are first written to a LOCAL staging array, already in destination order
  // (pc_map then holds, per local nonzero, its slot in that order), and copied to their owners afterwards by
  // pc_route_kernel — a pure streaming copy.  Scattered 8-byte stores straight into peer memory were far
  // slower than the gather they replace (r01f); staging in source order and gathering in the copy kernel cost
  // 5 ms per half-epoch at 2 GPUs (r02i).
  double* pc_stage;        // nullptr: store directly through pc_map / pc_out
  int use_cache;           // 1: pc_in is valid on entry, read it instead of recomputing
  PeerSet peers;        // other ranks' replicas of X (n = 0: none)
  int fence;            // 1: every thread ends with a system-scope fence when it may have stored to a peer (peers_release)
};

__device__ __forceinline__ void store_row_value(const CdSide& a, size_t off, double v) {
  a.X[off] = v;
  for (int p = 0; p < a.peers.n; p++) a.peers.x[p][off] = v;
}

// Optional (EALS_PEER_FENCE=1): every thread of a sweep kernel ends with a system-scope fence after its last
// peer store.  Off by default: grid completion already performs the kernel's stores at system scope before
// anything ordered after it in the stream (the event the other ranks wait on, or the NCCL all-reduce) can
// start, and the fence is expensive — a CTA-per-row kernel cannot retire its CTA until the NVLink stores are
// acknowledged: 129.0 -> 120.5 ms per epoch at 4 GPUs without it (profiles/r2/r02w_*), same loss bits, replicas
// identical.  It was added in round 1 when one 8-GPU run ended 0.19 % off; that turned out to be the start-up
// race (no barrier between init and the first peer store, fixed in round 2), not visibility.
__device__ __forceinline__ void peers_release(const CdSide& a) {
  if (a.fence && (a.peers.n > 0 || (a.pc_out.n > 1 && !a.pc_stage))) __threadfence_system();
}

// Store a prediction at global position g of the other orientation's cache (on whichever rank holds it).
__device__ __forceinline__ void pc_store_at(const CdSide& a, uint32_t g, double v) {
  int r = 0;
#pragma unroll
  for (int t = 1; t < 8; t++) r += (t < a.pc_out.n && g >= a.pc_out.bound[t]) ? 1 : 0;
  a.pc_out.base[r][g - a.pc_out.bound[r]] = v;
}

// The final prediction of local nonzero local_pos (its position g in the other orientation already
// looked up, or not needed when the values are staged).
__device__ __forceinline__ void pc_emit(const CdSide& a, int64_t local_pos, uint32_t g, double v) {
  (void)local_pos;
  if (a.pc_stage) a.pc_stage[g] = v;      // g = slot in destination order
  else pc_store_at(a, g, v);              // g = global position in the other orientation
}

__device__ __forceinline__ void pc_store(const CdSide& a, int64_t local_pos, double v) {
  pc_emit(a, local_pos, a.pc_map[local_pos], v);
}

// Second phase of the staged scheme: slot k of the staging array (DESTINATION order: route_dst ascending, so
// neighbouring threads read neighbouring values and write neighbouring addresses of the same peer) goes to its
// owner.  A small persistent grid, launched on a side stream right behind the Gram's main kernel.
__global__ void __launch_bounds__(256)
pc_route_kernel(const double* __restrict__ stage, const uint32_t* __restrict__ route_dst, int64_t n, PcOut out) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; k < n; k += stride) {
    const uint32_t g = route_dst[k];
    int r = 0;
#pragma unroll
    for (int t = 1; t < 8; t++) r += (t < out.n && g >= out.bound[t]) ? 1 : 0;
    out.base[r][g - out.bound[r]] = stage[k];
  }
  if (out.n > 1) __threadfence_system();   // see peers_release
}

// inv[perm[k]] = k
__global__ void invert_perm_kernel(const uint32_t* __restrict__ perm, int64_t n, uint32_t* __restrict__ inv) {
  const int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (k < n) inv[perm[k]] = (uint32_t)k;
}

}  // namespace eals
