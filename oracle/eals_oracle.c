/* TEST INFRASTRUCTURE ONLY — see eals_oracle.h.  Plain C11, build with -ffp-contract=off. */
#include "eals_oracle.h"

#include <math.h>
#include <stdlib.h>
#include <string.h>

/* ------------------------------------------------------------------------------------------
 * a2. DenseMat::init(mean, sigma) — DenseMat.cpp:54-62.
 * std::default_random_engine in libstdc++ is minstd_rand0 = LCG(a=16807, c=0, m=2^31-1), default
 * seed 1 (bits/random.h:1592-1593,1641).  std::normal_distribution<double> draws through
 * generate_canonical<double,53> (bits/random.tcc:3345-3381): range r = max-min+1 = 2^31-2, so
 * log2r = 30 and m = (53+30-1)/30 = 2 engine draws per uniform, combined as
 * (x1-1) + (x2-1)*r over r*r with the running scale held in double but multiplied in long double.
 * The normal itself is the Marsaglia polar method (random.tcc:1811-1847): the FIRST value returned
 * from a fresh pair is y*mult, x*mult is saved for the next call.
 * ------------------------------------------------------------------------------------------ */
typedef struct {
  uint32_t x;
  int have_saved;
  double saved;
} eo_rng;

static uint32_t eo_minstd_next(eo_rng* g) {
  g->x = (uint32_t)(((uint64_t)g->x * 16807u) % 2147483647u);
  return g->x;
}

static double eo_canonical(eo_rng* g) {
  const long double r = 2147483646.0L; /* max() - min() + 1 with min()=1, max()=2^31-2 */
  double sum = 0.0, tmp = 1.0;
  for (int k = 2; k != 0; --k) {
    sum += (double)(eo_minstd_next(g) - 1u) * tmp;
    tmp = (double)((long double)tmp * r);
  }
  double ret = sum / tmp;
  if (ret >= 1.0) ret = nextafter(1.0, 0.0);
  return ret;
}

static double eo_normal(eo_rng* g, double mean, double sigma) {
  double ret;
  if (g->have_saved) {
    g->have_saved = 0;
    ret = g->saved;
  } else {
    double x, y, r2;
    do {
      x = 2.0 * eo_canonical(g) - 1.0;
      y = 2.0 * eo_canonical(g) - 1.0;
      r2 = x * x + y * y;
    } while (r2 > 1.0 || r2 == 0.0);
    const double mult = sqrt(-2 * log(r2) / r2);
    g->saved = x * mult;
    g->have_saved = 1;
    ret = y * mult;
  }
  return ret * sigma + mean;
}

void eo_normal_fill(double* out, size_t n, double mean, double sigma) {
  eo_rng g = {1u, 0, 0.0};
  for (size_t i = 0; i < n; i++) out[i] = eo_normal(&g, mean, sigma);
}

/* ------------------------------------------------------------------------------------------
 * a1. Popularity weights — MF_fastALS.cpp:55-72: p = column counts; sum; p/=sum; p=pow(p,alpha);
 * Z = sum of those; Wi = w0*p/Z.
 * ------------------------------------------------------------------------------------------ */
void eo_item_weights(int N, const int64_t* col_ptr, double w0, double alpha, double* Wi) {
  double sum = 0, Z = 0;
  double* p = (double*)malloc(sizeof(double) * (size_t)(N > 0 ? N : 1));
  for (int i = 0; i < N; i++) {
    p[i] = (double)(int)(col_ptr[i + 1] - col_ptr[i]);
    sum += p[i];
  }
  for (int i = 0; i < N; i++) {
    p[i] /= sum;
    p[i] = pow(p[i], alpha);
    Z += p[i];
  }
  for (int i = 0; i < N; i++) Wi[i] = w0 * p[i] / Z;
  free(p);
}

/* ------------------------------------------------------------------------------------------
 * a3. initS — MF_fastALS.cpp:583-595.
 * SU = U.transpose().mult(U): entry (i,j) = sum over rows r (ascending) of U[r][i]*U[r][j]
 * (DenseMat.cpp:125-139 on the transposed copy).  SV: only k<=f computed, product is
 * (V[i][f]*V[i][k])*Wi[i], then mirrored.
 * ------------------------------------------------------------------------------------------ */
void eo_gram_plain(const double* X, int rows, int K, double* S) {
  for (int i = 0; i < K; i++)
    for (int j = 0; j < K; j++) {
      double product = 0;
      for (int r = 0; r < rows; r++) product += X[(size_t)r * K + i] * X[(size_t)r * K + j];
      S[(size_t)i * K + j] = product;
    }
}

void eo_gram_weighted(const double* X, const double* w, int rows, int K, double* S) {
  for (int f = 0; f < K; f++)
    for (int k = 0; k <= f; k++) {
      double val = 0;
      for (int i = 0; i < rows; i++) val += X[(size_t)i * K + f] * X[(size_t)i * K + k] * w[i];
      S[(size_t)f * K + k] = val;
      S[(size_t)k * K + f] = val;
    }
}

/* predict — MF_fastALS.cpp:208-221: sequential k, res += a*b. */
static double eo_dot_seq(const double* a, const double* b, int K) {
  double res = 0;
  for (int k = 0; k < K; k++) res += a[k] * b[k];
  return res;
}

/* ------------------------------------------------------------------------------------------
 * a5 + a7. User half-epoch.
 * update_user_thread (MF_fastALS.cpp:243-322): skip empty rows (:249); per nonzero cache
 * pred=<u,v_i>, rating, w (=rating: W is a copy of the train values, :75-82) (:261-270); then for
 * each factor f in order (:276-316) the numerator starts from -sum_{k!=f} u_k*SV[f][k]
 * (:284-287), the loop over the row's items (ascending item id) removes factor f from the cached
 * prediction, accumulates numer/denom (:297-305), denom += SV[f][f]+reg (:307), u_f=numer/denom
 * (:310) and the cache is refreshed (:314-315).
 * update_user_SU (:324-335): SU[f][k] = SU[f][k] - old_f*old_k + new_f*new_k for k<=f, mirrored,
 * applied for every row AFTER the whole sweep (:130-132) — empty rows included (old==new there).
 * ------------------------------------------------------------------------------------------ */
void eo_update_user_sweep(int M, int N, int K, const int64_t* row_ptr, const int32_t* col_idx,
                          const double* row_val, double* U, const double* V, double* SU,
                          const double* SV, const double* Wi, double reg, int u_begin, int u_end,
                          int patch_S) {
  (void)M; (void)N;
  int64_t max_n = 0;
  for (int u = u_begin; u < u_end; u++)
    if (row_ptr[u + 1] - row_ptr[u] > max_n) max_n = row_ptr[u + 1] - row_ptr[u];
  double* pred = (double*)malloc(sizeof(double) * (size_t)(max_n + 1));
  double* old = (double*)malloc(sizeof(double) * (size_t)(u_end - u_begin + 1) * K);
  memcpy(old, U + (size_t)u_begin * K, sizeof(double) * (size_t)(u_end - u_begin) * K);

  for (int u = u_begin; u < u_end; u++) {
    const int64_t p0 = row_ptr[u];
    const int n = (int)(row_ptr[u + 1] - p0);
    if (n == 0) continue;
    double* uget = U + (size_t)u * K;
    for (int j = 0; j < n; j++) pred[j] = eo_dot_seq(uget, V + (size_t)col_idx[p0 + j] * K, K);
    for (int f = 0; f < K; f++) {
      double numer = 0, denom = 0;
      const double* svget = SV + (size_t)f * K;
      for (int k = 0; k < K; k++)
        if (k != f) numer -= uget[k] * svget[k];
      const double ufget = uget[f];
      for (int j = 0; j < n; j++) {
        const int i = col_idx[p0 + j];
        const double rating = row_val ? row_val[p0 + j] : 1.0;
        const double w = rating;
        const double ifv = V[(size_t)i * K + f];
        pred[j] -= ufget * ifv;
        numer += (w * rating - (w - Wi[i]) * pred[j]) * ifv;
        denom += (w - Wi[i]) * ifv * ifv;
      }
      denom += svget[f] + reg;
      uget[f] = numer / denom;
      for (int j = 0; j < n; j++) pred[j] += uget[f] * V[(size_t)col_idx[p0 + j] * K + f];
    }
  }
  if (patch_S) {
    for (int u = u_begin; u < u_end; u++) {
      const double* o = old + (size_t)(u - u_begin) * K;
      const double* nw = U + (size_t)u * K;
      for (int f = 0; f < K; f++)
        for (int k = 0; k <= f; k++) {
          double val = SU[(size_t)f * K + k] - o[f] * o[k] + nw[f] * nw[k];
          SU[(size_t)f * K + k] = val;
          SU[(size_t)k * K + f] = val;
        }
    }
  }
  free(pred);
  free(old);
}

/* ------------------------------------------------------------------------------------------
 * a6 + a7. Item half-epoch — update_item_thread (MF_fastALS.cpp:338-407) and update_item_SV
 * (:409-422).  Differences from the user side: the S term is scaled by the row's own weight
 * (numer *= Wi[i] at :380, denom += Wi[i]*SU[f][f] + reg at :391) and Wi[i] is constant over the
 * inner loop (:388-389).  Users are visited in ascending user id (column order).
 * SV patch: SV[f][k] - old_f*old_k*Wi[i] + new_f*new_k*Wi[i].
 * ------------------------------------------------------------------------------------------ */
void eo_update_item_sweep(int M, int N, int K, const int64_t* col_ptr, const int32_t* row_idx,
                          const double* col_val, const double* U, double* V, const double* SU,
                          double* SV, const double* Wi, double reg, int i_begin, int i_end,
                          int patch_S) {
  (void)M; (void)N;
  int64_t max_n = 0;
  for (int i = i_begin; i < i_end; i++)
    if (col_ptr[i + 1] - col_ptr[i] > max_n) max_n = col_ptr[i + 1] - col_ptr[i];
  double* pred = (double*)malloc(sizeof(double) * (size_t)(max_n + 1));
  double* old = (double*)malloc(sizeof(double) * (size_t)(i_end - i_begin + 1) * K);
  memcpy(old, V + (size_t)i_begin * K, sizeof(double) * (size_t)(i_end - i_begin) * K);

  for (int i = i_begin; i < i_end; i++) {
    const int64_t p0 = col_ptr[i];
    const int n = (int)(col_ptr[i + 1] - p0);
    if (n == 0) continue;
    double* vget = V + (size_t)i * K;
    const double wi = Wi[i];
    for (int j = 0; j < n; j++) pred[j] = eo_dot_seq(U + (size_t)row_idx[p0 + j] * K, vget, K);
    for (int f = 0; f < K; f++) {
      double numer = 0, denom = 0;
      const double* suget = SU + (size_t)f * K;
      for (int k = 0; k < K; k++)
        if (k != f) numer -= vget[k] * suget[k];
      numer *= wi;
      const double ifv = vget[f];
      for (int j = 0; j < n; j++) {
        const int u = row_idx[p0 + j];
        const double rating = col_val ? col_val[p0 + j] : 1.0;
        const double w = rating;
        const double ufu = U[(size_t)u * K + f];
        pred[j] -= ufu * ifv;
        numer += (w * rating - (w - wi) * pred[j]) * ufu;
        denom += (w - wi) * ufu * ufu;
      }
      denom += wi * suget[f] + reg;
      vget[f] = numer / denom;
      for (int j = 0; j < n; j++) pred[j] += U[(size_t)row_idx[p0 + j] * K + f] * vget[f];
    }
  }
  if (patch_S) {
    for (int i = i_begin; i < i_end; i++) {
      const double* o = old + (size_t)(i - i_begin) * K;
      const double* nw = V + (size_t)i * K;
      const double wi = Wi[i];
      for (int f = 0; f < K; f++)
        for (int k = 0; k <= f; k++) {
          double val = SV[(size_t)f * K + k] - o[f] * o[k] * wi + nw[f] * nw[k] * wi;
          SV[(size_t)f * K + k] = val;
          SV[(size_t)k * K + f] = val;
        }
    }
  }
  free(pred);
  free(old);
}

/* ------------------------------------------------------------------------------------------
 * a8. loss — MF_fastALS.cpp:184-206.
 * L = reg*(|U|^2+|V|^2) (row-major sequential sums, DenseMat.cpp:86-92), then per user
 * l = sum_i [ w*(r-pred)^2 - Wi[i]*pred^2 ] + (SV*u).u where SV*u is formed row by row
 * (DenseMat.cpp:141-147) and the final inner product runs over k ascending (DenseVec.cpp:86-95).
 * The reference writes pow(x,2); g++ folds that to x*x, which is also the correctly rounded value.
 * ------------------------------------------------------------------------------------------ */
double eo_loss(int M, int N, int K, const int64_t* row_ptr, const int32_t* col_idx,
               const double* row_val, const double* U, const double* V, const double* SV,
               const double* Wi, double reg) {
  double su = 0, sv = 0;
  for (size_t t = 0; t < (size_t)M * K; t++) su += U[t] * U[t];
  for (size_t t = 0; t < (size_t)N * K; t++) sv += V[t] * V[t];
  double L = reg * (su + sv);
  double* tmp = (double*)malloc(sizeof(double) * (size_t)K);
  for (int u = 0; u < M; u++) {
    double l = 0;
    const double* uu = U + (size_t)u * K;
    for (int64_t p = row_ptr[u]; p < row_ptr[u + 1]; p++) {
      const int i = col_idx[p];
      const double rating = row_val ? row_val[p] : 1.0;
      const double pred = eo_dot_seq(uu, V + (size_t)i * K, K);
      const double d = rating - pred;
      l += rating * (d * d);
      l -= Wi[i] * (pred * pred);
    }
    for (int f = 0; f < K; f++) tmp[f] = eo_dot_seq(SV + (size_t)f * K, uu, K);
    l += eo_dot_seq(tmp, uu, K);
    L += l;
  }
  free(tmp);
  return L;
}

/* ------------------------------------------------------------------------------------------
 * a9. Evaluation — MF_fastALS.cpp:620-662 + metric helpers :597-618 + means main.cpp:60-62.
 * The rank list the reference builds comes from std::partial_sort_copy over (item, score) pairs in
 * item order with a comparator whose parameters are pair<const int,int>: the score is converted
 * to int (truncation toward zero) before `>` (:647-651).  libstdc++'s algorithm
 * (stl_algo.h:1647-1678): copy the first topK elements, __make_heap, then for every later element
 * e: if comp(e, heap[0]) __adjust_heap(heap, 0, len, e); finally __sort_heap.  The helpers below
 * follow stl_heap.h:131-147 (__push_heap), :220-250 (__adjust_heap), :252-267 (__pop_heap),
 * :336-362 (__make_heap), :416-428 (__sort_heap) with comp(a,b) := a.key > b.key.
 * ------------------------------------------------------------------------------------------ */
typedef struct {
  int item;
  int key;
} eo_ent;

static int eo_comp(eo_ent a, eo_ent b) { return a.key > b.key; }

static void eo_push_heap(eo_ent* first, long hole, long top, eo_ent value) {
  long parent = (hole - 1) / 2;
  while (hole > top && eo_comp(first[parent], value)) {
    first[hole] = first[parent];
    hole = parent;
    parent = (hole - 1) / 2;
  }
  first[hole] = value;
}

static void eo_adjust_heap(eo_ent* first, long hole, long len, eo_ent value) {
  const long top = hole;
  long second = hole;
  while (second < (len - 1) / 2) {
    second = 2 * (second + 1);
    if (eo_comp(first[second], first[second - 1])) second--;
    first[hole] = first[second];
    hole = second;
  }
  if ((len & 1) == 0 && second == (len - 2) / 2) {
    second = 2 * (second + 1);
    first[hole] = first[second - 1];
    hole = second - 1;
  }
  eo_push_heap(first, hole, top, value);
}

static void eo_make_heap(eo_ent* first, long len) {
  if (len < 2) return;
  long parent = (len - 2) / 2;
  for (;;) {
    eo_ent value = first[parent];
    eo_adjust_heap(first, parent, len, value);
    if (parent == 0) return;
    parent--;
  }
}

static void eo_sort_heap(eo_ent* first, long len) {
  while (len > 1) {
    --len;
    eo_ent value = first[len];
    first[len] = first[0];
    eo_adjust_heap(first, 0, len, value);
  }
}

void eo_evaluate(int M, int N, int K, const double* U, const double* V, const int32_t* gt_items,
                 int topK, int compat, double* hr, double* ndcg, double* prec,
                 int32_t* count_larger, double out_mean[3]) {
  double s_hr = 0, s_ndcg = 0, s_prec = 0;
  double* score = (double*)malloc(sizeof(double) * (size_t)(N > 0 ? N : 1));
  eo_ent* top = (eo_ent*)malloc(sizeof(eo_ent) * (size_t)(topK > 0 ? topK : 1));
  for (int u = 0; u < M; u++) {
    const double* uu = U + (size_t)u * K;
    const int gt = gt_items[u];
    const double maxScore = eo_dot_seq(uu, V + (size_t)gt * K, K);
    int countLarger = 0, early = 0;
    for (int i = 0; i < N; i++) {
      score[i] = eo_dot_seq(uu, V + (size_t)i * K, K);
      if (score[i] > maxScore) countLarger++;
      if (countLarger > topK) { early = 1; break; }
    }
    double r0 = 0, r1 = 0, r2 = 0;
    if (!early) {
      int pos = -1;
      if (compat) {
        /* top_K(topK) is value-initialised to (0, 0.0) pairs (:642); slots beyond N stay so. */
        for (int t = 0; t < topK; t++) { top[t].item = 0; top[t].key = 0; }
        long real = 0;
        int i = 0;
        for (; i < N && real < topK; i++, real++) { top[real].item = i; top[real].key = (int)score[i]; }
        eo_make_heap(top, real);
        for (; i < N; i++) {
          eo_ent e = {i, (int)score[i]};
          if (eo_comp(e, top[0])) eo_adjust_heap(top, 0, real, e);
        }
        eo_sort_heap(top, real);
        for (int t = 0; t < topK; t++)
          if (top[t].item == gt) { pos = t; break; }
      } else {
        if (countLarger < topK) pos = countLarger;
      }
      if (pos >= 0) {
        r0 = 1;
        r1 = log(2) / log(pos + 2);
        r2 = 1.0 / (pos + 1);
      }
    }
    if (hr) hr[u] = r0;
    if (ndcg) ndcg[u] = r1;
    if (prec) prec[u] = r2;
    if (count_larger) count_larger[u] = countLarger;
    s_hr += r0; s_ndcg += r1; s_prec += r2;
  }
  out_mean[0] = s_hr / M; out_mean[1] = s_ndcg / M; out_mean[2] = s_prec / M;
  free(score);
  free(top);
}
