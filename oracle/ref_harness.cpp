// TEST INFRASTRUCTURE ONLY — never linked into, imported by or called from the product path.
//
// ref_harness.cpp: a thin C-ABI shim around the UNMODIFIED reference implementation
// (/root/reference/{DenseMat,DenseVec,SparseMat,SparseVec,MF_fastALS}.cpp), compiled where the
// sources lie by oracle/Makefile into oracle/_ref/libeals_ref.so.  It lets the Python tests and
// bench.py's reference arm drive the real MF_fastALS object from CSR arrays without the
// reference's text-file loader, and read back U/V/SU/SV/Wi, the loss and per-user evaluation
// tuples at full precision.
//
// Nothing here restates reference arithmetic: every number comes out of the reference's own
// member functions.  The only logic of ours is (1) building the SparseMat exactly as
// main.cpp:165,192-205 does (setLength for rows/cols, then setValue in (u asc, i asc) order) and
// (2) the sweep drivers, which repeat the call ORDER of MF_fastALS::buildModel
// (MF_fastALS.cpp:115-160: clone U, user sweep, SU patch, clone V, item sweep, SV patch) so a
// test can stop between half-epochs.  ref_build_model() calls the real buildModel() as a
// cross-check of (2).
//
// The reference object is deliberately leaked / never copied: its destructor double-frees when a
// copy exists (main.cpp:37, MF_fastALS.cpp:664-673).

#include <fcntl.h>
#include <unistd.h>

#include <cstdint>
#include <cstdio>
#include <cstring>
#include <ctime>
#include <iostream>
#include <sstream>
#include <vector>

#define private public     // reach MF_fastALS::initS (MF_fastALS.h:78); all data is public already
#include "MF_fastALS.h"
#undef private

namespace {

struct RefModel {
  MF_fastALS* fals;
  int M, N, K;
};

SparseMat build_sparse(int M, int N, const int64_t* row_ptr, const int32_t* col_idx,
                       const double* val) {
  // main.cpp:165 — SparseMat(userCount, itemCount); :192-197 setLength; :198-205 setValue.
  SparseMat sm(M, N);
  std::vector<int> col_count(N, 0);
  for (int u = 0; u < M; u++)
    for (int64_t p = row_ptr[u]; p < row_ptr[u + 1]; p++) col_count[col_idx[p]]++;
  for (int u = 0; u < M; u++) sm.rows[u].setLength((int)(row_ptr[u + 1] - row_ptr[u]));
  for (int i = 0; i < N; i++) sm.cols[i].setLength(col_count[i]);
  for (int u = 0; u < M; u++)
    for (int64_t p = row_ptr[u]; p < row_ptr[u + 1]; p++)
      sm.setValue(u, col_idx[p], val ? val[p] : 1.0);
  return sm;
}

// Silences the reference's std::cout chatter by pointing fd 1 at /dev/null for the scope.
struct StdoutSilencer {
  int saved;
  StdoutSilencer() {
    std::cout.flush();
    fflush(stdout);
    saved = dup(1);
    int nul = open("/dev/null", O_WRONLY);
    if (nul >= 0) { dup2(nul, 1); close(nul); }
  }
  ~StdoutSilencer() {
    std::cout.flush();
    fflush(stdout);
    if (saved >= 0) { dup2(saved, 1); close(saved); }
  }
};

}  // namespace

extern "C" {

// Construct the reference model.  test_items may be NULL (then every test item is 0).
void* ref_create(int M, int N, const int64_t* row_ptr, const int32_t* col_idx, const double* val,
                 const int32_t* test_items, int topK, int factors, int maxIter, double w0,
                 double alpha, double reg, double init_mean, double init_stdev) {
  SparseMat train = build_sparse(M, N, row_ptr, col_idx, val);
  std::vector<Rating> tests;
  tests.reserve(M);
  for (int u = 0; u < M; u++) tests.push_back(Rating(u, test_items ? test_items[u] : 0, 1.0f, 0));
  RefModel* r = new RefModel;
  r->M = M; r->N = N; r->K = factors;
  r->fals = new MF_fastALS(train, tests, topK, /*threadNum*/ 1, factors, maxIter, w0, alpha, reg,
                           init_mean, init_stdev, /*showProgress*/ false, /*showLoss*/ true, M, N);
  return r;
}

static void copy_out(DenseMat& m, int rows, int cols, double* out) {
  for (int r = 0; r < rows; r++) std::memcpy(out + (size_t)r * cols, m.matrix[r], sizeof(double) * cols);
}

void ref_get_U(void* h, double* out) { RefModel* r = (RefModel*)h; copy_out(r->fals->U, r->M, r->K, out); }
void ref_get_V(void* h, double* out) { RefModel* r = (RefModel*)h; copy_out(r->fals->V, r->N, r->K, out); }
void ref_get_SU(void* h, double* out) { RefModel* r = (RefModel*)h; copy_out(r->fals->SU, r->K, r->K, out); }
void ref_get_SV(void* h, double* out) { RefModel* r = (RefModel*)h; copy_out(r->fals->SV, r->K, r->K, out); }
void ref_get_Wi(void* h, double* out) {
  RefModel* r = (RefModel*)h;
  std::memcpy(out, r->fals->Wi, sizeof(double) * r->N);
}

// Overwrite the factors in place and rebuild the S caches with the reference's own initS().
// (MF_fastALS::setUV is unusable: DenseMat::clone takes sizeof of a pointer, DenseMat.cpp:23-37.)
void ref_set_UV(void* h, const double* U, const double* V) {
  RefModel* r = (RefModel*)h;
  for (int u = 0; u < r->M; u++) std::memcpy(r->fals->U.matrix[u], U + (size_t)u * r->K, sizeof(double) * r->K);
  for (int i = 0; i < r->N; i++) std::memcpy(r->fals->V.matrix[i], V + (size_t)i * r->K, sizeof(double) * r->K);
  r->fals->initS();
}

void ref_set_Wi(void* h, const double* Wi) {
  RefModel* r = (RefModel*)h;
  std::memcpy(r->fals->Wi, Wi, sizeof(double) * r->N);
  r->fals->initS();
}

// One user half-epoch in buildModel's order (MF_fastALS.cpp:119-132). Returns clock() seconds of
// the region the reference itself times (:125-133).
double ref_update_user_sweep(void* h) {
  RefModel* r = (RefModel*)h;
  MF_fastALS& f = *r->fals;
  std::vector<double> old((size_t)r->M * r->K);
  copy_out(f.U, r->M, r->K, old.data());
  clock_t t0 = clock();
  for (int u = 0; u < r->M; u++) f.update_user_thread(u);
  for (int u = 0; u < r->M; u++) f.update_user_SU(old.data() + (size_t)u * r->K, f.U.matrix[u]);
  return (double)(clock() - t0) / CLOCKS_PER_SEC;
}

// One item half-epoch in buildModel's order (MF_fastALS.cpp:137-152).
double ref_update_item_sweep(void* h) {
  RefModel* r = (RefModel*)h;
  MF_fastALS& f = *r->fals;
  std::vector<double> old((size_t)r->N * r->K);
  copy_out(f.V, r->N, r->K, old.data());
  clock_t t0 = clock();
  for (int i = 0; i < r->N; i++) f.update_item_thread(i);
  for (int i = 0; i < r->N; i++) f.update_item_SV(i, old.data() + (size_t)i * r->K, f.V.matrix[i]);
  return (double)(clock() - t0) / CLOCKS_PER_SEC;
}

// Partial sweeps over a row range, for the bounded CPU-baseline sample in bench.py.  Same calls as
// above restricted to [begin, end); the S cache is patched for those rows only.
double ref_update_user_range(void* h, int begin, int end) {
  RefModel* r = (RefModel*)h;
  MF_fastALS& f = *r->fals;
  std::vector<double> old((size_t)(end - begin) * r->K);
  for (int u = begin; u < end; u++) std::memcpy(old.data() + (size_t)(u - begin) * r->K, f.U.matrix[u], sizeof(double) * r->K);
  clock_t t0 = clock();
  for (int u = begin; u < end; u++) f.update_user_thread(u);
  for (int u = begin; u < end; u++) f.update_user_SU(old.data() + (size_t)(u - begin) * r->K, f.U.matrix[u]);
  return (double)(clock() - t0) / CLOCKS_PER_SEC;
}

double ref_update_item_range(void* h, int begin, int end) {
  RefModel* r = (RefModel*)h;
  MF_fastALS& f = *r->fals;
  std::vector<double> old((size_t)(end - begin) * r->K);
  for (int i = begin; i < end; i++) std::memcpy(old.data() + (size_t)(i - begin) * r->K, f.V.matrix[i], sizeof(double) * r->K);
  clock_t t0 = clock();
  for (int i = begin; i < end; i++) f.update_item_thread(i);
  for (int i = begin; i < end; i++) f.update_item_SV(i, old.data() + (size_t)(i - begin) * r->K, f.V.matrix[i]);
  return (double)(clock() - t0) / CLOCKS_PER_SEC;
}

// Single-row forms, no S patch (what MF_fastALS::update_user_thread / update_item_thread do alone).
void ref_update_user_row(void* h, int u) { ((RefModel*)h)->fals->update_user_thread(u); }
void ref_update_item_row(void* h, int i) { ((RefModel*)h)->fals->update_item_thread(i); }

double ref_loss(void* h) { return ((RefModel*)h)->fals->loss(); }
double ref_predict(void* h, int u, int i) { return ((RefModel*)h)->fals->predict(u, i); }

// The real buildModel() (stdout captured).  iters overrides maxIter.  Writes the `loss:` values the
// reference printed (6 significant digits) is pointless, so losses are recomputed by loss() only
// if the caller asks for the final one.
void ref_build_model(void* h, int iters) {
  RefModel* r = (RefModel*)h;
  r->fals->maxIter = iters;
  bool keep = r->fals->showloss;
  r->fals->showloss = false;
  StdoutSilencer quiet;
  r->fals->buildModel();
  r->fals->showloss = keep;
}

// evaluate_for_user for every user (main.cpp:46-54), per-user tuples out; returns the three means
// exactly as main.cpp:60-62 forms them (std::accumulate from 0.0, then / size).
void ref_evaluate(void* h, const int32_t* gt_items, int topK, double* hr, double* ndcg, double* prec,
                  double out_mean[3]) {
  RefModel* r = (RefModel*)h;
  double s0 = 0, s1 = 0, s2 = 0;
  for (int u = 0; u < r->M; u++) {
    std::vector<double> res = r->fals->evaluate_for_user(u, gt_items[u], topK);
    if (hr) hr[u] = res[0];
    if (ndcg) ndcg[u] = res[1];
    if (prec) prec[u] = res[2];
    s0 += res[0]; s1 += res[1]; s2 += res[2];
  }
  out_mean[0] = s0 / r->M; out_mean[1] = s1 / r->M; out_mean[2] = s2 / r->M;
}

// DenseMat::init known-answer access (DenseMat.cpp:54-62) without building a model.
void ref_dense_init(int rows, int cols, double mean, double sigma, double* out) {
  DenseMat m(rows, cols);
  m.init(mean, sigma);
  copy_out(m, rows, cols, out);
}

}  // extern "C"
