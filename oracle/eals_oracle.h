/* TEST INFRASTRUCTURE ONLY — the CPU oracle.  Never linked into, imported by or called from the
 * product path (eals_cpp_b200/, include/).  Only tests/, __graft_entry__.smoke() and bench.py's
 * cpu_baseline / --impl reference legs may use it, and only as the checker / reported baseline.
 *
 * eals_oracle: a plain-C restatement of the reference's eALS hot path (QihanW/eals_cpp,
 * MF_fastALS.cpp + DenseMat.cpp) on flat CSR/CSC arrays.  Every function cites the reference
 * lines whose arithmetic — including operation ORDER, so results are bit-identical when built
 * with -ffp-contract=off — it follows.  The containers are not restated: the reference's
 * arrays-of-SparseVec / double** are replaced by row_ptr/col_idx (rows ascending by item id) and
 * col_ptr/row_idx (columns ascending by user id) and contiguous row-major fp64 matrices, which is
 * the layout contract of the product.
 *
 * Parity status: PINNED.  tests/test_oracle_vs_reference.py checks every function here
 * bit-for-bit (or to the last ulp where libm `pow` is involved) against oracle/_ref/libeals_ref.so,
 * i.e. the reference's own unmodified translation units compiled by oracle/Makefile, and against
 * the committed fixtures in tests/golden/ that were generated from that library.
 */
#ifndef EALS_ORACLE_H
#define EALS_ORACLE_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* DenseMat::init(mean, sigma) — DenseMat.cpp:54-62 with libstdc++ 13 semantics
 * (minstd_rand0 seed 1; generate_canonical<double,53>; Marsaglia polar normal with saved value).
 * Fills n = rows*cols values in row-major order from a FRESH default-seeded engine. */
void eo_normal_fill(double* out, size_t n, double mean, double sigma);

/* Wi[i] = w0 * p_i^alpha / sum_j p_j^alpha, p_i = colcount_i / nnz — MF_fastALS.cpp:55-72. */
void eo_item_weights(int N, const int64_t* col_ptr, double w0, double alpha, double* Wi);

/* SU = U^T U the way initS does it (MF_fastALS.cpp:584 -> DenseMat.cpp:94-139). */
void eo_gram_plain(const double* X, int rows, int K, double* S);
/* SV[f][k] = sum_i V[i][f]*V[i][k]*Wi[i], k<=f, mirrored — MF_fastALS.cpp:585-594. */
void eo_gram_weighted(const double* X, const double* w, int rows, int K, double* S);

/* One user half-epoch: update_user_thread for u in [u_begin,u_end) then update_user_SU for the
 * same rows — MF_fastALS.cpp:119-132, 243-322, 324-335.  row_val NULL means all ratings are 1
 * (and W = copy of ratings, MF_fastALS.cpp:75-82, so w = rating).  Reads V,SV,Wi; writes U,SU. */
void eo_update_user_sweep(int M, int N, int K, const int64_t* row_ptr, const int32_t* col_idx,
                          const double* row_val, double* U, const double* V, double* SU,
                          const double* SV, const double* Wi, double reg, int u_begin, int u_end,
                          int patch_S);

/* One item half-epoch — MF_fastALS.cpp:137-152, 338-407, 409-422.  Reads U,SU,Wi; writes V,SV. */
void eo_update_item_sweep(int M, int N, int K, const int64_t* col_ptr, const int32_t* row_idx,
                          const double* col_val, const double* U, double* V, const double* SU,
                          double* SV, const double* Wi, double reg, int i_begin, int i_end,
                          int patch_S);

/* loss() — MF_fastALS.cpp:184-206 (+ predict :208-221, DenseMat::squaredSum DenseMat.cpp:86-92,
 * DenseMat::mult(DenseVec) :141-147, DenseVec::inner DenseVec.cpp:86-95). */
double eo_loss(int M, int N, int K, const int64_t* row_ptr, const int32_t* col_idx,
               const double* row_val, const double* U, const double* V, const double* SV,
               const double* Wi, double reg);

/* evaluate_for_user for all users + the means of main.cpp:60-62 — MF_fastALS.cpp:597-662.
 * compat != 0: reproduce the reference's int-truncating comparator + libstdc++
 * partial_sort_copy heap order (stl_algo.h:1647-1678, stl_heap.h).  compat == 0: exact ranking
 * (position = number of strictly larger scores).  Per-user outputs may be NULL.
 * count_larger[u] receives the number of items scoring strictly above the held-out item, capped
 * at topK+1 (the reference stops counting there, :634). */
void eo_evaluate(int M, int N, int K, const double* U, const double* V, const int32_t* gt_items,
                 int topK, int compat, double* hr, double* ndcg, double* prec,
                 int32_t* count_larger, double out_mean[3]);

#ifdef __cplusplus
}
#endif
#endif
