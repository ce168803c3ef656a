// eals_host_types.h — minimal host containers for drivers that do not bring the reference's own.
//
// SparseMat here is CSR/CSC-backed (two flat index arrays and two offset arrays) but exposes the
// member names MF_fastALS_T duck-types on — n_r, n_c, rows[u].{n,spv_in,spv_do}, cols[i].{...} —
// which are the names of the reference's SparseMat/SparseVec (SparseMat.h:38-41, SparseVec.h:24-27).
// Rating has the reference's four fields (Rating.h:8-24).
#ifndef EALS_B200_HOST_TYPES_H
#define EALS_B200_HOST_TYPES_H

#include <cstdint>
#include <map>
#include <vector>

namespace eals_b200 {

struct Rating {
  int userId = 0;
  int itemId = 0;
  float score = 0;
  long timestamp = 0;
  Rating() = default;
  Rating(int u, int i, float s, long t) : userId(u), itemId(i), score(s), timestamp(t) {}
};

struct SparseVecView {
  int n = 0;
  int* spv_in = nullptr;
  double* spv_do = nullptr;
};

class SparseMat {
 public:
  int n_r = 0, n_c = 0;
  SparseVecView* rows = nullptr;
  SparseVecView* cols = nullptr;

  SparseMat() = default;
  // Build both orientations from per-user ordered maps item -> value (ascending item ids), the
  // shape main.cpp:168-205 goes through: rows ascend by item, columns ascend by user.
  SparseMat(int n_users, int n_items, const std::vector<std::map<int, double>>& by_user) : n_r(n_users), n_c(n_items) {
    std::vector<int64_t> col_count((size_t)n_items + 1, 0);
    int64_t nnz = 0;
    for (const auto& m : by_user) { nnz += (int64_t)m.size(); for (const auto& kv : m) col_count[(size_t)kv.first + 1]++; }
    row_idx_.resize((size_t)nnz); row_val_.resize((size_t)nnz);
    col_idx_.resize((size_t)nnz); col_val_.resize((size_t)nnz);
    rows_.resize((size_t)n_users); cols_.resize((size_t)n_items);
    for (int i = 0; i < n_items; i++) col_count[(size_t)i + 1] += col_count[i];
    std::vector<int64_t> fill(col_count.begin(), col_count.end() - 1);
    int64_t p = 0;
    for (int u = 0; u < n_users; u++) {
      rows_[u].n = (int)by_user[u].size();
      rows_[u].spv_in = col_idx_.data() + p;
      rows_[u].spv_do = row_val_.data() + p;
      for (const auto& kv : by_user[u]) {
        col_idx_[(size_t)p] = kv.first; row_val_[(size_t)p] = kv.second; p++;
        const int64_t q = fill[kv.first]++;
        row_idx_[(size_t)q] = u; col_val_[(size_t)q] = kv.second;
      }
    }
    for (int i = 0; i < n_items; i++) {
      cols_[i].n = (int)(col_count[(size_t)i + 1] - col_count[i]);
      cols_[i].spv_in = row_idx_.data() + col_count[i];
      cols_[i].spv_do = col_val_.data() + col_count[i];
    }
    rows = rows_.data(); cols = cols_.data();
  }
  SparseMat(const SparseMat&) = delete;
  SparseMat& operator=(const SparseMat&) = delete;
  int64_t itemCount() const { return (int64_t)col_idx_.size(); }   // number of stored ratings

 private:
  std::vector<int> col_idx_, row_idx_;
  std::vector<double> row_val_, col_val_;
  std::vector<SparseVecView> rows_, cols_;
};

}  // namespace eals_b200
#endif
