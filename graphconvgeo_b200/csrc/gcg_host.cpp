// Host-side pieces of the hot path: k-d tree region labels, A_hat construction,
// CSR transpose / row gather.  Plain C++ (no CUDA), exported through the same C ABI.
#include <algorithm>
#include <cmath>
#include <cstring>
#include <limits>
#include <numeric>
#include <utility>
#include <vector>

#include "gcg_common.cuh"

using gcg::set_error;

// ---------------------------------------------------------------- k-d tree
// Restates kdtree.py:84-118 (nodeSplit) and :126-147 (KDTreeClustering.fit):
// float64 compares, np.median split value, "value > split -> right", children
// visited left then right, leaves numbered in that order.
namespace {

double np_median(std::vector<double>& v) {
  const size_t n = v.size();
  const size_t mid = n / 2;
  std::nth_element(v.begin(), v.begin() + mid, v.end());
  const double hi = v[mid];
  if (n % 2 == 1) return hi;
  const double lo = *std::max_element(v.begin(), v.begin() + mid);
  return (lo + hi) / 2.0;  // numpy: mean of the two middle values
}

}  // namespace

extern "C" int gcg_kdtree_fit_host(const double* pts, int64_t n, int32_t dims, int64_t bucket_size,
                                   int64_t* labels, int64_t* n_leaves) {
  GCG_CHECK_ARG(pts && labels && n >= 0 && dims > 0, "gcg_kdtree_fit_host: bad argument");
  std::vector<int64_t> ids(n), scratch(n);
  std::iota(ids.begin(), ids.end(), 0);
  // explicit DFS stack of [begin, end) ranges over `ids`; right pushed first
  std::vector<std::pair<int64_t, int64_t>> stack;
  stack.emplace_back(0, n);
  std::vector<double> vals, mins(dims), maxs(dims);
  int64_t next_leaf = 0;
  while (!stack.empty()) {
    const auto range = stack.back();
    stack.pop_back();
    const int64_t b = range.first, e = range.second, cnt = e - b;
    bool leaf = true;
    if (cnt > bucket_size) {                                      // kdtree.py:85 (strict)
      for (int d = 0; d < dims; ++d) {
        mins[d] = std::numeric_limits<double>::infinity();        // :22-23
        maxs[d] = -std::numeric_limits<double>::infinity();
      }
      for (int64_t i = b; i < e; ++i)
        for (int d = 0; d < dims; ++d) {
          const double x = pts[ids[i] * dims + d];
          mins[d] = std::min(mins[d], x);                         // :47-49
          maxs[d] = std::max(maxs[d], x);
        }
      int sd = 0;                                                 // np.argmax: first max wins (:51-54)
      double best = maxs[0] - mins[0];
      for (int d = 1; d < dims; ++d)
        if (maxs[d] - mins[d] > best) { best = maxs[d] - mins[d]; sd = d; }
      vals.resize(cnt);
      for (int64_t i = 0; i < cnt; ++i) vals[i] = pts[ids[b + i] * dims + sd];
      double sv = np_median(vals);                                // :88
      if (mins[sd] != maxs[sd]) {                                 // :90-91 zero width -> leaf
        if (sv == maxs[sd]) sv = mins[sd];                        // :94-95
        // stable partition: left (<= sv) keeps order, then right (> sv)   (:106-114)
        int64_t nl = 0, nr = 0;
        for (int64_t i = b; i < e; ++i) {
          const int64_t id = ids[i];
          if (pts[id * dims + sd] > sv) scratch[nr++] = id;
          else ids[b + nl++] = id;
        }
        std::memcpy(ids.data() + b + nl, scratch.data(), sizeof(int64_t) * nr);
        stack.emplace_back(b + nl, e);                            // right (visited second)
        stack.emplace_back(b, b + nl);                            // left (visited first)  :117-118
        leaf = false;
      }
    }
    if (leaf) {
      for (int64_t i = b; i < e; ++i) labels[ids[i]] = next_leaf;  // :143-145
      ++next_leaf;
    }
  }
  if (n_leaves) *n_leaves = next_leaf;
  return GCG_OK;
}

// ------------------------------------------------------------------ A_hat
extern "C" int64_t gcg_ahat_nnz_host(int64_t n, const int32_t* indptr, const int32_t* indices) {
  if (!indptr || (n > 0 && indptr[n] > 0 && !indices)) return -1;
  int64_t nnz = indptr[n];
  for (int64_t r = 0; r < n; ++r) {
    const int32_t* b = indices + indptr[r];
    const int32_t* e = indices + indptr[r + 1];
    if (!std::binary_search(b, e, (int32_t)r)) ++nnz;
  }
  return nnz;
}

extern "C" int gcg_ahat_build_host(int64_t n, const int32_t* indptr, const int32_t* indices,
                                   const double* weights, int32_t* out_indptr, int32_t* out_indices,
                                   float* out_vals) {
  GCG_CHECK_ARG(indptr && out_indptr && out_indices && out_vals, "gcg_ahat_build_host: NULL argument");
  // pass 1: pattern with unit diagonal (adj.setdiag(1), tensormain.py:172) + row sums (:174)
  std::vector<double> w64;  // weights of the output pattern
  std::vector<double> dinv(n);
  int64_t o = 0;
  out_indptr[0] = 0;
  w64.reserve((size_t)indptr[n] + n);
  for (int64_t r = 0; r < n; ++r) {
    bool placed = false;
    double sum = 0.0;
    for (int32_t k = indptr[r]; k < indptr[r + 1]; ++k) {
      const int32_t c = indices[k];
      if (k > indptr[r] && indices[k - 1] >= c) {
        set_error("gcg_ahat_build_host: columns of row %lld are not strictly increasing", (long long)r);
        return GCG_ERR_SHAPE;
      }
      double w = weights ? weights[k] : 1.0;
      if (!placed && c >= r) {
        if (c > r) { out_indices[o++] = (int32_t)r; w64.push_back(1.0); sum += 1.0; }
        else w = 1.0;  // setdiag overwrites an existing diagonal value
        placed = true;
      }
      out_indices[o++] = c;
      w64.push_back(w);
      sum += w;
    }
    if (!placed) { out_indices[o++] = (int32_t)r; w64.push_back(1.0); sum += 1.0; }
    out_indptr[r + 1] = (int32_t)o;
    double di = 1.0 / std::sqrt(sum);                              // :175-176
    if (std::isinf(di)) di = 0.0;                                  // :177
    dinv[r] = di;
  }
  // pass 2: H = D * adj * D (:179) evaluated as (d_i * a_ij) * d_j in float64, cast (:180,:221)
  for (int64_t r = 0; r < n; ++r)
    for (int32_t k = out_indptr[r]; k < out_indptr[r + 1]; ++k)
      out_vals[k] = (float)((dinv[r] * w64[k]) * dinv[out_indices[k]]);
  return GCG_OK;
}

// -------------------------------------------------------------- CSR utilities
extern "C" int gcg_csr_transpose_host(int64_t n_rows, int64_t n_cols, const int32_t* indptr,
                                      const int32_t* indices, const float* vals, int32_t* t_indptr,
                                      int32_t* t_indices, float* t_vals) {
  GCG_CHECK_ARG(indptr && t_indptr && (indptr[n_rows] == 0 || (indices && vals && t_indices && t_vals)),
                "gcg_csr_transpose_host: NULL argument");
  const int64_t nnz = indptr[n_rows];
  std::fill(t_indptr, t_indptr + n_cols + 1, 0);
  for (int64_t k = 0; k < nnz; ++k) {
    const int32_t c = indices[k];
    if (c < 0 || c >= n_cols) {
      set_error("gcg_csr_transpose_host: column %d out of range", c);
      return GCG_ERR_SHAPE;
    }
    ++t_indptr[c + 1];
  }
  for (int64_t c = 0; c < n_cols; ++c) t_indptr[c + 1] += t_indptr[c];
  std::vector<int32_t> cursor(t_indptr, t_indptr + n_cols);
  for (int64_t r = 0; r < n_rows; ++r)
    for (int32_t k = indptr[r]; k < indptr[r + 1]; ++k) {
      const int32_t dst = cursor[indices[k]]++;
      t_indices[dst] = (int32_t)r;
      t_vals[dst] = vals[k];
    }
  return GCG_OK;
}

extern "C" int64_t gcg_csr_gather_rows_host(int64_t n_rows, const int32_t* indptr,
                                            const int32_t* indices, const float* vals,
                                            const int32_t* idx, int64_t n_idx, int32_t* out_indptr,
                                            int32_t* out_indices, float* out_vals) {
  if (!indptr || (n_idx > 0 && !idx)) { set_error("gcg_csr_gather_rows_host: NULL argument"); return -1; }
  int64_t nnz = 0;
  for (int64_t i = 0; i < n_idx; ++i) {
    const int64_t r = idx[i];
    if (r < 0 || r >= n_rows) { set_error("gcg_csr_gather_rows_host: index %lld out of range", (long long)r); return -1; }
    nnz += indptr[r + 1] - indptr[r];
  }
  if (!out_indices) return nnz;
  if (nnz >= INT32_MAX) { set_error("gcg_csr_gather_rows_host: nnz overflows int32"); return -1; }
  int64_t o = 0;
  out_indptr[0] = 0;
  for (int64_t i = 0; i < n_idx; ++i) {
    const int64_t r = idx[i];
    const int32_t len = indptr[r + 1] - indptr[r];
    std::memcpy(out_indices + o, indices + indptr[r], sizeof(int32_t) * len);
    std::memcpy(out_vals + o, vals + indptr[r], sizeof(float) * len);
    o += len;
    out_indptr[i + 1] = (int32_t)o;
  }
  return nnz;
}

extern "C" int gcg_csr_permute_host(int64_t n_rows, const int32_t* indptr, const int32_t* indices,
                                    const float* vals, const int32_t* order, const int32_t* col_map,
                                    int32_t* out_indptr, int32_t* out_indices, float* out_vals) {
  GCG_CHECK_ARG(indptr && order && out_indptr && (indptr[n_rows] == 0 || (indices && vals && out_indices && out_vals)),
                "gcg_csr_permute_host: NULL argument");
  out_indptr[0] = 0;
  for (int64_t i = 0; i < n_rows; ++i) {
    const int64_t r = order[i];
    if (r < 0 || r >= n_rows) { set_error("gcg_csr_permute_host: order[%lld] out of range", (long long)i); return GCG_ERR_SHAPE; }
    out_indptr[i + 1] = out_indptr[i] + (indptr[r + 1] - indptr[r]);
  }
#pragma omp parallel
  {
    std::vector<std::pair<int32_t, float>> tmp;
#pragma omp for schedule(dynamic, 1024)
    for (int64_t i = 0; i < n_rows; ++i) {
      const int64_t r = order[i];
      const int32_t b = indptr[r], len = indptr[r + 1] - b;
      int32_t* oi = out_indices + out_indptr[i];
      float* ov = out_vals + out_indptr[i];
      if (!col_map) {
        std::memcpy(oi, indices + b, sizeof(int32_t) * len);
        std::memcpy(ov, vals + b, sizeof(float) * len);
        continue;
      }
      tmp.resize(len);
      for (int32_t k = 0; k < len; ++k) tmp[k] = {col_map[indices[b + k]], vals[b + k]};
      std::sort(tmp.begin(), tmp.end(), [](const std::pair<int32_t, float>& x, const std::pair<int32_t, float>& y) { return x.first < y.first; });
      for (int32_t k = 0; k < len; ++k) { oi[k] = tmp[k].first; ov[k] = tmp[k].second; }
    }
  }
  return GCG_OK;
}

// Column-blocked view of selected rows of a CSR matrix: block b keeps the entries whose column lies in
// [b*block_cols, (b+1)*block_cols).  Used for X^T.dZ (Dot.grad of lasagne_layers.py:65): the rows of the
// frequent vocabulary terms are processed one DOCUMENT block at a time, so the gathered block of dZ rows
// stays L2-resident.  Outputs: out_indptr [n_blocks][n_sel+1] (offsets relative to the block's slice),
// out_block_off [n_blocks+1] (slice boundaries in out_indices/out_vals), entries grouped by block, row.
extern "C" int gcg_csr_split_colblocks_host(const int32_t* indptr, const int32_t* indices, const float* vals,
                                            const int32_t* row_sel, int64_t n_sel, int64_t block_cols,
                                            int64_t n_blocks, int32_t* out_indptr, int64_t* out_block_off,
                                            int32_t* out_indices, float* out_vals) {
  GCG_CHECK_ARG(indptr && indices && vals && row_sel && out_indptr && out_block_off && block_cols > 0 && n_blocks > 0,
                "gcg_csr_split_colblocks_host: bad argument");
  const int64_t stride = n_sel + 1;
  std::vector<int64_t> cnt((size_t)n_blocks * n_sel, 0);
  for (int64_t i = 0; i < n_sel; ++i) {
    const int64_t r = row_sel[i];
    for (int32_t k = indptr[r]; k < indptr[r + 1]; ++k) {
      const int64_t b = indices[k] / block_cols;
      if (b >= n_blocks) { set_error("gcg_csr_split_colblocks_host: column %d beyond the last block", indices[k]); return GCG_ERR_SHAPE; }
      ++cnt[(size_t)b * n_sel + i];
    }
  }
  int64_t total = 0;
  for (int64_t b = 0; b < n_blocks; ++b) {
    out_block_off[b] = total;
    int64_t run = 0;
    for (int64_t i = 0; i < n_sel; ++i) {
      out_indptr[b * stride + i] = (int32_t)run;
      run += cnt[(size_t)b * n_sel + i];
    }
    out_indptr[b * stride + n_sel] = (int32_t)run;
    total += run;
  }
  out_block_off[n_blocks] = total;
  if (!out_indices || !out_vals) return GCG_OK;          // sizing call
  std::vector<int64_t> cur((size_t)n_blocks * n_sel);
  for (int64_t b = 0; b < n_blocks; ++b)
    for (int64_t i = 0; i < n_sel; ++i) cur[(size_t)b * n_sel + i] = out_block_off[b] + out_indptr[b * stride + i];
  for (int64_t i = 0; i < n_sel; ++i) {
    const int64_t r = row_sel[i];
    for (int32_t k = indptr[r]; k < indptr[r + 1]; ++k) {
      const int64_t b = indices[k] / block_cols;
      const int64_t dst = cur[(size_t)b * n_sel + i]++;
      out_indices[dst] = indices[k];
      out_vals[dst] = vals[k];
    }
  }
  return GCG_OK;
}
