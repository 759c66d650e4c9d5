// Host-side maximum-weight full bipartite matching of a block-diagonal sparse score table (loss side of the bipartite
// classifier). The reference calls scipy.sparse.csgraph.min_weight_full_bipartite_matching(table, maximize=True)
// on the host once per event (bipartite_classification_base.py:173); for a collated batch of events the table is block
// diagonal (a supernode never spans events) and the blocks are independent assignment problems: they are solved here side by
// side on host threads (scipy holds the GIL for the whole call; matched as ONE table the cost grows faster than linearly).
//
// Per block: successive shortest augmenting paths with dual potentials (Jonker-Volgenant family, sparse rows, binary heap):
//   minimise sum of c(i, j) = -score(i, j) over matchings that cover every row; reduced costs c - u[i] - v[j] stay >= 0,
//   each row is inserted by one Dijkstra over alternating paths, potentials are moved by (delta - distance) on the scanned
//   part of the tree. float64 arithmetic on the fp32 scores. The optimum is unique up to ties between equal-weight matchings.
// No device code: this file is compiled into the library for the host side of the C ABI.
#include <algorithm>
#include <cstdint>
#include <limits>
#include <queue>
#include <thread>
#include <utility>
#include <vector>

#include "common.cuh"

using namespace hgnn;

namespace {

// rows [r0, r1) of the CSR table; writes col_of_row[r0..r1) (global column ids). Returns false when a row cannot be matched.
bool solve_block(const int32_t* indptr, const int32_t* indices, const float* data, int64_t r0, int64_t r1, int64_t* col_of_row) {
  const int64_t n = r1 - r0;
  if (n <= 0) return true;
  const int64_t lo = indptr[r0], hi = indptr[r1];
  // local column numbering: the sorted distinct column ids this block touches
  std::vector<int32_t> cols(indices + lo, indices + hi);
  std::sort(cols.begin(), cols.end());
  cols.erase(std::unique(cols.begin(), cols.end()), cols.end());
  const int64_t m = (int64_t)cols.size();
  if (m < n) return false;
  std::vector<int32_t> lcol(hi - lo);
  for (int64_t e = lo; e < hi; ++e)
    lcol[e - lo] = (int32_t)(std::lower_bound(cols.begin(), cols.end(), indices[e]) - cols.begin());

  const double INF = std::numeric_limits<double>::infinity();
  std::vector<double> u(n, 0.0), v(m, 0.0), dist(m, INF);
  std::vector<int32_t> row_of_col(m, -1), col_of(n, -1), pred(m, -1);
  std::vector<char> done(m, 0);
  std::vector<int32_t> touched, scanned;
  touched.reserve(256);
  scanned.reserve(256);
  auto cost = [&](int64_t e) { return -(double)data[e]; };

  // start: u[i] = cheapest edge of the row (all reduced costs >= 0), rows take that column when it is still free
  for (int64_t i = 0; i < n; ++i) {
    const int64_t a = indptr[r0 + i], b = indptr[r0 + i + 1];
    if (a == b) return false;
    double best = INF;
    int32_t bj = -1;
    for (int64_t e = a; e < b; ++e) {
      const double c = cost(e);
      if (c < best) { best = c; bj = lcol[e - lo]; }
    }
    u[i] = best;
    if (row_of_col[bj] < 0) { row_of_col[bj] = (int32_t)i; col_of[i] = bj; }
  }

  using Item = std::pair<double, int32_t>;  // (distance, local column)
  for (int64_t s = 0; s < n; ++s) {
    if (col_of[s] >= 0) continue;
    std::priority_queue<Item, std::vector<Item>, std::greater<Item>> heap;
    touched.clear();
    scanned.clear();
    int32_t i = (int32_t)s, sink = -1;
    double base = 0.0, delta = 0.0;
    while (true) {
      const int64_t a = indptr[r0 + i], b = indptr[r0 + i + 1];
      for (int64_t e = a; e < b; ++e) {
        const int32_t j = lcol[e - lo];
        if (done[j]) continue;
        const double nd = base + (cost(e) - u[i] - v[j]);
        if (nd < dist[j]) {
          if (dist[j] == INF) touched.push_back(j);
          dist[j] = nd;
          pred[j] = i;
          heap.push(Item(nd, j));
        }
      }
      int32_t j = -1;
      double d = 0.0;
      while (!heap.empty()) {
        const Item top = heap.top();
        heap.pop();
        if (!done[top.second] && top.first <= dist[top.second]) { j = top.second; d = top.first; break; }
      }
      if (j < 0) break;  // no augmenting path: the block has no full matching
      done[j] = 1;
      scanned.push_back(j);
      if (row_of_col[j] < 0) { sink = j; delta = d; break; }
      i = row_of_col[j];
      base = d;
    }
    if (sink >= 0) {
      // potentials: rows of the tree gain (delta - their distance), scanned columns lose the same amount
      u[s] += delta;
      for (int32_t j : scanned) {
        const double gain = delta - dist[j];
        if (row_of_col[j] >= 0) u[row_of_col[j]] += gain;
        v[j] -= gain;
      }
      // augment along the predecessor chain
      int32_t j = sink;
      while (true) {
        const int32_t r = pred[j];
        const int32_t prev = col_of[r];
        row_of_col[j] = r;
        col_of[r] = j;
        if (r == (int32_t)s) break;
        j = prev;
      }
    }
    for (int32_t j : touched) { dist[j] = INF; done[j] = 0; pred[j] = -1; }
    if (sink < 0) return false;
  }
  for (int64_t i = 0; i < n; ++i) col_of_row[r0 + i] = cols[col_of[i]];
  return true;
}

}  // namespace

extern "C" int hgnn_match_blocks_max(const int32_t* indptr, const int32_t* indices, const float* data, int64_t n_rows,
                                     const int64_t* row_ptr, int64_t n_blocks, int64_t* col_of_row, int n_threads) {
  HGNN_REQUIRE(indptr && row_ptr && col_of_row && n_rows >= 0 && n_blocks >= 0, "match_blocks_max: bad argument");
  HGNN_REQUIRE(n_rows == 0 || (indices && data), "match_blocks_max: NULL table");
  for (int64_t b = 0; b < n_blocks; ++b)
    HGNN_REQUIRE(row_ptr[b] <= row_ptr[b + 1] && row_ptr[b] >= 0 && row_ptr[b + 1] <= n_rows, "match_blocks_max: row_ptr must ascend within [0, n_rows]");
  if (n_blocks == 0) return HGNN_OK;
  HGNN_REQUIRE(row_ptr[0] == 0 && row_ptr[n_blocks] == n_rows, "match_blocks_max: the blocks must cover every row");
  int nt = n_threads > 0 ? n_threads : (int)std::thread::hardware_concurrency();
  nt = (int)std::max<int64_t>(1, std::min<int64_t>(nt, n_blocks));
  std::vector<char> ok(n_blocks, 1);
  auto work = [&](int t) {
    for (int64_t b = t; b < n_blocks; b += nt) ok[b] = solve_block(indptr, indices, data, row_ptr[b], row_ptr[b + 1], col_of_row) ? 1 : 0;
  };
  if (nt == 1) {
    work(0);
  } else {
    std::vector<std::thread> th;
    for (int t = 1; t < nt; ++t) th.emplace_back(work, t);
    work(0);
    for (auto& x : th) x.join();
  }
  for (int64_t b = 0; b < n_blocks; ++b)
    if (!ok[b]) return fail(HGNN_ERR_BAD_ARG, "match_blocks_max: block %lld has no matching that covers every row", (long long)b);
  return HGNN_OK;
}
