// Host CSR packer: HNSW neighbour lists -> row_ptr / col_idx / edge_w.
//
// Ground truth is hnsw_rs's Point::get_neighborhood_id()
// (patches/hnsw_rs/src/hnsw.rs:220-240): per point, per layer, a Vec of
// #[repr(C)] Neighbour {d_id: usize, distance: f32, p_id: PointId} (:97-106),
// stored in ascending distance (:1186).  The reference has no HNSW->GNN bridge
// (SURVEY.md fact 2); this is the new code a ruvector-gnn CUDA backend calls
// once per HNSW level.  Integer/byte work only: copies ids and distances
// verbatim, preserves order, never infers emptiness from p_id.
#include <string.h>

#include <string>

#include "../../include/ruvgnn.h"
#include "abi_guard.h"

namespace ruvgnn {
int32_t fail(int32_t code, const std::string &msg);
}

extern "C" int32_t ruvgnn_pack_neighbours(uint64_t n_nodes, const void *const *lists,
                                          const uint64_t *counts, size_t stride, size_t id_offset,
                                          size_t dist_offset, uint64_t *row_ptr, uint32_t *col_idx,
                                          float *edge_w) {
  RUVGNN_ABI_TRY
  if (!counts || !row_ptr || (n_nodes && !lists))
    return ruvgnn::fail(RUVGNN_ERR_INVALID_INPUT, "null argument");
  if (stride < sizeof(uint64_t) + sizeof(float) || id_offset + sizeof(uint64_t) > stride ||
      dist_offset + sizeof(float) > stride)
    return ruvgnn::fail(RUVGNN_ERR_INVALID_INPUT, "record stride/offsets inconsistent");
  uint64_t e = 0;
  row_ptr[0] = 0;
  for (uint64_t i = 0; i < n_nodes; ++i) {
    const unsigned char *rec = static_cast<const unsigned char *>(lists[i]);
    const uint64_t c = counts[i];
    if (c && !rec)
      return ruvgnn::fail(RUVGNN_ERR_INVALID_INPUT,
                          "node " + std::to_string(i) + " has a count but no list");
    if (c && (!col_idx || !edge_w)) return ruvgnn::fail(RUVGNN_ERR_INVALID_INPUT, "null output");
    for (uint64_t r = 0; r < c; ++r, rec += stride) {
      uint64_t id;
      float d;
      memcpy(&id, rec + id_offset, sizeof id);
      memcpy(&d, rec + dist_offset, sizeof d);
      if (id >= n_nodes || id > 0xffffffffull)
        return ruvgnn::fail(RUVGNN_ERR_INVALID_INPUT,
                            "neighbour id " + std::to_string(id) + " of node " +
                                std::to_string(i) + " out of range");
      col_idx[e] = (uint32_t)id;
      edge_w[e] = d;
      ++e;
    }
    row_ptr[i + 1] = e;
  }
  return RUVGNN_OK;
  RUVGNN_ABI_CATCH
}

// Union-of-levels view of the adjacency, mirroring hnsw_rs flatten_point
// (patches/hnsw_rs/src/flatten.rs:74-89): node i's lists of every level are concatenated (level 0
// first) and sorted by distance (Ord for Neighbour compares the distance only, :29-43).  The
// reference uses sort_unstable, so the order of equal distances is unspecified there; here the
// sort is stable (equal distances keep the level order), which is one of the orders the reference
// may produce.  A NaN distance panics in the reference (:37-41) and is RUVGNN_ERR_INVALID_INPUT here.
// Duplicates (the same neighbour on several levels) are kept, as the reference keeps them.
#include <algorithm>
#include <cmath>
#include <vector>

extern "C" int32_t ruvgnn_flatten_levels(uint64_t n_nodes, uint32_t n_levels,
                                         const uint64_t *const *row_ptrs,
                                         const uint32_t *const *col_idxs,
                                         const float *const *edge_ws, uint64_t *row_ptr,
                                         uint32_t *col_idx, float *edge_w) {
  RUVGNN_ABI_TRY
  if (!row_ptr || (n_levels && (!row_ptrs || !col_idxs || !edge_ws)))
    return ruvgnn::fail(RUVGNN_ERR_INVALID_INPUT, "null argument");
  for (uint32_t l = 0; l < n_levels; ++l)
    if (!row_ptrs[l]) return ruvgnn::fail(RUVGNN_ERR_INVALID_INPUT, "null level row_ptr");
  uint64_t e = 0;
  row_ptr[0] = 0;
  std::vector<std::pair<float, uint32_t>> row;
  for (uint64_t i = 0; i < n_nodes; ++i) {
    row.clear();
    for (uint32_t l = 0; l < n_levels; ++l) {
      const uint64_t b = row_ptrs[l][i], en = row_ptrs[l][i + 1];
      if (en < b) return ruvgnn::fail(RUVGNN_ERR_INVALID_INPUT, "level row_ptr is not monotone");
      if (en > b && (!col_idxs[l] || !edge_ws[l]))
        return ruvgnn::fail(RUVGNN_ERR_INVALID_INPUT, "null level col_idx / edge_w");
      for (uint64_t k = b; k < en; ++k) {
        if (std::isnan(edge_ws[l][k]))
          return ruvgnn::fail(RUVGNN_ERR_INVALID_INPUT, "got a NaN in a distance (node " + std::to_string(i) + ")");
        row.push_back({edge_ws[l][k], col_idxs[l][k]});
      }
    }
    std::stable_sort(row.begin(), row.end(),
                     [](const std::pair<float, uint32_t> &a, const std::pair<float, uint32_t> &b) {
                       return a.first < b.first;
                     });
    if (!row.empty() && (!col_idx || !edge_w)) return ruvgnn::fail(RUVGNN_ERR_INVALID_INPUT, "null output");
    for (const auto &r : row) {
      col_idx[e] = r.second;
      edge_w[e] = r.first;
      ++e;
    }
    row_ptr[i + 1] = e;
  }
  return RUVGNN_OK;
  RUVGNN_ABI_CATCH
}

// BFS vertex order for locality, the reference's own reorder
// (DataLoader::reorder_bfs, crates/ruvector-gnn/src/cold_tier.rs:307-342) on a CSR: the edge list
// is the CSR in stored order, (u, v) = (row, column); every edge is pushed to both endpoints'
// adjacency lists in that order (:313-318), so node x's list is [rows u < x that point at x, in
// row order] + [x's own neighbours in stored order] + [rows u > x that point at x].  BFS from node
// 0, restarting at the lowest unvisited id for further components (:325-339).  O(N + E) time and
// memory: usable at 1e8 nodes, where the SciPy RCM of graph.locality_order is not.
// order[p] = the node that moves to position p (a permutation of 0..n-1).
extern "C" int32_t ruvgnn_bfs_order(uint64_t n_nodes, const uint64_t *row_ptr,
                                    const uint32_t *col_idx, uint64_t *order) {
  RUVGNN_ABI_TRY
  if (n_nodes == 0) return RUVGNN_OK;
  if (!row_ptr || !order) return ruvgnn::fail(RUVGNN_ERR_INVALID_INPUT, "null argument");
  if (n_nodes > 0xffffffffull) return ruvgnn::fail(RUVGNN_ERR_INVALID_INPUT, "node ids exceed u32");
  const uint64_t E = row_ptr[n_nodes];
  if (E && !col_idx) return ruvgnn::fail(RUVGNN_ERR_INVALID_INPUT, "col_idx is null");
  // transposed adjacency (sources of every node's in-edges, in edge order = ascending source row)
  std::vector<uint64_t> in_ptr(n_nodes + 1, 0);
  for (uint64_t u = 0; u < n_nodes; ++u) {
    if (row_ptr[u + 1] < row_ptr[u]) return ruvgnn::fail(RUVGNN_ERR_INVALID_INPUT, "row_ptr is not monotone");
    for (uint64_t k = row_ptr[u]; k < row_ptr[u + 1]; ++k) {
      if (col_idx[k] >= n_nodes) continue;  // the reference skips out-of-range endpoints (:314)
      ++in_ptr[(uint64_t)col_idx[k] + 1];
    }
  }
  for (uint64_t i = 0; i < n_nodes; ++i) in_ptr[i + 1] += in_ptr[i];
  std::vector<uint32_t> in_src(in_ptr[n_nodes]);
  {
    std::vector<uint64_t> at(in_ptr.begin(), in_ptr.end() - 1);
    for (uint64_t u = 0; u < n_nodes; ++u)
      for (uint64_t k = row_ptr[u]; k < row_ptr[u + 1]; ++k)
        if (col_idx[k] < n_nodes) in_src[at[col_idx[k]]++] = (uint32_t)u;
  }
  std::vector<uint64_t> visited((n_nodes + 63) / 64, 0);
  auto seen = [&](uint64_t v) { return (visited[v >> 6] >> (v & 63)) & 1ull; };
  auto mark = [&](uint64_t v) { visited[v >> 6] |= 1ull << (v & 63); };
  uint64_t head = 0, tail = 0;  // `order` doubles as the BFS queue
  auto visit = [&](uint64_t v) {
    if (!seen(v)) {
      mark(v);
      order[tail++] = v;
    }
  };
  for (uint64_t start = 0; start < n_nodes; ++start) {
    if (seen(start)) continue;
    visit(start);
    while (head < tail) {
      const uint64_t x = order[head++];
      uint64_t k = in_ptr[x];
      const uint64_t ke = in_ptr[x + 1];
      for (; k < ke && in_src[k] < x; ++k) visit(in_src[k]);  // pushed by rows before x
      // a self loop (x, x) pushes x to its own list twice, in row x's position: already visited
      for (; k < ke && in_src[k] == x; ++k) {
      }
      for (uint64_t j = row_ptr[x]; j < row_ptr[x + 1]; ++j)
        if (col_idx[j] < n_nodes) visit(col_idx[j]);
      for (; k < ke; ++k) visit(in_src[k]);  // pushed by rows after x
    }
  }
  return RUVGNN_OK;
  RUVGNN_ABI_CATCH
}
