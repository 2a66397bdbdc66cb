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
