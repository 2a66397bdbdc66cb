// Reader for hnsw_rs graph dumps (`<basename>.hnsw.graph`) -> per-level CSR.
//
// Format as written by the reference (all integers native-endian, usize = 8 bytes):
//   Description::dump            patches/hnsw_rs/src/hnswio.rs:878-918 (magics :49-65)
//     u32 MAGICDESCR_{2,3,4}; u8 dumpmode (1 Full, 0 Light); u8 max_nb_connection;
//     [v4: f64 level_scale]; u8 nb_layer; usize ef; usize nb_point; usize dimension;
//     usize len + distance name; usize len + T type name
//   PointIndexation::dump        hnswio.rs:1303-1340
//     u8 nb_layer; per layer: u32 MAGICLAYER, usize nb_point, points; then the entry point id
//   dump_point                   hnswio.rs:1063-1096
//     u32 MAGICPOINT; usize origin_id; [Full: u8 layer, i32 rank];
//     per neighbour layer (Description.nb_layer of them): usize count, then per neighbour
//     usize d_id, [Full: u8, i32], f32 distance
// Neighbour lists are kept exactly as stored (order = ascending distance, hnsw.rs:1186);
// the CSR row of a point is its origin_id (= ruvector-core HnswIndex idx, index/hnsw.rs:273-277).
// Host-only integer/byte work; no device involved.
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include <memory>
#include <string>
#include <vector>

#include "../../include/ruvgnn.h"
#include "abi_guard.h"

namespace ruvgnn {
int32_t fail(int32_t code, const std::string &msg);
}
using ruvgnn::fail;

namespace {
constexpr uint32_t MAGICPOINT = 0x000a678f, MAGICDESCR_2 = 0x002a677f, MAGICDESCR_3 = 0x002a6771,
                   MAGICDESCR_4 = 0x002a6779, MAGICLAYER = 0x000a676f;

struct Reader {
  const uint8_t *p, *end;
  bool ok = true;
  template <class T>
  T get() {
    T v{};
    if ((size_t)(end - p) < sizeof(T)) {
      ok = false;
      p = end;
      return v;
    }
    memcpy(&v, p, sizeof(T));
    p += sizeof(T);
    return v;
  }
  bool skip(size_t n) {
    if ((size_t)(end - p) < n) {
      ok = false;
      p = end;
      return false;
    }
    p += n;
    return true;
  }
};
}  // namespace

struct ruvgnn_hnsw_dump {
  uint32_t format_version = 0, dumpmode = 0, max_nb_connection = 0, nb_layer = 0;
  uint64_t ef = 0, nb_point = 0, dimension = 0, entry_point = 0, n_ids = 0;
  double level_scale = 1.0;
  std::string distname, t_name;
  // one record per point, in file order
  std::vector<uint64_t> origin;       // [nb_point]
  std::vector<uint8_t> point_layer;   // [nb_point] (Full mode; else the layer block index)
  std::vector<uint64_t> list_off;     // [nb_point * nb_layer + 1] offsets into ids / dist
  std::vector<uint64_t> ids;
  std::vector<float> dist;
  std::vector<uint64_t> level_edges;  // [nb_layer]
};

extern "C" {

int32_t ruvgnn_hnsw_dump_open(const char *graph_path, ruvgnn_hnsw_dump **out) {
  RUVGNN_ABI_TRY
  if (!graph_path || !out) return fail(RUVGNN_ERR_INVALID_INPUT, "null argument");
  *out = nullptr;
  FILE *f = fopen(graph_path, "rb");
  if (!f) return fail(RUVGNN_ERR_INVALID_INPUT, std::string("cannot open ") + graph_path);
  fseek(f, 0, SEEK_END);
  long sz = ftell(f);
  fseek(f, 0, SEEK_SET);
  std::vector<uint8_t> buf((size_t)(sz > 0 ? sz : 0));
  if (sz > 0 && fread(buf.data(), 1, (size_t)sz, f) != (size_t)sz) {
    fclose(f);
    return fail(RUVGNN_ERR_INVALID_INPUT, "short read");
  }
  fclose(f);
  Reader r{buf.data(), buf.data() + buf.size()};
  std::unique_ptr<ruvgnn_hnsw_dump> owner(new ruvgnn_hnsw_dump());
  ruvgnn_hnsw_dump *d = owner.get();
  auto bad = [&](const std::string &m) { return fail(RUVGNN_ERR_INVALID_INPUT, "hnsw dump: " + m); };
  const uint32_t magic = r.get<uint32_t>();
  if (magic == MAGICDESCR_2) d->format_version = 2;
  else if (magic == MAGICDESCR_3) d->format_version = 3;
  else if (magic == MAGICDESCR_4) d->format_version = 4;
  else return bad("bad magic at descr beginning");
  d->dumpmode = r.get<uint8_t>();
  d->max_nb_connection = r.get<uint8_t>();
  if (d->format_version == 4) d->level_scale = r.get<double>();
  d->nb_layer = r.get<uint8_t>();
  d->ef = r.get<uint64_t>();
  d->nb_point = r.get<uint64_t>();
  d->dimension = r.get<uint64_t>();
  for (std::string *s : {&d->distname, &d->t_name}) {
    const uint64_t len = r.get<uint64_t>();
    if (!r.ok || len > 256) return bad("bad name length");
    if ((uint64_t)(r.end - r.p) < len) return bad("truncated name");
    s->assign(reinterpret_cast<const char *>(r.p), (size_t)len);
    r.skip((size_t)len);
  }
  if (!r.ok) return bad("truncated description");
  if (d->nb_layer == 0 || d->nb_layer > 64) return bad("implausible nb_layer");
  const bool full = d->dumpmode == 1;
  const uint32_t L = d->nb_layer;
  d->level_edges.assign(L, 0);
  d->origin.reserve(d->nb_point);
  d->list_off.reserve(d->nb_point * L + 1);
  d->list_off.push_back(0);
  const uint32_t blocks = r.get<uint8_t>();
  for (uint32_t b = 0; b < blocks; ++b) {
    if (r.get<uint32_t>() != MAGICLAYER || !r.ok) return bad("bad magic at layer beginning");
    const uint64_t npts = r.get<uint64_t>();
    if (!r.ok || npts > (uint64_t)(r.end - r.p) / 12) return bad("implausible point count in layer");
    for (uint64_t j = 0; j < npts; ++j) {
      if (r.get<uint32_t>() != MAGICPOINT || !r.ok) return bad("bad magic at point beginning");
      const uint64_t origin = r.get<uint64_t>();
      // rows are addressed by origin_id and columns are u32: anything larger cannot be a node
      // (and origin + 1 below must not wrap)
      if (!r.ok || origin >= 0xffffffffull) return bad("origin_id exceeds the u32 node index range");
      uint8_t layer = (uint8_t)b;
      if (full) {
        layer = r.get<uint8_t>();
        (void)r.get<int32_t>();
      }
      d->origin.push_back(origin);
      d->point_layer.push_back(layer);
      if (origin + 1 > d->n_ids) d->n_ids = origin + 1;
      for (uint32_t l = 0; l < L; ++l) {
        const uint64_t cnt = r.get<uint64_t>();
        const size_t rec = 8 + (full ? 5 : 0) + 4;
        if (!r.ok || cnt > (uint64_t)(r.end - r.p) / rec) return bad("truncated neighbour list");
        for (uint64_t k = 0; k < cnt; ++k) {
          d->ids.push_back(r.get<uint64_t>());
          if (full) r.skip(5);
          d->dist.push_back(r.get<float>());
        }
        d->level_edges[l] += cnt;
        d->list_off.push_back(d->ids.size());
      }
    }
  }
  d->entry_point = r.get<uint64_t>();
  if (!r.ok) return bad("truncated before the entry point");
  if (d->origin.size() != d->nb_point) return bad("point count differs from the description");
  *out = owner.release();
  return RUVGNN_OK;
  RUVGNN_ABI_CATCH
}

void ruvgnn_hnsw_dump_close(ruvgnn_hnsw_dump *d) { delete d; }

int32_t ruvgnn_hnsw_dump_info(const ruvgnn_hnsw_dump *d, uint64_t *nb_point, uint64_t *n_ids,
                              uint32_t *nb_layer, uint32_t *max_nb_connection, uint64_t *dimension,
                              uint32_t *dumpmode, uint64_t *entry_point) {
  RUVGNN_ABI_TRY
  if (!d) return fail(RUVGNN_ERR_INVALID_INPUT, "dump is null");
  if (nb_point) *nb_point = d->nb_point;
  if (n_ids) *n_ids = d->n_ids;
  if (nb_layer) *nb_layer = d->nb_layer;
  if (max_nb_connection) *max_nb_connection = d->max_nb_connection;
  if (dimension) *dimension = d->dimension;
  if (dumpmode) *dumpmode = d->dumpmode;
  if (entry_point) *entry_point = d->entry_point;
  return RUVGNN_OK;
  RUVGNN_ABI_CATCH
}

int32_t ruvgnn_hnsw_dump_level_edges(const ruvgnn_hnsw_dump *d, uint32_t level, uint64_t *n_edges) {
  RUVGNN_ABI_TRY
  if (!d || !n_edges) return fail(RUVGNN_ERR_INVALID_INPUT, "null argument");
  if (level >= d->nb_layer) return fail(RUVGNN_ERR_INVALID_INPUT, "level out of range");
  *n_edges = d->level_edges[level];
  return RUVGNN_OK;
  RUVGNN_ABI_CATCH
}

int32_t ruvgnn_hnsw_dump_level_csr(const ruvgnn_hnsw_dump *d, uint32_t level, uint64_t n_nodes,
                                   uint64_t *row_ptr, uint32_t *col_idx, float *edge_w) {
  RUVGNN_ABI_TRY
  if (!d || !row_ptr) return fail(RUVGNN_ERR_INVALID_INPUT, "null argument");
  if (level >= d->nb_layer) return fail(RUVGNN_ERR_INVALID_INPUT, "level out of range");
  if (n_nodes < d->n_ids)
    return fail(RUVGNN_ERR_INVALID_INPUT, "n_nodes is smaller than the largest origin_id + 1");
  if (n_nodes > 0xffffffffull) return fail(RUVGNN_ERR_INVALID_INPUT, "n_nodes exceeds u32 col_idx");
  const uint32_t L = d->nb_layer;
  // degree per origin id
  for (uint64_t i = 0; i <= n_nodes; ++i) row_ptr[i] = 0;
  std::vector<uint8_t> seen((size_t)n_nodes, 0);
  for (uint64_t p = 0; p < d->nb_point; ++p) {
    const uint64_t cnt = d->list_off[p * L + level + 1] - d->list_off[p * L + level];
    if (seen[d->origin[p]]) return fail(RUVGNN_ERR_INVALID_INPUT, "duplicate origin_id in dump");
    seen[d->origin[p]] = 1;
    row_ptr[d->origin[p] + 1] = cnt;
  }
  for (uint64_t i = 0; i < n_nodes; ++i) row_ptr[i + 1] += row_ptr[i];
  if (row_ptr[n_nodes] && (!col_idx || !edge_w))
    return fail(RUVGNN_ERR_INVALID_INPUT, "null output");
  for (uint64_t p = 0; p < d->nb_point; ++p) {
    const uint64_t b = d->list_off[p * L + level], e = d->list_off[p * L + level + 1];
    uint64_t o = row_ptr[d->origin[p]];
    for (uint64_t k = b; k < e; ++k, ++o) {
      if (d->ids[k] >= n_nodes) return fail(RUVGNN_ERR_INVALID_INPUT, "neighbour id out of range");
      col_idx[o] = (uint32_t)d->ids[k];
      edge_w[o] = d->dist[k];
    }
  }
  return RUVGNN_OK;
  RUVGNN_ABI_CATCH
}

}  // extern "C"
