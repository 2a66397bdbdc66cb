/*
 * ruvgnn.h -- C ABI of libruvgnn.so: the B200 (sm_100a) backend for
 * ruvector-gnn's RuvectorLayer forward over an HNSW neighbourhood graph.
 *
 * This is the boundary a `cuda` sub-module of ruvector_gnn::layer would bind
 * (INTEGRATION.md shows the Rust `extern "C"` block).  Plain pointers and
 * sizes only; no C++/torch types.  All reference citations are relative to
 * /root/reference/ (Mu-L/ruvector 2.0.5).
 *
 * Conventions (pattern of crates/ruvector-fpga-transformer/src/ffi/c_abi.rs:20-76):
 *   - every call returns an int32 status, 0 = OK; never panics/aborts;
 *   - opaque handles are created and destroyed by paired functions;
 *   - the library copies what it needs and never frees caller memory;
 *   - output buffers are caller-allocated;
 *   - there is NO CPU fallback: without a CUDA device every compute entry
 *     returns RUVGNN_ERR_CUDA.
 */
#ifndef RUVGNN_H
#define RUVGNN_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* Status codes.  1..7 mirror GnnError's variants (crates/ruvector-gnn/src/error.rs:9-60). */
#define RUVGNN_OK 0
#define RUVGNN_ERR_DIMENSION_MISMATCH 1 /* GnnError::DimensionMismatch */
#define RUVGNN_ERR_INVALID_SHAPE 2      /* GnnError::InvalidShape */
#define RUVGNN_ERR_LAYER_CONFIG 3       /* GnnError::LayerConfig (layer.rs:101-106,353-358) */
#define RUVGNN_ERR_INVALID_INPUT 7      /* GnnError::InvalidInput */
#define RUVGNN_ERR_CUDA 100
#define RUVGNN_ERR_NCCL 101
#define RUVGNN_ERR_OOM 102

typedef struct ruvgnn_ctx ruvgnn_ctx;     /* device, stream, workspace */
typedef struct ruvgnn_layer ruvgnn_layer; /* one RuvectorLayer's weights on the device */
typedef struct ruvgnn_graph ruvgnn_graph; /* device CSR of one HNSW level */

/* Message of the last failing call on the calling thread (thread-local copy;
 * for RUVGNN_ERR_LAYER_CONFIG it is the reference's exact text, e.g.
 * "Dropout must be between 0.0 and 1.0, got 1.5"). */
const char *ruvgnn_last_error(void);

/* ABI version; bumps on any signature change. */
uint32_t ruvgnn_abi_version(void);

/* ---- context ----------------------------------------------------------- */

/* `stream` is a cudaStream_t to enqueue on (NULL: the library creates its own).
 * One forward in flight per ctx.  Replaces nothing in the reference (the CPU
 * layer has no device state). */
int32_t ruvgnn_ctx_create(int32_t device, void *stream, ruvgnn_ctx **out);
void ruvgnn_ctx_destroy(ruvgnn_ctx *ctx);
int32_t ruvgnn_sync(ruvgnn_ctx *ctx);
void *ruvgnn_ctx_stream(ruvgnn_ctx *ctx);
/* Number of kernels this ctx has launched since creation (bench.py's gpu_launches). */
uint64_t ruvgnn_ctx_launch_count(ruvgnn_ctx *ctx);
/* Device milliseconds per stage (0 project, 1 aggregate, 2 update, 3 h2d, 4 d2h,
 * 5 halo exchange wait), averaged over the timed forwards (up to 32) since the
 * previous query; CUDA events on the launching stream.  Synchronises the stream.
 * aggregate = sum over the K-agg launches (one per row block / row range / gather pass). */
int32_t ruvgnn_ctx_last_timing(ruvgnn_ctx *ctx, float *ms6);
/* Launch grids sized for at most `sms` SMs (0 = all of the device; negative = all but
 * |sms|): leaves SMs to a collective that runs concurrently on another stream, so persistent
 * kernels do not queue behind its blocks. */
int32_t ruvgnn_ctx_set_sm_limit(ruvgnn_ctx *ctx, int32_t sms);
/* Also time the phased entries (ruvgnn_layer_project / ruvgnn_layer_aggregate_update). */
int32_t ruvgnn_ctx_set_timing(ruvgnn_ctx *ctx, int32_t on);

/* ---- layer: RuvectorLayer::new / serde -------------------------------- */

/* One `Linear` (layer.rs:14-17): weights row-major [out][in] exactly as serde
 * writes them, bias [out]. */
typedef struct {
  const float *w;
  const float *b;
} ruvgnn_linear;

/* Field order = serde declaration order: RuvectorLayer (layer.rs:315-333),
 * GRUCell (:215-227), MultiHeadAttention (:86-93), LayerNorm (:51-55). */
typedef struct {
  ruvgnn_linear w_msg, w_agg;
  ruvgnn_linear w_z, u_z, w_r, u_r, w_h, u_h;
  ruvgnn_linear q, k, v, out;
  const float *gamma;
  const float *beta;
  float eps; /* 1e-5 in RuvectorLayer::new (layer.rs:365) */
} ruvgnn_weights;

/* Arithmetic mode of the dense projections.  The gather/aggregate kernel is
 * always FP32.  A new layer starts in RUVGNN_MODE_FP32_TC when its shape allows it
 * (input_dim % 32 == 0 and hidden_dim % 32 == 0), else in RUVGNN_MODE_FP32_SIMT; there is no
 * silent switch afterwards: ruvgnn_layer_set_mode on an unsupported shape fails. */
#define RUVGNN_MODE_FP32_SIMT 0 /* FP32 FFMA GEMMs (any shape; the parity reference on device) */
#define RUVGNN_MODE_FP32_TC 1   /* tcgen05 3xTF32 error-compensated, FP32 accumulate: FP32-grade */
#define RUVGNN_MODE_TF32 2      /* tcgen05 single-pass TF32: NOT parity, own tolerance */

/* Replaces RuvectorLayer::new(input_dim, hidden_dim, heads, dropout)
 * (layer.rs:347-368) with injected weights (host pointers).  Same validation,
 * same messages: dropout outside [0,1] and hidden_dim % heads != 0
 * -> RUVGNN_ERR_LAYER_CONFIG. */
int32_t ruvgnn_layer_create(ruvgnn_ctx *ctx, uint32_t input_dim, uint32_t hidden_dim,
                            uint32_t heads, float dropout, const ruvgnn_weights *w,
                            ruvgnn_layer **out);
/* Validation only (no device needed): what `new` would return. */
int32_t ruvgnn_layer_validate(uint32_t input_dim, uint32_t hidden_dim, uint32_t heads,
                              float dropout);
/* Replaces serde_json::from_str::<RuvectorLayer> (ruvector-gnn-node/src/lib.rs:108-116):
 * accepts the JSON the reference's to_json() emits. */
int32_t ruvgnn_layer_from_json(ruvgnn_ctx *ctx, const char *json, size_t len, ruvgnn_layer **out);
/* Replaces serde_json::to_string(&layer) (lib.rs:97-105).  Writes up to `cap`
 * bytes (NUL-terminated when it fits) and the needed size (without NUL). */
int32_t ruvgnn_layer_to_json(const ruvgnn_layer *layer, char *buf, size_t cap, size_t *needed);
/* Parses a layer JSON and re-emits it in canonical form (same rules as the two calls
 * above) without touching a device: a validity check for a reference-produced file. */
int32_t ruvgnn_layer_json_normalize(const char *json, size_t len, char *buf, size_t cap,
                                    size_t *needed);
int32_t ruvgnn_layer_dims(const ruvgnn_layer *layer, uint32_t *input_dim, uint32_t *hidden_dim,
                          uint32_t *heads, float *dropout);
int32_t ruvgnn_layer_set_mode(ruvgnn_layer *layer, uint32_t mode);
uint32_t ruvgnn_layer_mode(const ruvgnn_layer *layer);
/* Kernel launches of the update phase (out-proj + combine + w_agg, GRU, dropout scale, LayerNorm:
 * layer.rs:156-159,407-417) per row block in the layer's current mode: 1 when the whole chain runs as the
 * fused kernel (hidden_dim == 128, tensor-core modes; every embedding row is then written once and no
 * intermediate touches HBM), 5 otherwise (four GEMMs + LayerNorm). */
uint32_t ruvgnn_layer_update_launches(const ruvgnn_layer *layer);
void ruvgnn_layer_destroy(ruvgnn_layer *layer);

/* ---- graph: HNSW neighbour lists -> device CSR ------------------------- */

/* Packs a list-of-lists of hnsw_rs `Neighbour` records (#[repr(C)]
 * {d_id: usize, distance: f32, p_id}, patches/hnsw_rs/src/hnsw.rs:97-106) as
 * returned by Point::get_neighborhood_id()[level] (hnsw.rs:220-240) into CSR.
 * `lists[i]` points at node i's records (may be NULL when counts[i]==0);
 * record r of node i lives at lists[i] + r*stride; d_id is read as u64 at
 * +id_offset and distance as f32 at +dist_offset.  Order is preserved
 * (ascending distance, hnsw.rs:1186).  row_ptr: [n_nodes+1] u64; col_idx,
 * edge_w: [sum(counts)].  d_id >= n_nodes or > u32 -> RUVGNN_ERR_INVALID_INPUT.
 * Host only; no device needed. */
int32_t ruvgnn_pack_neighbours(uint64_t n_nodes, const void *const *lists, const uint64_t *counts,
                               size_t stride, size_t id_offset, size_t dist_offset,
                               uint64_t *row_ptr, uint32_t *col_idx, float *edge_w);

/* Union-of-levels adjacency, the CSR analogue of hnsw_rs flatten_point
 * (patches/hnsw_rs/src/flatten.rs:74-89): per node the lists of all n_levels level CSRs are
 * concatenated (level 0 first) and sorted by ascending distance (stable; the reference's
 * sort_unstable leaves equal distances unordered), duplicates kept.  NaN distance ->
 * RUVGNN_ERR_INVALID_INPUT (the reference panics).  row_ptr: [n_nodes+1]; col_idx / edge_w:
 * [sum of the levels' edges].  Host only. */
int32_t ruvgnn_flatten_levels(uint64_t n_nodes, uint32_t n_levels, const uint64_t *const *row_ptrs,
                              const uint32_t *const *col_idxs, const float *const *edge_ws,
                              uint64_t *row_ptr, uint32_t *col_idx, float *edge_w);

/* BFS vertex order for locality before range partitioning: DataLoader::reorder_bfs
 * (crates/ruvector-gnn/src/cold_tier.rs:307-342) with the CSR in stored order as the edge list
 * (u = row, v = column; every edge joins both endpoints' lists).  order[p] = the node that moves
 * to position p.  O(n_nodes + n_edges).  Host only. */
int32_t ruvgnn_bfs_order(uint64_t n_nodes, const uint64_t *row_ptr, const uint32_t *col_idx,
                         uint64_t *order);

/* Reads an hnsw_rs graph dump (`<basename>.hnsw.graph`, written by Hnsw::file_dump:
 * Description patches/hnsw_rs/src/hnswio.rs:878-918, PointIndexation :1303-1340,
 * dump_point :1063-1096; Full and Light modes, description versions 2-4) and exposes each
 * HNSW level as CSR rows indexed by origin_id, neighbours in stored order.  Host only. */
typedef struct ruvgnn_hnsw_dump ruvgnn_hnsw_dump;
int32_t ruvgnn_hnsw_dump_open(const char *graph_path, ruvgnn_hnsw_dump **out);
void ruvgnn_hnsw_dump_close(ruvgnn_hnsw_dump *d);
/* n_ids = largest origin_id + 1 (the minimum n_nodes of the CSR). */
int32_t ruvgnn_hnsw_dump_info(const ruvgnn_hnsw_dump *d, uint64_t *nb_point, uint64_t *n_ids,
                              uint32_t *nb_layer, uint32_t *max_nb_connection, uint64_t *dimension,
                              uint32_t *dumpmode, uint64_t *entry_point);
int32_t ruvgnn_hnsw_dump_level_edges(const ruvgnn_hnsw_dump *d, uint32_t level, uint64_t *n_edges);
/* row_ptr: [n_nodes+1]; col_idx / edge_w: [level_edges]. */
int32_t ruvgnn_hnsw_dump_level_csr(const ruvgnn_hnsw_dump *d, uint32_t level, uint64_t n_nodes,
                                   uint64_t *row_ptr, uint32_t *col_idx, float *edge_w);

/* Uploads a CSR (host arrays; the caller keeps ownership).  Validates
 * row_ptr[0]==0, monotone row_ptr, col_idx < n_cols -> RUVGNN_ERR_INVALID_INPUT.
 * n_cols = number of addressable embedding rows (n_nodes for a single GPU;
 * n_nodes + n_halo for a partition, see the range-partitioned forward below).
 * RUVGNN_GRAPH_PRENORMALIZED: edge_w already holds the normalised weights of
 * layer.rs:426-432 (used by the per-node entry to keep zip truncation). */
#define RUVGNN_GRAPH_PRENORMALIZED 1u
#define RUVGNN_GRAPH_DEVICE_PTRS 2u /* row_ptr / col_idx / edge_w are device pointers (copied) */
/* A graph in which fewer than half of the rows have neighbours (upper HNSW levels: ~n / M^l nodes
 * own a list) also gets the list of those rows; the forward then runs Q projection, gather and
 * update on them alone and LN(W_msg x + b) (layer.rs:385-389) on the rest -- same bits, far less
 * work.  RUVGNN_GRAPH_NO_ACTIVE_LIST keeps the dense path (tests compare the two). */
#define RUVGNN_GRAPH_NO_ACTIVE_LIST 4u
int32_t ruvgnn_graph_create(ruvgnn_ctx *ctx, uint64_t n_nodes, uint64_t n_cols,
                            const uint64_t *row_ptr, const uint32_t *col_idx, const float *edge_w,
                            uint32_t flags, ruvgnn_graph **out);
int32_t ruvgnn_graph_dims(const ruvgnn_graph *g, uint64_t *n_nodes, uint64_t *n_cols,
                          uint64_t *n_edges);
/* Copies the device CSR back (bit-exact round-trip check of the upload). */
int32_t ruvgnn_graph_download(const ruvgnn_graph *g, uint64_t *row_ptr, uint32_t *col_idx,
                              float *edge_w);
void ruvgnn_graph_destroy(ruvgnn_graph *g);

/* ---- forward ----------------------------------------------------------- */

#define RUVGNN_FWD_DEVICE_PTRS 1u /* x / out are device pointers (default: host) */
#define RUVGNN_FWD_ASYNC 2u       /* return after enqueue; ruvgnn_sync() to wait */
#define RUVGNN_FWD_TIMING 4u      /* record per-stage CUDA events */

/* Batched forward: out[i] = layer.forward(x[i], [x[j] for j in N(i)], w[i])
 * for every node, all nodes reading the OLD embeddings -- the loop of
 * verified_training.rs:422-436 / stage_iii.rs:143-168 in one call.
 * x: [n_nodes][input_dim] row-major, out: [n_nodes][hidden_dim]; must not alias. */
int32_t ruvgnn_layer_forward(ruvgnn_ctx *ctx, const ruvgnn_layer *layer, const ruvgnn_graph *graph,
                             const float *x, float *out, uint32_t flags);

/* Stacked forward (BASELINE config 3: one RuvectorLayer per HNSW level, applied in the
 * order given, a full Jacobi barrier between layers -- the batched analogue of the loop
 * in hierarchical_forward, crates/ruvector-gnn/src/search.rs:117-159):
 *   h_0 = x;  h_{l+1}[i] = layers[l].forward(h_l[i], [h_l[j] for j in N_l(i)], w_l[i]);  out = h_L.
 * Requires graphs[l].n_nodes equal for all l and layers[l+1].input_dim == layers[l].hidden_dim
 * (else RUVGNN_ERR_DIMENSION_MISMATCH).  Intermediates stay in HBM. */
int32_t ruvgnn_stack_forward(ruvgnn_ctx *ctx, const ruvgnn_layer *const *layers,
                             const ruvgnn_graph *const *graphs, uint32_t n_layers, const float *x,
                             float *out, uint32_t flags);

/* Replaces RuvectorLayer::forward(&self, node_embedding, neighbor_embeddings,
 * edge_weights) (layer.rs:379-384) for one node, host buffers.  nbrs is K rows
 * of x_len floats.  x_len != input_dim -> RUVGNN_ERR_DIMENSION_MISMATCH (the
 * reference panics inside ndarray).  n_w != K keeps the reference's zip
 * truncation (layer.rs:438).  Parity/debug entry, not the fast path. */
int32_t ruvgnn_layer_forward_node(ruvgnn_ctx *ctx, const ruvgnn_layer *layer, const float *x,
                                  uint32_t x_len, const float *nbrs, uint32_t k, const float *w,
                                  uint32_t n_w, float *out);

/* ---- query-time search path (crates/ruvector-gnn/src/search.rs) -------------------- */

/* Batched differentiable_search (search.rs:56-94): for each of n_queries queries, cosine
 * similarity against all n_cand candidates (f32 dot, f64 norms, :4-26), temperature softmax
 * over all candidates (:29-44), the k = min(k, n_cand) best by descending weight (ties in
 * index order).  idx_out / w_out: [n_queries][k_eff], k_eff returned in *k_out.  k <= 16.
 * Host pointers (RUVGNN_FWD_DEVICE_PTRS: device pointers). */
int32_t ruvgnn_differentiable_search(ruvgnn_ctx *ctx, const float *queries, uint64_t n_queries,
                                     const float *cands, uint64_t n_cand, uint32_t dim, uint32_t k,
                                     float temperature, uint32_t *idx_out, float *w_out,
                                     uint32_t *k_out, uint32_t flags);

/* Batched hierarchical_forward (search.rs:105-162) for n_queries independent queries:
 * per level l (skipped when level_rows[l] == 0): top-5 differentiable_search at T = 1 over
 * level_emb[l] ([level_rows[l]][level_dim[l]]), aggregated = sum w e, combined =
 * (cur + aggregated)/2, cur = layers[l].forward(combined, selected rows, weights).
 * level_dim[l] must equal the current width and layers[l].input_dim
 * (RUVGNN_ERR_DIMENSION_MISMATCH; the reference panics inside ndarray).
 * out: [n_queries][width after the last applied layer], out_capacity_floats = the TOTAL number
 * of floats `out` can hold (>= n_queries * final width, else RUVGNN_ERR_INVALID_INPUT and *out_dim
 * still receives the width); *out_dim receives that width.  Host pointers only. */
int32_t ruvgnn_hierarchical_forward(ruvgnn_ctx *ctx, const float *queries, uint64_t n_queries,
                                    uint32_t query_dim, const float *const *level_emb,
                                    const uint64_t *level_rows, const uint32_t *level_dim,
                                    const ruvgnn_layer *const *layers, uint32_t n_levels,
                                    float *out, uint64_t out_capacity_floats, uint32_t *out_dim);

/* ---- range-partitioned forward (one process per GPU) -------- */
/* Phase A: project the partition's own rows; x is [n_local][input_dim]
 * (device).  Afterwards ruvgnn_halo_pack gathers the projected K|V|M rows
 * that peers need. */
int32_t ruvgnn_layer_project(ruvgnn_ctx *ctx, const ruvgnn_layer *layer, uint64_t n_local,
                             const float *x_dev);
/* Same for rows [row_begin, row_end) only (x_dev is still the base of all n_local rows): lets a
 * caller project row blocks as their host-to-device copies land.  Every row must be projected
 * before ruvgnn_halo_pack / ruvgnn_layer_messages / ruvgnn_layer_aggregate_update. */
int32_t ruvgnn_layer_project_rows(ruvgnn_ctx *ctx, const ruvgnn_layer *layer, uint64_t n_local,
                                  const float *x_dev, uint64_t row_begin, uint64_t row_end);
/* Floats per halo row of this layer: 3*hidden_dim (K|V|M rows) or hidden_dim (message row
 * only, when the layer runs the single-row gather formulation: hidden_dim == 128, <= 4 heads). */
uint32_t ruvgnn_layer_halo_width(const ruvgnn_layer *layer);
/* Gathers rows `idx[0..n_idx)` of the projected block into a dense
 * [n_idx][ruvgnn_layer_halo_width] device buffer (the halo send buffer). */
int32_t ruvgnn_halo_pack(ruvgnn_ctx *ctx, const ruvgnn_layer *layer, const uint32_t *idx_dev,
                         uint64_t n_idx, float *send_dev);
/* Message-only halo (used when almost every remote row is needed): instead of moving K|V|M
 * rows (3*hidden_dim floats) peers exchange the message rows M = W_msg x + b (hidden_dim
 * floats) and rebuild K and V where they arrive.
 * ruvgnn_layer_messages copies message rows [row_begin, row_begin + n_rows) of the projected
 * block into a dense [n_rows][hidden_dim] device buffer.
 * ruvgnn_layer_expand_halo turns received message rows into halo rows
 * [W_k m + b_k | W_v m + b_v | m] ([n_rows][ruvgnn_layer_halo_width]; a plain copy for
 * layers whose halo width is hidden_dim) with the very GEMM arithmetic of
 * ruvgnn_layer_project (layer.rs:136-137), so the rebuilt rows carry the owner's bits. */
int32_t ruvgnn_layer_messages(ruvgnn_ctx *ctx, const ruvgnn_layer *layer, uint64_t row_begin,
                              uint64_t n_rows, float *out_dev);
int32_t ruvgnn_layer_expand_halo(ruvgnn_ctx *ctx, const ruvgnn_layer *layer,
                                 const float *m_rows_dev, uint64_t n_rows, float *halo_dev);
/* Phase B: gather/aggregate + update for local rows [row_begin,row_end).
 * graph col_idx < n_local address own rows, col_idx >= n_local address
 * halo_dev[col - n_local] ([n_halo][ruvgnn_layer_halo_width], received from peers). */
int32_t ruvgnn_layer_aggregate_update(ruvgnn_ctx *ctx, const ruvgnn_layer *layer,
                                      const ruvgnn_graph *graph, const float *halo_dev,
                                      uint64_t row_begin, uint64_t row_end, float *out_dev);

/* ---- embedding store files as X source / output sink ----------------------------
 * A store is a file of f32 rows at a fixed stride.  The two on-disk layouts of the reference are
 * used in place: MmapManager's headerless dense [max_nodes][d_embed] file
 * (crates/ruvector-gnn/src/mmap.rs:118-186, node i at byte i * d_embed * 4, :199-203) and the
 * cold tier's FeatureStorage (crates/ruvector-gnn/src/cold_tier.rs:17-18,45-75: 24-byte
 * little-endian header dim | num_nodes | block_size, then one page-aligned block per node). */
typedef struct {
  const char *path;          /* the caller keeps the string alive for the duration of the call */
  uint64_t offset_bytes;     /* first row: 0 (MmapManager) or 24 (FeatureStorage) */
  uint64_t row_stride_bytes; /* d_embed * 4, or the FeatureStorage block_size */
  uint32_t dim;              /* floats per row */
  uint64_t rows;             /* rows the file holds */
} ruvgnn_store;
/* MmapManager::new(path, d_embed, max_nodes) view of an existing file (size checked). */
int32_t ruvgnn_store_describe_mmap(const char *path, uint32_t d_embed, uint64_t max_nodes,
                                   ruvgnn_store *out);
/* FeatureStorage::open(path): reads and validates the header. */
int32_t ruvgnn_store_describe_feature_storage(const char *path, ruvgnn_store *out);
/* The batched forward of ruvgnn_layer_forward with X streamed from x_store and the new embeddings
 * streamed into out_store (distinct files; rows [0, graph n_nodes) of each): rows move file ->
 * pinned block -> device block -> projection and back in blocks of block_rows rows (0: 131072),
 * double-buffered, so neither X nor the output is resident as a whole on the host or in HBM.
 * Results are the bits of ruvgnn_layer_forward on the same rows. */
int32_t ruvgnn_layer_forward_store(ruvgnn_ctx *ctx, const ruvgnn_layer *layer,
                                   const ruvgnn_graph *graph, const ruvgnn_store *x_store,
                                   const ruvgnn_store *out_store, uint64_t block_rows);

/* Diagnostic form of the multi-pass gather the partitioned forward uses (below): the layer
 * forward over a single-device graph with the neighbour list of every node processed in
 * n_splits + 1 passes over the column ranges [0, splits[0]), [splits[0], splits[1]), ...,
 * [splits[n_splits-1], n_cols), the online-softmax state carried from pass to pass.  Equals
 * ruvgnn_layer_forward up to f32 summation order.  Device pointers only.  Needs the fast layout
 * (hidden_dim % 128 == 0, heads | 32), else RUVGNN_ERR_INVALID_SHAPE. */
int32_t ruvgnn_layer_forward_split(ruvgnn_ctx *ctx, const ruvgnn_layer *layer,
                                   const ruvgnn_graph *graph, const float *x_dev, float *out_dev,
                                   const uint64_t *splits, uint32_t n_splits);

/* ---- range-partitioned forward inside the library ----------------------------
 * One ruvgnn_dist per GPU: an NCCL communicator (libnccl.so.2 is loaded at run time; absent
 * -> RUVGNN_ERR_NCCL) and a communication stream beside the ctx's compute stream.  The host
 * program is either one process per GPU (ruvgnn_dist_unique_id on rank 0, shipped to the other
 * ranks by any transport, then ruvgnn_dist_create everywhere) or one process with one host
 * thread per GPU (ruvgnn_dist_create_all = ncclCommInitAll; every collective entry below must
 * then be called by all ranks concurrently, one thread each).  The reference has no
 * multi-device path; weights are replicated like its Clone-able layer
 * (crates/ruvector-cli/src/mcp/gnn_cache.rs:183-188). */
#define RUVGNN_NCCL_ID_BYTES 128
typedef struct ruvgnn_dist ruvgnn_dist;
typedef struct ruvgnn_partition ruvgnn_partition;

int32_t ruvgnn_dist_unique_id(uint8_t *id /* [RUVGNN_NCCL_ID_BYTES] */);
/* id == NULL with world > 1 makes a plan-only handle: ruvgnn_partition_create builds this rank's
 * plan (what it needs from whom) without talking to anyone, ruvgnn_dist_forward is refused. */
int32_t ruvgnn_dist_create(ruvgnn_ctx *ctx, const uint8_t *id, int32_t rank, int32_t world,
                           ruvgnn_dist **out);
/* ctxs[r] (one per device, distinct devices) becomes rank r; out: [world] handles. */
int32_t ruvgnn_dist_create_all(ruvgnn_ctx *const *ctxs, int32_t world, ruvgnn_dist **out);
void ruvgnn_dist_destroy(ruvgnn_dist *d);
int32_t ruvgnn_dist_info(const ruvgnn_dist *d, int32_t *rank, int32_t *world);

/* Collective.  This rank's rows of a range-partitioned graph: rank p owns the node range
 * [bounds[p], bounds[p+1]); row_ptr [n_local+1], col_idx GLOBAL node ids, edge_w.  Builds the halo
 * plan on the device: remote columns are de-duplicated (bitmap over the node ids + prefix
 * popcount), every row keeps its neighbours in stored order with columns renamed to
 * [own rows | halo slots], and the request lists are exchanged over NCCL.  When every rank needs
 * >= 75% of all remote rows and the ranges are equal the plan switches to the dense plane (all-gather
 * of the projected rows in row chunks); otherwise the sparse plane (pack + all-to-all-v). */
#define RUVGNN_PART_DEVICE_PTRS 1u /* row_ptr / col_idx / edge_w are device pointers */
#define RUVGNN_PART_FORCE_SPARSE 2u
#define RUVGNN_PART_FORCE_DENSE 4u /* needs equal ranges */
int32_t ruvgnn_partition_create(ruvgnn_dist *d, const uint64_t *bounds, const uint64_t *row_ptr,
                                const uint32_t *col_idx, const float *edge_w, uint32_t flags,
                                ruvgnn_partition **out);
void ruvgnn_partition_destroy(ruvgnn_partition *p);
typedef struct {
  uint64_t n_local, n_total, n_edges;
  uint64_t n_halo;        /* distinct remote rows this rank's rows reference */
  uint64_t n_send;        /* rows peers asked this rank for (sparse plane) */
  uint64_t remote_edges;  /* edges whose endpoint is remote */
  uint64_t interior_rows; /* rows in the contiguous all-local runs (exact mode overlaps them) */
  uint32_t dense;         /* 1: all-gather plane */
  uint32_t chunks;        /* row chunks of the dense plane */
} ruvgnn_partition_info;
int32_t ruvgnn_partition_stats(const ruvgnn_partition *p, ruvgnn_partition_info *out);
/* Host copies of the plan, for checks: the renamed columns [n_edges], the global ids of the halo
 * rows in slot order [n_halo] (sparse plane), the own-local rows sent to peers [n_send], rows
 * received from / sent to each peer [world] each.  Any pointer may be NULL. */
int32_t ruvgnn_partition_download(const ruvgnn_partition *p, uint32_t *col_idx_local,
                                  uint32_t *need_ids, uint32_t *send_idx, uint64_t *recv_counts,
                                  uint64_t *send_counts);

/* Collective.  One layer over the partition: x [n_local][input_dim] -> out [n_local][hidden_dim]
 * (host pointers by default: H2D / D2H pipelined in row blocks; RUVGNN_FWD_DEVICE_PTRS: device).
 *   default     project -> exchange on the communication stream (dense plane: chunked all-gather, K | V of
 *               each chunk rebuilt while the next one is on the wire) -> single-pass gather in stored order
 *               -> update: every row goes through exactly the arithmetic of the single-GPU run, the outputs
 *               are bit-identical to it.  All-local row runs go first, under the exchange.
 *               (RUVGNN_DIST_EXACT names this mode explicitly.)
 *   RUVGNN_DIST_OVERLAP   a gather pass over the own-row edges runs while the halo is on the wire, then
 *               one pass per halo chunk as it lands, the online-softmax state carried in HBM.  Equals the
 *               single-GPU result up to f32 summation order.  Opt-in: on B200 a copy that runs beside the
 *               HBM-saturating gather slows down 2.5-3x, so the passes cost more than they hide
 *               (DESIGN.md section 6).
 *   RUVGNN_DIST_ROW_HALO   dense plane, K|V|M layers: ship the K|V|M rows themselves.  Default there is the
 *               message-row halo: peers ship M = W_msg x + b only (a third of the bytes) and K, V are
 *               rebuilt on arrival with the projection's own GEMM (layer.rs:136-137), so a rebuilt row
 *               carries the owner's bits.  (RUVGNN_DIST_MSG_HALO names the default explicitly.)
 *   RUVGNN_DIST_PEER_DIRECT   sparse plans, one process per GPU: the ranks map each other's
 *               projected tables with CUDA IPC (probed collectively at ruvgnn_dist_create) and the gather
 *               kernel reads the remote rows itself over NVLink -- no pack, no halo buffer, no collective
 *               on the data path; single pass in stored order (bit-identical to the single-GPU run).
 *               Opt-in: the reads are not de-duplicated and SM-issued peer loads reach 290-700 GB/s, so the
 *               pack + all-to-all-v path is faster on every graph measured.  Ignored where IPC mapping failed. */
#define RUVGNN_DIST_EXACT 16u
#define RUVGNN_DIST_MSG_HALO 32u
#define RUVGNN_DIST_PEER_DIRECT 64u
#define RUVGNN_DIST_OVERLAP 128u
#define RUVGNN_DIST_ROW_HALO 256u
int32_t ruvgnn_dist_forward(ruvgnn_dist *d, const ruvgnn_partition *p, const ruvgnn_layer *layer,
                            const float *x, float *out, uint32_t flags);
typedef struct {
  double bytes_sent, bytes_received; /* per forward, this rank */
  float exposed_wait_ms;             /* compute stream stalled on the exchange (last forward) */
  float exchange_ms;                 /* first collective enqueued -> last chunk landed (comm stream) */
  uint32_t passes;                   /* gather passes of the last forward */
  uint32_t exact;                    /* 1: the last forward ran the exact (single-pass) form */
  uint32_t peer_direct;              /* 1: remote rows were read from the peers' tables by the gather */
  uint32_t reserved;
} ruvgnn_dist_stats;
/* Synchronises the ctx stream. */
int32_t ruvgnn_dist_last_stats(ruvgnn_dist *d, ruvgnn_dist_stats *out);
/* Row chunks of the dense plane (1..16; 0 = default: 4)
 * for partitions created afterwards: the all-gather runs chunk by chunk and the overlapped forward
 * runs one gather pass per chunk. */
int32_t ruvgnn_dist_set_chunks(ruvgnn_dist *d, uint32_t chunks);

#ifdef __cplusplus
}
#endif
#endif /* RUVGNN_H */
