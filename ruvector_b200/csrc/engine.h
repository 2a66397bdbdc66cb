// Internal handles and engine entry points shared by ffi.cu (single-device C ABI) and dist.cu
// (range-partitioned forward with the NCCL halo exchange).  Not part of the public ABI.
#pragma once
#include <vector>

#include "common.cuh"
#include "gemm_tc.cuh"
#include "layer_host.h"

struct ruvgnn_layer;

struct ruvgnn_ctx {
  int device = 0;
  cudaStream_t stream = nullptr;
  bool own_stream = false;
  int sm_count = 148;      // SMs grids are sized for (ruvgnn_ctx_set_sm_limit)
  int sm_count_dev = 148;  // SMs of the device
  uint64_t launches = 0;
  // activation workspace: P[n][4H] | OA[n][2H] | T0..T3[n][H]
  float *ws = nullptr;
  size_t ws_floats = 0;
  // host-pointer staging
  float *x_stage = nullptr, *out_stage = nullptr;
  size_t x_stage_floats = 0, out_stage_floats = 0;
  float *stack_buf[2] = {nullptr, nullptr};  // ping-pong activations of ruvgnn_stack_forward
  size_t stack_floats[2] = {0, 0};
  int *d_flag = nullptr;
  // timing ring: every timed forward records 7 events into the next slot;
  // ruvgnn_ctx_last_timing averages the slots recorded since the previous query.
  // events: 0 start, 1 after h2d, 2 after project, 6 aggregate start (after the halo
  // exchange in the partitioned path), 3 after gather, 4 after update, 5 after d2h
  static constexpr int kRing = 32;
  cudaEvent_t ring[kRing][7] = {};
  // (start, end) event pairs around every K-agg launch of a slot (a forward may run K-agg once
  // per row block / row range / pass); created on first use
  std::vector<cudaEvent_t> agg_ev[kRing];
  uint32_t agg_used[kRing] = {};
  int cur_slot = 0;
  bool timing_on = false;  // ruvgnn_ctx_set_timing: also time project / aggregate_update
  // host-buffer forward: copy streams + per-block events for the H2D | compute | D2H pipeline
  static constexpr int kMaxBlocks = 16;
  cudaStream_t s_in = nullptr, s_out = nullptr;
  cudaEvent_t ev_in[kMaxBlocks] = {}, ev_out[kMaxBlocks] = {}, ev_start = nullptr, ev_done = nullptr;
  cudaEvent_t *ev = ring[0];
  uint32_t ring_count = 0;
  bool timing_valid = false;
  // State of the projected workspace: which layer projected how many rows, with which carve, and
  // whether every row has been projected.  Every entry that rewrites `ws` for another purpose
  // invalidates it, so the phased entries cannot read stale or differently laid-out memory.
  uint64_t proj_rows = 0;
  const ruvgnn_layer *proj_layer = nullptr;
  uint64_t proj_generation = 0;  // generation id of proj_layer at projection time
  bool proj_state_rows = false;  // carve used per-node gather state (multi-pass gather)
  uint64_t proj_done_rows = 0;   // rows projected so far (project_rows calls may arrive in pieces)
  // multi-pass gather (dist.cu): keep the gather state o|agg (or S_g|agg) and the running softmax
  // (max, sum) of every node between passes instead of one row block
  bool state_per_node = false;
  // peer-direct plane (dist.cu): the gather reads remote rows from the peers' tables; col_idx = global ids
  const ruvgnn::PeerTable *peer_dev = nullptr;
  const uint32_t *peer_cols = nullptr;
};

struct ruvgnn_layer {
  ruvgnn::LayerHost host;  // logical layout, for to_json and re-upload
  uint32_t D = 0, H = 0, heads = 0;
  uint32_t mode = RUVGNN_MODE_FP32_SIMT;
  uint64_t generation = 0;  // unique per created layer (address reuse cannot alias a stale projection)
  bool fast = false;
  ruvgnn::ColumnOrder ord;
  int device = 0;
  float *blob = nullptr;  // one allocation, pointers below point inside
  float *Wmsg = nullptr, *bmsg = nullptr;  // [H][D]
  float *Wqkv = nullptr, *bqkv = nullptr;  // [3H][H]
  float *Wo = nullptr, *bo = nullptr;      // [H][H]
  float *Wagg = nullptr, *bagg = nullptr;  // [H][H]
  float *Wzr = nullptr, *bzr = nullptr;    // [2H][2H]: rows z|r, cols a|m
  float *Wh = nullptr, *bh = nullptr;      // [H][2H]: cols a|rh
  float *gamma = nullptr, *beta = nullptr;
  // R = 1 gather formulation (H == 128, heads in {1,2,4,8}; see GatherR1Args): folded weights
  bool r1 = false;
  float *Wqk = nullptr, *bqk = nullptr;  // [heads*H][H]: Qt_g = W_k[g]^T (W_q[g] m + b_q[g])
  float *Wov = nullptr, *bov = nullptr;  // [H][heads*H]: attn = sum_g (W_o[:,g] W_v[g]) S_g + W_o b_v + b_o
  // fused K-upd (H = 128, tensor-core modes): bias of the folded first stage; the matrix is tc TC_W_F
  float *bf = nullptr;
  bool fused_upd = false;
  ruvgnn::TcWeights tc;  // split (hi/lo) copies for the tcgen05 path; empty unless usable
};

struct ruvgnn_graph {
  int device = 0;
  uint64_t n_nodes = 0, n_cols = 0, n_edges = 0;
  uint32_t flags = 0;
  uint64_t *row_ptr = nullptr;
  uint32_t *col_idx = nullptr;
  float *edge_w = nullptr;
  bool owns = true;  // false: the arrays belong to someone else (ruvgnn_partition)
  // Sparse levels (upper HNSW levels: ~n / M^l nodes have a list): the rows with at least one
  // neighbour, and the CSR offsets of just those rows.  Set only when fewer than half of the rows
  // are active; the forward then runs Q projection, gather and update on the active rows alone.
  uint64_t n_active = 0;
  uint32_t *active_idx = nullptr;   // [n_active] ascending node ids
  uint64_t *row_ptr_active = nullptr;  // [n_active + 1]
};

namespace ruvgnn {

struct DeviceGuard {
  int prev = -1;
  bool ok = true;
  explicit DeviceGuard(int dev) {
    if (cudaGetDevice(&prev) != cudaSuccess) {
      ok = false;
      prev = -1;
    }
    if (ok && prev != dev && cudaSetDevice(dev) != cudaSuccess) ok = false;
    if (!ok) cudaGetLastError();
  }
  ~DeviceGuard() {
    if (prev >= 0) cudaSetDevice(prev);
  }
};

struct Ws {
  float *P, *OA, *T0, *T1, *T2, *T3;
  int64_t ldp;
  // views shared by both formulations
  float *m;      // message rows (hidden state of the GRU)
  int64_t ldm;
  float *halo_src;  // rows a peer needs: K|V|M (3H) or M (H)
  uint32_t halo_w;
  // R = 1 only
  float *Qt, *SA;
  float *ML;  // [n][2*heads] running softmax (max, sum) per head; state_per_node only
};

// Pass-ordered copy of a graph's edge arrays (launch_build_pass_csr): every row's edges grouped by
// gather pass, weights normalised, seg[row][0..npass] = the group offsets inside the row.
struct PassCsr {
  uint32_t *col = nullptr;
  float *w = nullptr;
  uint32_t *seg = nullptr;
  int npass = 0;
};
// One pass of the gather.  csr == nullptr: the ordinary single-pass gather over the graph's own
// edge arrays.  Otherwise group `index` of every row of the pass-ordered copy; first: the gather
// state starts from zero; last: the state is normalised and written in the form K-upd reads.
struct GatherPass {
  const PassCsr *csr = nullptr;
  int index = 0;
  bool first = true, last = true;
  bool whole() const { return csr == nullptr; }
};

uint64_t upd_block();
size_t ws_floats_needed(const ruvgnn_ctx *c, const ruvgnn_layer *L, uint64_t n);
int32_t ensure_ws_floats(ruvgnn_ctx *c, size_t need);
int32_t ensure_ws(ruvgnn_ctx *c, const ruvgnn_layer *L, uint64_t n);
int32_t ensure_stage(ruvgnn_ctx *c, float **buf, size_t *have, size_t need);
Ws carve(const ruvgnn_ctx *c, const ruvgnn_layer *L, uint64_t n);
void begin_timing_slot(ruvgnn_ctx *c);
bool projection_ready(const ruvgnn_ctx *c, const ruvgnn_layer *L, uint64_t n, bool all_rows);
void invalidate_projection(ruvgnn_ctx *c);

int32_t do_project_rows(ruvgnn_ctx *c, const ruvgnn_layer *L, uint64_t n, const float *x_dev,
                        uint64_t rb, uint64_t re, int stages = 3);
int32_t do_project(ruvgnn_ctx *c, const ruvgnn_layer *L, uint64_t n, const float *x_dev);
// K-agg over rows [rb, re) (one pass), timed as a K-agg launch when stage events are on.
int32_t gather_rows(ruvgnn_ctx *c, const ruvgnn_layer *L, const ruvgnn_graph *G, const float *halo,
                    uint64_t rb, uint64_t re, const GatherPass &pass, bool stage_events);
// K-upd over rows [rb, re) of at most upd_block() rows (the gather's final pass has run).
int32_t update_rows(ruvgnn_ctx *c, const ruvgnn_layer *L, const ruvgnn_graph *G, uint64_t rb,
                    uint64_t re, float *out_dev);
// Whole layer on device buffers: project -> gather -> update, with the sparse-level form (only the
// rows that have neighbours go through Q projection, gather and update; the others are
// LN(W_msg x + b), layer.rs:385-389) when the graph carries an active-row list and the layer runs
// the single-row gather.  Records the per-stage timing events when `timing`.
int32_t forward_device(ruvgnn_ctx *c, const ruvgnn_layer *L, const ruvgnn_graph *G, const float *x_dev,
                       float *out_dev, bool timing);
// K-agg (single pass, or `pass` as the final pass) + K-upd, block by block.
int32_t do_aggregate_update(ruvgnn_ctx *c, const ruvgnn_layer *L, const ruvgnn_graph *G,
                            const float *halo, uint64_t rb, uint64_t re, float *out_dev,
                            bool stage_events = true, const GatherPass &pass = GatherPass());

}  // namespace ruvgnn
