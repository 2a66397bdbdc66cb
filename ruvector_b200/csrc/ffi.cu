// C ABI of libruvgnn.so (include/ruvgnn.h): context, layer, graph, forward.
// No CPU fallback anywhere: every compute entry needs a CUDA device.
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <atomic>
#include <memory>
#include <string>
#include <vector>

#include "common.cuh"
#include "gemm_tc.cuh"
#include "layer_host.h"
#include "abi_guard.h"
#include "engine.h"

namespace ruvgnn {

static thread_local std::string g_last_error;
void set_error(const std::string &msg) { g_last_error = msg; }
int32_t fail(int32_t code, const std::string &msg) {
  g_last_error = msg;
  return code;
}

}  // namespace ruvgnn

using namespace ruvgnn;

namespace {

int32_t check_config(uint32_t input_dim, uint32_t hidden_dim, uint32_t heads, float dropout) {
  char buf[160];
  // RuvectorLayer::new checks dropout first (layer.rs:353-358) ...
  if (!(dropout >= 0.0f && dropout <= 1.0f)) {
    snprintf(buf, sizeof buf, "Dropout must be between 0.0 and 1.0, got %g", (double)dropout);
    return fail(RUVGNN_ERR_LAYER_CONFIG, buf);
  }
  // ... then MultiHeadAttention::new the head split (layer.rs:101-106).  heads == 0 is a
  // division-by-zero panic in the reference; it is reported as the same config error here.
  if (heads == 0 || hidden_dim % heads != 0) {
    snprintf(buf, sizeof buf,
             "Embedding dimension (%u) must be divisible by number of heads (%u)", hidden_dim,
             heads);
    return fail(RUVGNN_ERR_LAYER_CONFIG, buf);
  }
  if (input_dim == 0 || hidden_dim == 0)
    return fail(RUVGNN_ERR_INVALID_SHAPE, "input_dim and hidden_dim must be non-zero");
  return RUVGNN_OK;
}

}  // namespace

namespace ruvgnn {

// Activation workspace.  Per node (kept for the whole forward, the gather reads them at arbitrary
// rows): P[n][4H] = Q|K|V|M (R = 3) or M[n][H] + Qt[n][hH] (R = 1).  Per row block of at most
// upd_block() rows (reused block after block by K-agg -> K-upd): OA[B][2H] or SA[B][(h+1)H], and
// T0..T3[B][H].  1M nodes x 384: 6.1 + 9.2 GB; 10M nodes x 128 (R = 1, 4 heads): 25.6 + 4.8 GB
// instead of 71.7 GB with everything per node.  With ctx->state_per_node (multi-pass gather of the
// partitioned forward) OA / SA cover all n rows and ML[n][2 heads] holds the running softmax state.
uint64_t upd_block() {
  static uint64_t b = 0;
  if (!b) {
    const char *e = getenv("RUVGNN_UPD_BLOCK");  // tests shrink it to exercise the block loop
    const long long v = e ? atoll(e) : 0;
    b = v > 0 ? (uint64_t)v : (uint64_t)1 << 20;
  }
  return b;
}
static uint32_t ws_node_units(const ruvgnn_layer *L) { return L->r1 ? 1 + L->heads : 4; }
static uint32_t ws_state_units(const ruvgnn_layer *L) { return L->r1 ? L->heads + 1 : 2; }
size_t ws_floats_needed(const ruvgnn_ctx *c, const ruvgnn_layer *L, uint64_t n) {
  const uint64_t B = n < upd_block() ? n : upd_block();
  const uint64_t S = c->state_per_node ? n : B;
  return (size_t)n * L->H * ws_node_units(L) + (size_t)S * L->H * ws_state_units(L) +
         (size_t)B * L->H * 4 + (c->state_per_node ? (size_t)n * 2 * L->heads : 0) + 64;
}

int32_t ensure_ws_floats(ruvgnn_ctx *c, size_t need) {
  if (need <= c->ws_floats) return RUVGNN_OK;
  if (c->ws) {
    RUVGNN_CUDA_TRY(cudaStreamSynchronize(c->stream));
    cudaFree(c->ws);
    c->ws = nullptr;
    c->ws_floats = 0;
  }
  invalidate_projection(c);
  RUVGNN_CUDA_TRY(cudaMalloc(&c->ws, need * sizeof(float)));
  c->ws_floats = need;
  return RUVGNN_OK;
}
int32_t ensure_ws(ruvgnn_ctx *c, const ruvgnn_layer *L, uint64_t n) {
  return ensure_ws_floats(c, ws_floats_needed(c, L, n));
}

int32_t ensure_stage(ruvgnn_ctx *c, float **buf, size_t *have, size_t need) {
  if (need <= *have) return RUVGNN_OK;
  if (*buf) {
    RUVGNN_CUDA_TRY(cudaStreamSynchronize(c->stream));
    cudaFree(*buf);
    *buf = nullptr;
    *have = 0;
  }
  RUVGNN_CUDA_TRY(cudaMalloc(buf, need * sizeof(float)));
  *have = need;
  return RUVGNN_OK;
}

void begin_timing_slot(ruvgnn_ctx *c) {
  c->cur_slot = (int)(c->ring_count % ruvgnn_ctx::kRing);
  c->ev = c->ring[c->cur_slot];
  c->agg_used[c->cur_slot] = 0;
  ++c->ring_count;
}

void invalidate_projection(ruvgnn_ctx *c) {
  c->proj_rows = 0;
  c->proj_layer = nullptr;
  c->proj_done_rows = 0;
}
// Is the workspace a projection of exactly this layer over n rows, carved the way the caller
// is about to carve it?  all_rows: every row has been projected.
bool projection_ready(const ruvgnn_ctx *c, const ruvgnn_layer *L, uint64_t n, bool all_rows) {
  return c->proj_layer == L && c->proj_generation == L->generation && c->proj_rows == n && n > 0 &&
         c->proj_state_rows == c->state_per_node && (!all_rows || c->proj_done_rows >= n);
}

Ws carve(const ruvgnn_ctx *c, const ruvgnn_layer *L, uint64_t n) {
  const uint32_t H = L->H;
  const uint64_t B = n < upd_block() ? n : upd_block();
  const uint64_t S = c->state_per_node ? n : B;  // rows of gather state
  Ws w{};
  if (!L->r1) {
    w.P = c->ws;
    w.ldp = 4 * (int64_t)H;
    w.OA = w.P + (size_t)n * 4 * H;
    w.T0 = w.OA + (size_t)S * 2 * H;  // block-local from here on
    w.m = w.P + 3 * H;
    w.ldm = w.ldp;
    w.halo_src = w.P + H;
    w.halo_w = 3 * H;
  } else {
    const uint32_t h = L->heads;
    w.m = c->ws;
    w.ldm = H;
    w.Qt = w.m + (size_t)n * H;
    w.SA = w.Qt + (size_t)n * h * H;
    w.T0 = w.SA + (size_t)S * (h + 1) * H;  // block-local from here on
    w.halo_src = w.m;
    w.halo_w = H;
  }
  w.T1 = w.T0 + (size_t)B * H;
  w.T2 = w.T1 + (size_t)B * H;
  w.T3 = w.T2 + (size_t)B * H;
  w.ML = c->state_per_node ? w.T3 + (size_t)B * H : nullptr;
  return w;
}

// One GEMM of the layer, dispatched on the layer's arithmetic mode.
static int32_t run_gemm(ruvgnn_ctx *c, const ruvgnn_layer *L, const GemmArgs &g, EpiKind epi,
                        int which) {
  if (L->mode != RUVGNN_MODE_FP32_SIMT) {
    if (!L->tc.usable)
      return fail(RUVGNN_ERR_INVALID_SHAPE, "tensor-core mode on a layer without tensor-core weights");
    RUVGNN_CUDA_TRY(
        launch_gemm_tc(g, epi, L->tc, which, L->mode == RUVGNN_MODE_TF32, c->sm_count, c->stream,
                       &c->launches));
    return RUVGNN_OK;
  }
  RUVGNN_CUDA_TRY(launch_gemm_simt(g, epi, c->stream, &c->launches));
  return RUVGNN_OK;
}

// Phase A (K-proj): M = X W_msg^T + b (layer.rs:392-396, once per node instead of once
// per edge), then Q|K|V = M [W_q;W_k;W_v]^T + b (:135-137).
// Projects rows [rb, re) of an n-row partition (x_dev is the base of all n rows).
// stages: 1 = the message GEMM, 2 = the query (Q|K|V) GEMM, 3 = both.  The partitioned forward
// runs them apart when the halo rows are the message rows, so the exchange starts earlier.
int32_t do_project_rows(ruvgnn_ctx *c, const ruvgnn_layer *L, uint64_t n, const float *x_dev,
                        uint64_t rb, uint64_t re, int stages) {
  const uint32_t H = L->H, D = L->D;
  int32_t rc = ensure_ws(c, L, n);
  if (rc) return rc;
  if (!projection_ready(c, L, n, false) || (rb == 0 && (stages & 1))) {  // a new projection starts
    c->proj_rows = n;
    c->proj_layer = L;
    c->proj_generation = L->generation;
    c->proj_state_rows = c->state_per_node;
    c->proj_done_rows = 0;
  }
  if (re <= rb) return RUVGNN_OK;
  Ws w = carve(c, L, n);
  GemmArgs g;
  g.M = (int64_t)(re - rb);
  g.N = (int32_t)H;
  g.K1 = (int32_t)D;
  g.A1 = x_dev + rb * D;
  g.lda1 = D;
  g.W = L->Wmsg;
  g.bias = L->bmsg;
  g.out = w.m + rb * w.ldm;
  g.ldo = w.ldm;
  if ((stages & 1) && (rc = run_gemm(c, L, g, EPI_BIAS, TC_W_MSG))) return rc;
  if (!(stages & 2)) return RUVGNN_OK;
  GemmArgs q;
  q.M = (int64_t)(re - rb);
  q.K1 = (int32_t)H;
  q.A1 = w.m + rb * w.ldm;
  q.lda1 = w.ldm;
  if (!L->r1) {  // Q | K | V
    q.N = (int32_t)(3 * H);
    q.W = L->Wqkv;
    q.bias = L->bqkv;
    q.out = w.P + rb * w.ldp;
    q.ldo = w.ldp;
  } else {  // folded queries, one H-vector per head
    q.N = (int32_t)(L->heads * H);
    q.W = L->Wqk;
    q.bias = L->bqk;
    q.out = w.Qt + rb * (size_t)(L->heads * H);
    q.ldo = (int64_t)L->heads * H;
  }
  if ((rc = run_gemm(c, L, q, EPI_BIAS, TC_W_QKV))) return rc;
  c->proj_done_rows += re - rb;
  return RUVGNN_OK;
}
int32_t do_project(ruvgnn_ctx *c, const ruvgnn_layer *L, uint64_t n, const float *x_dev) {
  return do_project_rows(c, L, n, x_dev, 0, n, 3);
}

// K-agg over rows [rb, re): one pass of the gather (GatherPass).  The gather state of row r lives
// at index r - sb of OA / SA, where sb is the first row of the current row block (0 when the state
// is kept per node).
int32_t gather_rows(ruvgnn_ctx *c, const ruvgnn_layer *L, const ruvgnn_graph *G, const float *halo,
                    uint64_t rb, uint64_t re, const GatherPass &pass, bool stage_events) {
  const uint32_t H = L->H;
  const uint64_t n = c->proj_rows;
  if (!pass.whole() && !c->state_per_node)
    return fail(RUVGNN_ERR_INVALID_INPUT, "multi-pass gather needs per-node gather state");
  Ws w = carve(c, L, n);
  const int64_t sb = c->state_per_node ? 0 : (int64_t)rb;
  cudaEvent_t agg_end = nullptr;
  if (c->timing_valid && stage_events) {
    std::vector<cudaEvent_t> &pool = c->agg_ev[c->cur_slot];
    uint32_t &used = c->agg_used[c->cur_slot];
    while (pool.size() < 2 * (size_t)(used + 1)) {
      cudaEvent_t e;
      RUVGNN_CUDA_TRY(cudaEventCreate(&e));
      pool.push_back(e);
    }
    if (used == 0) cudaEventRecord(c->ev[6], c->stream);
    cudaEventRecord(pool[2 * used], c->stream);
    agg_end = pool[2 * used + 1];
    ++used;
  }
  if (!L->r1) {
    GatherArgs ga;
    ga.P = w.P;
    ga.ldp = w.ldp;
    ga.halo = halo;
    ga.n_local = n;
    ga.row_ptr = G->row_ptr;
    ga.col_idx = G->col_idx;
    ga.edge_w = G->edge_w;
    ga.oa = w.OA - (ptrdiff_t)(sb * 2 * (int64_t)H);
    ga.row_begin = rb;
    ga.row_end = re;
    ga.H = H;
    ga.heads = L->heads;
    ga.prenormalized = (G->flags & RUVGNN_GRAPH_PRENORMALIZED) ? 1 : 0;
    if (pass.csr) {
      ga.multipass = 1;
      ga.col_idx = pass.csr->col;
      ga.edge_w = pass.csr->w;
      ga.prenormalized = 1;
      ga.seg = pass.csr->seg;
      ga.npass1 = pass.csr->npass + 1;
      ga.pass = pass.index;
    }
    ga.first = pass.first ? 1 : 0;
    ga.last = pass.last ? 1 : 0;
    ga.ml = w.ML;
    if (c->peer_dev && !pass.csr) {
      ga.peer = c->peer_dev;
      ga.col_idx = c->peer_cols;
    }
    RUVGNN_CUDA_TRY(launch_gather(ga, L->fast, c->sm_count, c->stream, &c->launches));
  } else {
    GatherR1Args ga;
    ga.M = w.m;
    ga.ldm = w.ldm;
    ga.halo = halo;
    ga.n_local = n;
    ga.Qt = w.Qt;
    ga.row_ptr = G->row_ptr;
    ga.col_idx = G->col_idx;
    ga.edge_w = G->edge_w;
    ga.sa = w.SA - (ptrdiff_t)(sb * (int64_t)(L->heads + 1) * H);
    ga.row_begin = rb;
    ga.row_end = re;
    ga.heads = L->heads;
    ga.prenormalized = (G->flags & RUVGNN_GRAPH_PRENORMALIZED) ? 1 : 0;
    if (pass.csr) {
      ga.multipass = 1;
      ga.col_idx = pass.csr->col;
      ga.edge_w = pass.csr->w;
      ga.prenormalized = 1;
      ga.seg = pass.csr->seg;
      ga.npass1 = pass.csr->npass + 1;
      ga.pass = pass.index;
    }
    ga.first = pass.first ? 1 : 0;
    ga.last = pass.last ? 1 : 0;
    ga.ml = w.ML;
    if (c->peer_dev && !pass.csr) {
      ga.peer = c->peer_dev;
      ga.col_idx = c->peer_cols;
    }
    RUVGNN_CUDA_TRY(launch_gather_r1(ga, c->sm_count, c->stream, &c->launches));
  }
  if (agg_end) {
    cudaEventRecord(agg_end, c->stream);
    cudaEventRecord(c->ev[3], c->stream);
  }
  return RUVGNN_OK;
}

// K-upd over rows [rb, re), re - rb <= upd_block(): out-proj + combine + W_agg, GRU, dropout
// scale, LayerNorm (layer.rs:156-159,407-417).  All bases are indexed by the row number (already
// shifted by the caller); `scatter` sends row r's output to out[scatter[r]].
struct UpdIO {
  const float *oa = nullptr;  // R = 3: o | agg [row][2H];  R = 1: S_0..S_{h-1} | agg [row][(h+1)H]
  const float *m = nullptr;   // message rows (GRU hidden state)
  int64_t ldm = 0;
  float *T0 = nullptr, *T1 = nullptr, *T2 = nullptr, *T3 = nullptr;  // [row][H]
  const uint64_t *row_ptr = nullptr;  // LayerNorm row select (deg == 0 -> LN(m)); nullptr: every row updated
  const uint32_t *scatter = nullptr;
  float *out = nullptr;
};
static int32_t update_rows_io(ruvgnn_ctx *c, const ruvgnn_layer *L, const UpdIO &io, uint64_t rb,
                              uint64_t re) {
  const uint32_t H = L->H;
  const int64_t rows = (int64_t)(re - rb);
  const float *m = io.m + rb * io.ldm;
  int32_t rc;
  if (L->fused_upd && L->mode != RUVGNN_MODE_FP32_SIMT) {  // H = 128: the whole chain in one kernel
    FusedUpdArgs f;
    f.rows = rows;
    f.row_begin = rb;
    f.A1 = io.oa;
    f.lda1 = L->r1 ? (int64_t)(L->heads + 1) * H : 2 * (int64_t)H;
    f.K1 = (int32_t)f.lda1;
    f.m = io.m;
    f.ldm = io.ldm;
    f.bf = L->bf; f.bzr = L->bzr; f.bh = L->bh; f.gamma = L->gamma; f.beta = L->beta;
    f.eps = L->host.eps;
    f.keep = 1.0f - L->host.dropout;
    f.row_ptr = io.row_ptr;
    f.scatter = io.scatter;
    f.out = io.out;
    RUVGNN_CUDA_TRY(tc::launch_upd_fused(f, L->tc, L->mode == RUVGNN_MODE_TF32, c->sm_count, c->stream,
                                         &c->launches));
    return RUVGNN_OK;
  }
  if (!L->r1) {  // combined = W_o o + b_o + agg
    GemmArgs g;
    g.M = rows; g.N = (int32_t)H; g.K1 = (int32_t)H;
    g.A1 = io.oa + rb * 2 * H; g.lda1 = 2 * H;
    g.W = L->Wo; g.bias = L->bo;
    g.add = io.oa + rb * 2 * H + H; g.ldadd = 2 * H;
    g.out = io.T0 + rb * H; g.ldo = H;
    if ((rc = run_gemm(c, L, g, EPI_BIAS_ADD, TC_W_O))) return rc;
  } else {  // combined = sum_g (W_o[:,g] W_v[g]) S_g + (W_o b_v + b_o) + agg
    const int64_t ld = (int64_t)(L->heads + 1) * H;
    GemmArgs g;
    g.M = rows; g.N = (int32_t)H; g.K1 = (int32_t)(L->heads * H);
    g.A1 = io.oa + rb * ld; g.lda1 = ld;
    g.W = L->Wov; g.bias = L->bov;
    g.add = io.oa + rb * ld + (size_t)L->heads * H; g.ldadd = ld;
    g.out = io.T0 + rb * H; g.ldo = H;
    if ((rc = run_gemm(c, L, g, EPI_BIAS_ADD, TC_W_O))) return rc;
  }
  {  // a = W_agg combined + b_agg
    GemmArgs g;
    g.M = rows; g.N = (int32_t)H; g.K1 = (int32_t)H;
    g.A1 = io.T0 + rb * H; g.lda1 = H;
    g.W = L->Wagg; g.bias = L->bagg;
    g.out = io.T1 + rb * H; g.ldo = H;
    if ((rc = run_gemm(c, L, g, EPI_BIAS, TC_W_AGG))) return rc;
  }
  {  // z = sigma(W_z a + U_z m + b), rh = sigma(W_r a + U_r m + b) * m
    GemmArgs g;
    g.M = rows; g.N = (int32_t)(2 * H); g.K1 = (int32_t)H; g.K2 = (int32_t)H;
    g.A1 = io.T1 + rb * H; g.lda1 = H;
    g.A2 = m; g.lda2 = io.ldm;
    g.W = L->Wzr; g.bias = L->bzr;
    g.m = m; g.ldm = io.ldm;
    g.z = io.T2 + rb * H; g.ldz = H;
    g.rh = io.T3 + rb * H; g.ldrh = H;
    g.H = (int32_t)H;
    if ((rc = run_gemm(c, L, g, EPI_GATES, TC_W_ZR))) return rc;
  }
  {  // d = ((1-z) m + z tanh(W_h a + U_h rh + b)) * (1-p)
    GemmArgs g;
    g.M = rows; g.N = (int32_t)H; g.K1 = (int32_t)H; g.K2 = (int32_t)H;
    g.A1 = io.T1 + rb * H; g.lda1 = H;
    g.A2 = io.T3 + rb * H; g.lda2 = H;
    g.W = L->Wh; g.bias = L->bh;
    g.m = m; g.ldm = io.ldm;
    g.z = io.T2 + rb * H; g.ldz = H;
    g.out = io.T0 + rb * H; g.ldo = H;
    g.H = (int32_t)H;
    g.keep = 1.0f - L->host.dropout;
    if ((rc = run_gemm(c, L, g, EPI_GRU, TC_W_H))) return rc;
  }
  LayerNormArgs ln;
  ln.d = io.T0; ln.ldd = H;
  ln.m = io.m; ln.ldm = io.ldm;
  ln.row_ptr = io.row_ptr;
  ln.gamma = L->gamma; ln.beta = L->beta; ln.eps = L->host.eps;
  ln.out = io.out;
  ln.scatter = io.scatter;
  ln.row_begin = rb; ln.row_end = re;
  ln.H = H; ln.P = L->ord.P;
  RUVGNN_CUDA_TRY(launch_layernorm(ln, c->stream, &c->launches));
  return RUVGNN_OK;
}

int32_t update_rows(ruvgnn_ctx *c, const ruvgnn_layer *L, const ruvgnn_graph *G, uint64_t rb,
                    uint64_t re, float *out_dev) {
  const uint32_t H = L->H;
  Ws w = carve(c, L, c->proj_rows);
  // T0..T3 hold rows [rb, re) only (row r at index r - rb); OA / SA likewise unless the gather
  // state is kept per node.
  const int64_t sb = c->state_per_node ? 0 : (int64_t)rb;
  const int64_t tb = (int64_t)rb;
  UpdIO io;
  io.oa = L->r1 ? w.SA - (ptrdiff_t)(sb * (int64_t)(L->heads + 1) * H) : w.OA - (ptrdiff_t)(sb * 2 * H);
  io.m = w.m;
  io.ldm = w.ldm;
  io.T0 = w.T0 - (ptrdiff_t)(tb * H);
  io.T1 = w.T1 - (ptrdiff_t)(tb * H);
  io.T2 = w.T2 - (ptrdiff_t)(tb * H);
  io.T3 = w.T3 - (ptrdiff_t)(tb * H);
  io.row_ptr = G->row_ptr;
  io.out = out_dev;
  return update_rows_io(c, L, io, rb, re);
}

// Sparse level (R = 1 layers): M = W_msg X + b for every row; out = LN(M) for every row (the
// no-neighbour path, layer.rs:385-389); then only the active rows -- compacted -- go through the
// folded-query GEMM, the gather and the update, and their LayerNorm rows are scattered over the
// LN(M) rows.  Each row sees exactly the arithmetic of the dense path (a GEMM output row depends on
// its own input row only), so the results are bit-identical to it.
static int32_t forward_sparse_level(ruvgnn_ctx *c, const ruvgnn_layer *L, const ruvgnn_graph *G,
                                    const float *x_dev, float *out_dev, bool timing) {
  const uint32_t H = L->H, h = L->heads;
  const uint64_t n = G->n_nodes, na = G->n_active;
  int32_t rc;
  if ((rc = do_project_rows(c, L, n, x_dev, 0, n, 1))) return rc;  // message rows of every node
  Ws w = carve(c, L, n);
  {
    LayerNormArgs ln;  // every row: LN(m); the active rows are overwritten below
    ln.d = w.m; ln.ldd = w.ldm;
    ln.m = w.m; ln.ldm = w.ldm;
    ln.row_ptr = nullptr;
    ln.gamma = L->gamma; ln.beta = L->beta; ln.eps = L->host.eps;
    ln.out = out_dev;
    ln.row_begin = 0; ln.row_end = n;
    ln.H = H; ln.P = L->ord.P;
    RUVGNN_CUDA_TRY(launch_layernorm(ln, c->stream, &c->launches));
  }
  if (timing) {
    cudaEventRecord(c->ev[2], c->stream);
    cudaEventRecord(c->ev[6], c->stream);
    cudaEventRecord(c->ev[3], c->stream);
  }
  // compact buffers live in the (unused) per-node query region: Qc [B][hH] | Mc [B][H]; the
  // per-block regions SA / T0..T3 are used as they are
  const uint64_t B = na < upd_block() ? na : upd_block();
  float *Qc = w.Qt, *Mc = w.Qt + (size_t)B * h * H;
  const int64_t ldsa = (int64_t)(h + 1) * H;
  for (uint64_t cb = 0; cb < na; cb += B) {
    const uint64_t ce = cb + B < na ? cb + B : na;
    RUVGNN_CUDA_TRY(launch_pack_rows(w.m, w.ldm, H, G->active_idx + cb, ce - cb, Mc, c->stream, &c->launches));
    GemmArgs q;  // folded queries of the active rows
    q.M = (int64_t)(ce - cb);
    q.N = (int32_t)(h * H);
    q.K1 = (int32_t)H;
    q.A1 = Mc;
    q.lda1 = H;
    q.W = L->Wqk;
    q.bias = L->bqk;
    q.out = Qc;
    q.ldo = (int64_t)h * H;
    if ((rc = run_gemm(c, L, q, EPI_BIAS, TC_W_QKV))) return rc;
    GatherR1Args ga;  // compact row i = node active_idx[i]; neighbours address the full message table
    ga.M = w.m;
    ga.ldm = w.ldm;
    ga.halo = nullptr;
    ga.n_local = n;
    ga.Qt = Qc - (ptrdiff_t)(cb * (uint64_t)h * H);
    ga.row_ptr = G->row_ptr_active;
    ga.col_idx = G->col_idx;
    ga.edge_w = G->edge_w;
    ga.sa = w.SA - (ptrdiff_t)(cb * (uint64_t)ldsa);
    ga.row_begin = cb;
    ga.row_end = ce;
    ga.heads = h;
    ga.prenormalized = (G->flags & RUVGNN_GRAPH_PRENORMALIZED) ? 1 : 0;
    RUVGNN_CUDA_TRY(launch_gather_r1(ga, c->sm_count, c->stream, &c->launches));
    UpdIO io;
    io.oa = w.SA - (ptrdiff_t)(cb * (uint64_t)ldsa);
    io.m = Mc - (ptrdiff_t)(cb * H);
    io.ldm = H;
    io.T0 = w.T0 - (ptrdiff_t)(cb * H);
    io.T1 = w.T1 - (ptrdiff_t)(cb * H);
    io.T2 = w.T2 - (ptrdiff_t)(cb * H);
    io.T3 = w.T3 - (ptrdiff_t)(cb * H);
    io.row_ptr = nullptr;  // every compact row has neighbours
    io.scatter = G->active_idx;
    io.out = out_dev;
    if ((rc = update_rows_io(c, L, io, cb, ce))) return rc;
  }
  invalidate_projection(c);  // the workspace does not hold a full projection
  return RUVGNN_OK;
}

int32_t forward_device(ruvgnn_ctx *c, const ruvgnn_layer *L, const ruvgnn_graph *G, const float *x_dev,
                       float *out_dev, bool timing) {
  const uint64_t n = G->n_nodes;
  int32_t rc;
  if (L->r1 && G->active_idx && L->heads <= 4) return forward_sparse_level(c, L, G, x_dev, out_dev, timing);
  if ((rc = do_project(c, L, n, x_dev))) return rc;
  if (timing) cudaEventRecord(c->ev[2], c->stream);
  return do_aggregate_update(c, L, G, nullptr, 0, n, out_dev);
}

// Phase B over rows [rb, re): one K-agg -> K-upd pass per block of at most upd_block() rows.
// (Round 2 tried running K-upd of block b on a second, high-priority stream on 32-64 SMs beside K-agg of
// block b+1: the two contend for L2 / HBM instead of overlapping -- config 2 went from 53.4 to 59.2 ms,
// 4M x 128 from 41.6 to 50.4 ms, profiles/r02h_bench_*_pipeline*.json -- so the blocks run back to back.)
int32_t do_aggregate_update(ruvgnn_ctx *c, const ruvgnn_layer *L, const ruvgnn_graph *G,
                            const float *halo, uint64_t rb, uint64_t re, float *out_dev,
                            bool stage_events, const GatherPass &pass) {
  const uint64_t B = upd_block();
  for (uint64_t b0 = rb; b0 < re; b0 += B) {
    const uint64_t b1 = b0 + B < re ? b0 + B : re;
    int32_t rc = gather_rows(c, L, G, halo, b0, b1, pass, stage_events);
    if (!rc) rc = update_rows(c, L, G, b0, b1, out_dev);
    if (rc) return rc;
  }
  return RUVGNN_OK;
}

}  // namespace ruvgnn

namespace {

// Builds the device image of a layer: permuted / concatenated weights.
int32_t upload_layer(ruvgnn_layer *L) {
  const uint32_t H = L->H, D = L->D;
  const LayerHost &h = L->host;
  L->fast = fast_layout_ok(H, L->heads);
  L->ord.H = H;
  L->ord.P = L->fast ? H / 128 : 0;
  const ColumnOrder ord = L->ord;
  const size_t nW = (size_t)H * D + 3 * (size_t)H * H + 2 * (size_t)H * H + 4 * (size_t)H * H +
                    2 * (size_t)H * H;
  const size_t nB = (size_t)H + 3 * H + H + H + 2 * H + H + 2 * H;
  // + room for the R = 1 folded matrices (2 * heads * H * H + biases), allocated when used
  const size_t nR1 = (H == 128 && L->heads <= 8) ? 2 * (size_t)L->heads * H * H + (size_t)L->heads * H + H + 16 : 0;
  // + the fused K-upd's first-stage matrix [H][(heads + 1) H] (or [H][2H]) and its bias, H = 128 only
  const size_t nF = H == 128 ? (size_t)(L->heads + 1) * H * H + 2 * (size_t)H * H + H + 16 : 0;
  std::vector<float> blob(nW + nB + nR1 + nF + 64, 0.f);
  size_t off = 0;
  auto take = [&](size_t n) {
    float *p = blob.data() + off;
    off += (n + 3) & ~(size_t)3;
    return p;
  };
  float *Wmsg = take((size_t)H * D), *bmsg = take(H);
  float *Wqkv = take(3 * (size_t)H * H), *bqkv = take(3 * H);
  float *Wo = take((size_t)H * H), *bo = take(H);
  float *Wagg = take((size_t)H * H), *bagg = take(H);
  float *Wzr = take(4 * (size_t)H * H), *bzr = take(2 * H);
  float *Wh = take(2 * (size_t)H * H), *bh = take(H);
  float *gamma = take(H), *beta = take(H);
  {
    const char *e = getenv("RUVGNN_R1");
    const uint32_t hh = L->heads;
    // measured on B200 (1M nodes, deg ~48): R=1 wins for <= 4 heads (K-agg 8.7 vs 13.2 ms at 4
    // heads); at 8 heads the per-edge work (8 full-width dots + updates) makes it slower (25.6 ms),
    // so 8 heads stays on the R=3 kernel unless RUVGNN_R1=2 forces it.
    const int want = e ? atoi(e) : 1;
    L->r1 = L->fast && H == 128 && want != 0 &&
            (hh == 1 || hh == 2 || hh == 4 || (hh == 8 && want == 2));
  }
  // rows_perm / cols_perm: does the output / input of this Linear live in phys order?
  auto place = [&](const HostLinear &l, float *W, size_t ldw, size_t row_off, size_t col_off,
                   bool cols_perm) {
    for (uint32_t o = 0; o < l.out; ++o) {
      float *dst = W + (row_off + ord.phys(o)) * ldw + col_off;
      const float *src = l.w.data() + (size_t)o * l.in;
      if (cols_perm)
        for (uint32_t i = 0; i < l.in; ++i) dst[ord.phys(i)] = src[i];
      else
        memcpy(dst, src, sizeof(float) * l.in);
    }
  };
  auto place_b = [&](const HostLinear &l, float *b, size_t row_off, bool accumulate) {
    for (uint32_t o = 0; o < l.out; ++o) {
      if (accumulate)
        b[row_off + ord.phys(o)] += l.b[o];
      else
        b[row_off + ord.phys(o)] = l.b[o];
    }
  };
  place(h.lin[L_W_MSG], Wmsg, D, 0, 0, false);
  place_b(h.lin[L_W_MSG], bmsg, 0, false);
  place(h.lin[L_Q], Wqkv, H, 0, 0, true);
  place(h.lin[L_K], Wqkv, H, H, 0, true);
  place(h.lin[L_V], Wqkv, H, 2 * H, 0, true);
  place_b(h.lin[L_Q], bqkv, 0, false);
  place_b(h.lin[L_K], bqkv, H, false);
  place_b(h.lin[L_V], bqkv, 2 * H, false);
  place(h.lin[L_OUT], Wo, H, 0, 0, true);
  place_b(h.lin[L_OUT], bo, 0, false);
  place(h.lin[L_W_AGG], Wagg, H, 0, 0, true);
  place_b(h.lin[L_W_AGG], bagg, 0, false);
  place(h.lin[L_W_Z], Wzr, 2 * H, 0, 0, true);
  place(h.lin[L_U_Z], Wzr, 2 * H, 0, H, true);
  place(h.lin[L_W_R], Wzr, 2 * H, H, 0, true);
  place(h.lin[L_U_R], Wzr, 2 * H, H, H, true);
  place_b(h.lin[L_W_Z], bzr, 0, false);
  place_b(h.lin[L_U_Z], bzr, 0, true);
  place_b(h.lin[L_W_R], bzr, H, false);
  place_b(h.lin[L_U_R], bzr, H, true);
  place(h.lin[L_W_H], Wh, 2 * H, 0, 0, true);
  place(h.lin[L_U_H], Wh, 2 * H, 0, H, true);
  place_b(h.lin[L_W_H], bh, 0, false);
  place_b(h.lin[L_U_H], bh, 0, true);
  for (uint32_t i = 0; i < H; ++i) {
    gamma[ord.phys(i)] = h.gamma[i];
    beta[ord.phys(i)] = h.beta[i];
  }
  float *Wqk = nullptr, *bqk = nullptr, *Wov = nullptr, *bov = nullptr;
  if (L->r1) {
    // P == 1, so the column order is the identity.  Products are formed in double precision.
    const uint32_t hh = L->heads, dh = H / hh;
    Wqk = take((size_t)hh * H * H);
    bqk = take((size_t)hh * H);
    Wov = take((size_t)hh * H * H);
    bov = take(H);
    const HostLinear &lq = h.lin[L_Q], &lk = h.lin[L_K], &lv = h.lin[L_V], &lo = h.lin[L_OUT];
    for (uint32_t g = 0; g < hh; ++g)
      for (uint32_t cidx = 0; cidx < H; ++cidx) {
        // Qt_g[c] = sum_{t in head g} W_k[t][c] * (W_q[t] . m + b_q[t])
        float *row = Wqk + ((size_t)g * H + cidx) * H;
        double bsum = 0.0;
        std::vector<double> acc(H, 0.0);
        for (uint32_t t = g * dh; t < (g + 1) * dh; ++t) {
          const double wk = lk.w[(size_t)t * H + cidx];
          const float *wq = lq.w.data() + (size_t)t * H;
          for (uint32_t i = 0; i < H; ++i) acc[i] += wk * (double)wq[i];
          bsum += wk * (double)lq.b[t];
        }
        for (uint32_t i = 0; i < H; ++i) row[i] = (float)acc[i];
        bqk[(size_t)g * H + cidx] = (float)bsum;
      }
    for (uint32_t o = 0; o < H; ++o) {
      // attn[o] = sum_g sum_c (sum_{t in head g} W_o[o][t] W_v[t][c]) S_g[c] + sum_t W_o[o][t] b_v[t] + b_o[o]
      float *row = Wov + (size_t)o * hh * H;
      double bsum = (double)lo.b[o];
      for (uint32_t g = 0; g < hh; ++g) {
        std::vector<double> acc(H, 0.0);
        for (uint32_t t = g * dh; t < (g + 1) * dh; ++t) {
          const double wo = lo.w[(size_t)o * H + t];
          const float *wv = lv.w.data() + (size_t)t * H;
          for (uint32_t i = 0; i < H; ++i) acc[i] += wo * (double)wv[i];
          bsum += wo * (double)lv.b[t];
        }
        for (uint32_t i = 0; i < H; ++i) row[(size_t)g * H + i] = (float)acc[i];
      }
      bov[o] = (float)bsum;
    }
  }
  // Fused K-upd (gemm_upd_fused.cu), H = 128: out-proj, combine and w_agg (layer.rs:159,407,408) are
  // one linear map of the gather state, a = W_agg (W_o' S + b_o' + agg) + b_agg
  //   = [W_agg W_o' | W_agg] [S | agg] + (W_agg b_o' + b_agg),
  // with (W_o', b_o') = (W_ov, b_ov) over the per-head sums (R = 1) or (W_o, b_o) over o (R = 3).
  // Products in double precision; P == 1, so the column order is the identity.
  float *Wf = nullptr, *bf = nullptr;
  uint32_t KF = 0;
  if (H == 128 && L->fast) {
    const float *Wo1 = L->r1 ? Wov : Wo, *bo1 = L->r1 ? bov : bo;
    const uint32_t Ko = L->r1 ? L->heads * H : H;
    KF = Ko + H;
    Wf = take((size_t)H * KF);
    bf = take(H);
    std::vector<double> acc(Ko);
    for (uint32_t o = 0; o < H; ++o) {
      std::fill(acc.begin(), acc.end(), 0.0);
      double bsum = (double)bagg[o];
      for (uint32_t t = 0; t < H; ++t) {
        const double wa = (double)Wagg[(size_t)o * H + t];
        const float *src = Wo1 + (size_t)t * Ko;
        for (uint32_t i = 0; i < Ko; ++i) acc[i] += wa * (double)src[i];
        bsum += wa * (double)bo1[t];
      }
      float *row = Wf + (size_t)o * KF;
      for (uint32_t i = 0; i < Ko; ++i) row[i] = (float)acc[i];
      memcpy(row + Ko, Wagg + (size_t)o * H, sizeof(float) * H);
      bf[o] = (float)bsum;
    }
  }
  RUVGNN_CUDA_TRY(cudaMalloc(&L->blob, blob.size() * sizeof(float)));
  RUVGNN_CUDA_TRY(
      cudaMemcpy(L->blob, blob.data(), blob.size() * sizeof(float), cudaMemcpyHostToDevice));
  auto dev = [&](const float *p) { return L->blob + (p - blob.data()); };
  L->Wmsg = dev(Wmsg); L->bmsg = dev(bmsg);
  L->Wqkv = dev(Wqkv); L->bqkv = dev(bqkv);
  L->Wo = dev(Wo); L->bo = dev(bo);
  L->Wagg = dev(Wagg); L->bagg = dev(bagg);
  L->Wzr = dev(Wzr); L->bzr = dev(bzr);
  L->Wh = dev(Wh); L->bh = dev(bh);
  L->gamma = dev(gamma); L->beta = dev(beta);
  if (L->r1) {
    L->Wqk = dev(Wqk); L->bqk = dev(bqk);
    L->Wov = dev(Wov); L->bov = dev(bov);
  }
  if (Wf) L->bf = dev(bf);
  // tcgen05 path: split copies (hi/lo) of the device weight matrices
  const float *mats[TC_W_COUNT] = {Wmsg, L->r1 ? Wqk : Wqkv, L->r1 ? Wov : Wo, Wagg, Wzr, Wh,
                                   Wqkv + (size_t)H * H, Wf};
  const uint32_t rows[TC_W_COUNT] = {H, L->r1 ? L->heads * H : 3 * H, H, H, 2 * H, H, 2 * H, Wf ? H : 0};
  const uint32_t cols[TC_W_COUNT] = {D, H, L->r1 ? L->heads * H : H, H, 2 * H, 2 * H, H, KF};
  cudaError_t e = tc_weights_create(mats, rows, cols, &L->tc);
  if (e != cudaSuccess)
    return fail(RUVGNN_ERR_CUDA, std::string("tc_weights_create: ") + cudaGetErrorString(e));
  {
    // K-upd of an H = 128 layer runs as the one fused kernel (gemm_upd_fused.cu) in the tensor-core modes;
    // RUVGNN_FUSED_UPD=0 (read at layer creation) keeps the four GEMMs + LayerNorm (A/B experiments).
    const char *fe = getenv("RUVGNN_FUSED_UPD");
    const int want = fe ? atoi(fe) : 1;
    L->fused_upd = want != 0 && Wf && tc::upd_fused_supported(L->tc, H);
  }
  return RUVGNN_OK;
}

int32_t make_layer(ruvgnn_ctx *ctx, std::unique_ptr<ruvgnn_layer> L, ruvgnn_layer **out) {
  int32_t rc = check_config(L->host.input_dim, L->host.hidden_dim, L->host.heads, L->host.dropout);
  if (rc) return rc;
  L->D = L->host.input_dim;
  L->H = L->host.hidden_dim;
  L->heads = L->host.heads;
  L->device = ctx->device;
  static std::atomic<uint64_t> next_generation{1};
  L->generation = next_generation.fetch_add(1);
  DeviceGuard dg(ctx->device);
  if (!dg.ok) return fail(RUVGNN_ERR_CUDA, "cudaSetDevice failed");
  rc = upload_layer(L.get());
  if (rc) {
    if (L->blob) cudaFree(L->blob);
    tc_weights_destroy(&L->tc);
    return rc;
  }
  // Default arithmetic: the tcgen05 FP32-grade path (3xTF32) wherever the shape allows it, so a
  // drop-in caller that never touches ruvgnn_layer_set_mode gets the measured path; FP32 FFMA
  // only for shapes the tensor-core kernels do not cover (e.g. the reference's 4 -> 8 unit test).
  L->mode = L->tc.usable ? RUVGNN_MODE_FP32_TC : RUVGNN_MODE_FP32_SIMT;
  *out = L.release();
  return RUVGNN_OK;
}

}  // namespace

extern "C" {

const char *ruvgnn_last_error(void) { return g_last_error.c_str(); }
uint32_t ruvgnn_abi_version(void) { return 2; }

int32_t ruvgnn_ctx_create(int32_t device, void *stream, ruvgnn_ctx **out) {
  RUVGNN_ABI_TRY
  if (!out) return fail(RUVGNN_ERR_INVALID_INPUT, "out is null");
  *out = nullptr;
  int ndev = 0;
  cudaError_t e = cudaGetDeviceCount(&ndev);
  if (e != cudaSuccess || ndev == 0) {
    cudaGetLastError();
    return fail(RUVGNN_ERR_CUDA,
                std::string("no CUDA device available (libruvgnn has no CPU fallback): ") +
                    cudaGetErrorString(e));
  }
  if (device < 0 || device >= ndev) return fail(RUVGNN_ERR_INVALID_INPUT, "device out of range");
  DeviceGuard dg(device);
  if (!dg.ok) return fail(RUVGNN_ERR_CUDA, "cudaSetDevice failed");
  cudaDeviceProp prop;
  RUVGNN_CUDA_TRY(cudaGetDeviceProperties(&prop, device));
  if (prop.major != 10)
    return fail(RUVGNN_ERR_CUDA, std::string("libruvgnn is built for sm_100a only; device is sm_") +
                                     std::to_string(prop.major) + std::to_string(prop.minor));
  std::unique_ptr<ruvgnn_ctx> c(new ruvgnn_ctx());
  c->device = device;
  c->sm_count = c->sm_count_dev = prop.multiProcessorCount;
  if (stream) {
    c->stream = static_cast<cudaStream_t>(stream);
  } else {
    RUVGNN_CUDA_TRY(cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking));
    c->own_stream = true;
  }
  RUVGNN_CUDA_TRY(cudaMalloc(&c->d_flag, sizeof(int)));
  for (auto &slot : c->ring)
    for (auto &ev : slot) RUVGNN_CUDA_TRY(cudaEventCreate(&ev));
  RUVGNN_CUDA_TRY(cudaStreamCreateWithFlags(&c->s_in, cudaStreamNonBlocking));
  RUVGNN_CUDA_TRY(cudaStreamCreateWithFlags(&c->s_out, cudaStreamNonBlocking));
  for (auto &ev : c->ev_in) RUVGNN_CUDA_TRY(cudaEventCreateWithFlags(&ev, cudaEventDisableTiming));
  for (auto &ev : c->ev_out) RUVGNN_CUDA_TRY(cudaEventCreateWithFlags(&ev, cudaEventDisableTiming));
  RUVGNN_CUDA_TRY(cudaEventCreateWithFlags(&c->ev_start, cudaEventDisableTiming));
  RUVGNN_CUDA_TRY(cudaEventCreateWithFlags(&c->ev_done, cudaEventDisableTiming));
  *out = c.release();
  return RUVGNN_OK;
  RUVGNN_ABI_CATCH
}

void ruvgnn_ctx_destroy(ruvgnn_ctx *c) {
  if (!c) return;
  DeviceGuard dg(c->device);
  cudaStreamSynchronize(c->stream);
  if (c->ws) cudaFree(c->ws);
  if (c->x_stage) cudaFree(c->x_stage);
  if (c->out_stage) cudaFree(c->out_stage);
  for (auto &b : c->stack_buf)
    if (b) cudaFree(b);
  if (c->d_flag) cudaFree(c->d_flag);
  for (auto &slot : c->ring)
    for (auto &ev : slot)
      if (ev) cudaEventDestroy(ev);
  for (auto &pool : c->agg_ev)
    for (auto &ev : pool) cudaEventDestroy(ev);
  for (auto &ev : c->ev_in)
    if (ev) cudaEventDestroy(ev);
  for (auto &ev : c->ev_out)
    if (ev) cudaEventDestroy(ev);
  if (c->ev_start) cudaEventDestroy(c->ev_start);
  if (c->ev_done) cudaEventDestroy(c->ev_done);
  if (c->s_in) cudaStreamDestroy(c->s_in);
  if (c->s_out) cudaStreamDestroy(c->s_out);
  if (c->own_stream) cudaStreamDestroy(c->stream);
  delete c;
}

int32_t ruvgnn_sync(ruvgnn_ctx *c) {
  RUVGNN_ABI_TRY
  if (!c) return fail(RUVGNN_ERR_INVALID_INPUT, "ctx is null");
  RUVGNN_CUDA_TRY(cudaStreamSynchronize(c->stream));
  return RUVGNN_OK;
  RUVGNN_ABI_CATCH
}
void *ruvgnn_ctx_stream(ruvgnn_ctx *c) { return c ? c->stream : nullptr; }
uint64_t ruvgnn_ctx_launch_count(ruvgnn_ctx *c) { return c ? c->launches : 0; }

int32_t ruvgnn_ctx_set_sm_limit(ruvgnn_ctx *c, int32_t sms) {
  RUVGNN_ABI_TRY
  if (!c) return fail(RUVGNN_ERR_INVALID_INPUT, "ctx is null");
  if (sms < 0) sms = c->sm_count_dev + sms > 1 ? c->sm_count_dev + sms : 1;  // reserve |sms|
  c->sm_count = (sms == 0 || sms > c->sm_count_dev) ? c->sm_count_dev : sms;
  return RUVGNN_OK;
  RUVGNN_ABI_CATCH
}

int32_t ruvgnn_ctx_set_timing(ruvgnn_ctx *c, int32_t on) {
  RUVGNN_ABI_TRY
  if (!c) return fail(RUVGNN_ERR_INVALID_INPUT, "ctx is null");
  c->timing_on = on != 0;
  return RUVGNN_OK;
  RUVGNN_ABI_CATCH
}

int32_t ruvgnn_ctx_last_timing(ruvgnn_ctx *c, float *ms6) {
  RUVGNN_ABI_TRY
  if (!c || !ms6) return fail(RUVGNN_ERR_INVALID_INPUT, "null argument");
  if (c->ring_count == 0) return fail(RUVGNN_ERR_INVALID_INPUT, "no timed forward since last query");
  RUVGNN_CUDA_TRY(cudaStreamSynchronize(c->stream));
  const uint32_t nslots = c->ring_count < (uint32_t)ruvgnn_ctx::kRing ? c->ring_count : ruvgnn_ctx::kRing;
  double acc[6] = {0, 0, 0, 0, 0, 0};
  for (uint32_t sl = 0; sl < nslots; ++sl) {
    cudaEvent_t *e = c->ring[sl];
    float h2d, proj, xchg, agg, upd, d2h;
    RUVGNN_CUDA_TRY(cudaEventElapsedTime(&h2d, e[0], e[1]));
    RUVGNN_CUDA_TRY(cudaEventElapsedTime(&proj, e[1], e[2]));
    RUVGNN_CUDA_TRY(cudaEventElapsedTime(&xchg, e[2], e[6]));
    // K-agg = sum over its launches (one per row block / row range); K-upd = the rest of phase B
    float phase_b;
    RUVGNN_CUDA_TRY(cudaEventElapsedTime(&phase_b, e[6], e[4]));
    agg = 0.f;
    for (uint32_t k = 0; k < c->agg_used[sl]; ++k) {
      float t;
      RUVGNN_CUDA_TRY(cudaEventElapsedTime(&t, c->agg_ev[sl][2 * k], c->agg_ev[sl][2 * k + 1]));
      agg += t;
    }
    if (c->agg_used[sl] == 0) RUVGNN_CUDA_TRY(cudaEventElapsedTime(&agg, e[6], e[3]));
    upd = phase_b - agg;
    RUVGNN_CUDA_TRY(cudaEventElapsedTime(&d2h, e[4], e[5]));
    acc[0] += proj; acc[1] += agg; acc[2] += upd; acc[3] += h2d; acc[4] += d2h; acc[5] += xchg;
  }
  for (int i = 0; i < 6; ++i) ms6[i] = (float)(acc[i] / nslots);
  c->ring_count = 0;
  return RUVGNN_OK;
  RUVGNN_ABI_CATCH
}

int32_t ruvgnn_layer_validate(uint32_t input_dim, uint32_t hidden_dim, uint32_t heads,
                              float dropout) {
  RUVGNN_ABI_TRY
  return check_config(input_dim, hidden_dim, heads, dropout);
  RUVGNN_ABI_CATCH
}

int32_t ruvgnn_layer_create(ruvgnn_ctx *ctx, uint32_t input_dim, uint32_t hidden_dim,
                            uint32_t heads, float dropout, const ruvgnn_weights *w,
                            ruvgnn_layer **out) {
  RUVGNN_ABI_TRY
  if (!out) return fail(RUVGNN_ERR_INVALID_INPUT, "out is null");
  *out = nullptr;
  int32_t rc = check_config(input_dim, hidden_dim, heads, dropout);
  if (rc) return rc;
  if (!ctx || !w) return fail(RUVGNN_ERR_INVALID_INPUT, "ctx / weights is null");
  std::unique_ptr<ruvgnn_layer> L(new ruvgnn_layer());
  LayerHost &h = L->host;
  h.input_dim = input_dim;
  h.hidden_dim = hidden_dim;
  h.heads = heads;
  h.dropout = dropout;
  h.eps = w->eps;
  const ruvgnn_linear *src[L_COUNT] = {&w->w_msg, &w->w_agg, &w->w_z, &w->u_z, &w->w_r, &w->u_r,
                                       &w->w_h,   &w->u_h,   &w->q,   &w->k,   &w->v,   &w->out};
  for (int i = 0; i < L_COUNT; ++i) {
    if (!src[i]->w || !src[i]->b) return fail(RUVGNN_ERR_INVALID_INPUT, "null weight pointer");
    h.lin[i].out = hidden_dim;
    h.lin[i].in = (i == L_W_MSG) ? input_dim : hidden_dim;
    h.lin[i].w.assign(src[i]->w, src[i]->w + (size_t)h.lin[i].out * h.lin[i].in);
    h.lin[i].b.assign(src[i]->b, src[i]->b + hidden_dim);
  }
  if (!w->gamma || !w->beta) return fail(RUVGNN_ERR_INVALID_INPUT, "null gamma/beta");
  h.gamma.assign(w->gamma, w->gamma + hidden_dim);
  h.beta.assign(w->beta, w->beta + hidden_dim);
  return make_layer(ctx, std::move(L), out);
  RUVGNN_ABI_CATCH
}

int32_t ruvgnn_layer_from_json(ruvgnn_ctx *ctx, const char *json, size_t len, ruvgnn_layer **out) {
  RUVGNN_ABI_TRY
  if (!out) return fail(RUVGNN_ERR_INVALID_INPUT, "out is null");
  *out = nullptr;
  if (!ctx || !json) return fail(RUVGNN_ERR_INVALID_INPUT, "ctx / json is null");
  std::unique_ptr<ruvgnn_layer> L(new ruvgnn_layer());
  std::string err = parse_layer_json(json, len, L->host);
  if (!err.empty()) return fail(RUVGNN_ERR_INVALID_INPUT, err);
  return make_layer(ctx, std::move(L), out);
  RUVGNN_ABI_CATCH
}

int32_t ruvgnn_layer_to_json(const ruvgnn_layer *layer, char *buf, size_t cap, size_t *needed) {
  RUVGNN_ABI_TRY
  if (!layer) return fail(RUVGNN_ERR_INVALID_INPUT, "layer is null");
  const std::string s = emit_layer_json(layer->host);
  if (needed) *needed = s.size();
  if (buf && cap) {
    const size_t n = s.size() < cap - 1 ? s.size() : cap - 1;
    memcpy(buf, s.data(), n);
    buf[n] = 0;
  }
  return RUVGNN_OK;
  RUVGNN_ABI_CATCH
}

int32_t ruvgnn_layer_json_normalize(const char *json, size_t len, char *buf, size_t cap,
                                    size_t *needed) {
  RUVGNN_ABI_TRY
  if (!json) return fail(RUVGNN_ERR_INVALID_INPUT, "json is null");
  LayerHost h;
  std::string err = parse_layer_json(json, len, h);
  if (!err.empty()) return fail(RUVGNN_ERR_INVALID_INPUT, err);
  int32_t rc = check_config(h.input_dim, h.hidden_dim, h.heads, h.dropout);
  if (rc) return rc;
  const std::string s = emit_layer_json(h);
  if (needed) *needed = s.size();
  if (buf && cap) {
    const size_t n = s.size() < cap - 1 ? s.size() : cap - 1;
    memcpy(buf, s.data(), n);
    buf[n] = 0;
  }
  return RUVGNN_OK;
  RUVGNN_ABI_CATCH
}

int32_t ruvgnn_layer_dims(const ruvgnn_layer *layer, uint32_t *input_dim, uint32_t *hidden_dim,
                          uint32_t *heads, float *dropout) {
  RUVGNN_ABI_TRY
  if (!layer) return fail(RUVGNN_ERR_INVALID_INPUT, "layer is null");
  if (input_dim) *input_dim = layer->D;
  if (hidden_dim) *hidden_dim = layer->H;
  if (heads) *heads = layer->heads;
  if (dropout) *dropout = layer->host.dropout;
  return RUVGNN_OK;
  RUVGNN_ABI_CATCH
}

int32_t ruvgnn_layer_set_mode(ruvgnn_layer *layer, uint32_t mode) {
  RUVGNN_ABI_TRY
  if (!layer) return fail(RUVGNN_ERR_INVALID_INPUT, "layer is null");
  if (mode > RUVGNN_MODE_TF32) return fail(RUVGNN_ERR_INVALID_INPUT, "unknown mode");
  if (mode != RUVGNN_MODE_FP32_SIMT && !layer->tc.usable)
    return fail(RUVGNN_ERR_INVALID_SHAPE,
                "tensor-core modes need input_dim % 32 == 0 and hidden_dim % 32 == 0");
  layer->mode = mode;
  return RUVGNN_OK;
  RUVGNN_ABI_CATCH
}

uint32_t ruvgnn_layer_mode(const ruvgnn_layer *layer) {
  return layer ? layer->mode : RUVGNN_MODE_FP32_SIMT;
}

uint32_t ruvgnn_layer_update_launches(const ruvgnn_layer *layer) {
  if (!layer) return 0;
  return (layer->fused_upd && layer->mode != RUVGNN_MODE_FP32_SIMT) ? 1u : 5u;
}

void ruvgnn_layer_destroy(ruvgnn_layer *L) {
  if (!L) return;
  DeviceGuard dg(L->device);
  if (L->blob) cudaFree(L->blob);
  tc_weights_destroy(&L->tc);
  delete L;
}

int32_t ruvgnn_graph_create(ruvgnn_ctx *ctx, uint64_t n_nodes, uint64_t n_cols,
                            const uint64_t *row_ptr, const uint32_t *col_idx, const float *edge_w,
                            uint32_t flags, ruvgnn_graph **out) {
  RUVGNN_ABI_TRY
  if (!out) return fail(RUVGNN_ERR_INVALID_INPUT, "out is null");
  *out = nullptr;
  if (!ctx || !row_ptr) return fail(RUVGNN_ERR_INVALID_INPUT, "ctx / row_ptr is null");
  if (n_cols < n_nodes) return fail(RUVGNN_ERR_INVALID_INPUT, "n_cols < n_nodes");
  if (n_cols > 0xffffffffull) return fail(RUVGNN_ERR_INVALID_INPUT, "n_cols exceeds u32 col_idx");
  DeviceGuard dg(ctx->device);
  if (!dg.ok) return fail(RUVGNN_ERR_CUDA, "cudaSetDevice failed");
  const bool dev_ptrs = flags & RUVGNN_GRAPH_DEVICE_PTRS;
  const cudaMemcpyKind kind = dev_ptrs ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice;
  uint64_t ends[2] = {0, 0};
  if (dev_ptrs) {
    RUVGNN_CUDA_TRY(cudaMemcpyAsync(&ends[0], row_ptr, sizeof(uint64_t), cudaMemcpyDeviceToHost, ctx->stream));
    RUVGNN_CUDA_TRY(cudaMemcpyAsync(&ends[1], row_ptr + n_nodes, sizeof(uint64_t), cudaMemcpyDeviceToHost, ctx->stream));
    RUVGNN_CUDA_TRY(cudaStreamSynchronize(ctx->stream));
  } else {
    ends[0] = row_ptr[0];
    ends[1] = row_ptr[n_nodes];
  }
  if (ends[0] != 0) return fail(RUVGNN_ERR_INVALID_INPUT, "row_ptr[0] must be 0");
  const uint64_t n_edges = ends[1];
  if (n_edges && (!col_idx || !edge_w))
    return fail(RUVGNN_ERR_INVALID_INPUT, "col_idx / edge_w is null");
  std::unique_ptr<ruvgnn_graph> g(new ruvgnn_graph());
  g->device = ctx->device;
  g->n_nodes = n_nodes;
  g->n_cols = n_cols;
  g->n_edges = n_edges;
  g->flags = flags & RUVGNN_GRAPH_PRENORMALIZED;
  auto cleanup = [&]() {
    if (g->row_ptr) cudaFree(g->row_ptr);
    if (g->col_idx) cudaFree(g->col_idx);
    if (g->edge_w) cudaFree(g->edge_w);
    if (g->active_idx) cudaFree(g->active_idx);
    if (g->row_ptr_active) cudaFree(g->row_ptr_active);
    g->active_idx = nullptr;
    g->row_ptr_active = nullptr;
    g->row_ptr = nullptr;
    g->col_idx = nullptr;
    g->edge_w = nullptr;
  };
#define GTRY(expr)                                                                           \
  do {                                                                                       \
    cudaError_t _e = (expr);                                                                 \
    if (_e != cudaSuccess) {                                                                 \
      cleanup();                                                                             \
      return fail(_e == cudaErrorMemoryAllocation ? RUVGNN_ERR_OOM : RUVGNN_ERR_CUDA,        \
                  std::string(#expr) + ": " + cudaGetErrorString(_e));                       \
    }                                                                                        \
  } while (0)
  GTRY(cudaMalloc(&g->row_ptr, (n_nodes + 1) * sizeof(uint64_t)));
  GTRY(cudaMalloc(&g->col_idx, (n_edges + 4) * sizeof(uint32_t)));
  GTRY(cudaMalloc(&g->edge_w, (n_edges + 4) * sizeof(float)));
  GTRY(cudaMemcpyAsync(g->row_ptr, row_ptr, (n_nodes + 1) * sizeof(uint64_t), kind, ctx->stream));
  if (n_edges) {
    GTRY(cudaMemcpyAsync(g->col_idx, col_idx, n_edges * sizeof(uint32_t), kind, ctx->stream));
    GTRY(cudaMemcpyAsync(g->edge_w, edge_w, n_edges * sizeof(float), kind, ctx->stream));
  }
  // device-side validation: monotone row_ptr, col_idx < n_cols
  GTRY(cudaMemsetAsync(ctx->d_flag, 0, sizeof(int), ctx->stream));
  GTRY(launch_validate_csr(g->row_ptr, g->col_idx, n_nodes, n_cols, n_edges, ctx->d_flag,
                           ctx->stream, &ctx->launches));
  int flag = 0;
  GTRY(cudaMemcpyAsync(&flag, ctx->d_flag, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
  GTRY(cudaStreamSynchronize(ctx->stream));
  if (flag) {
    cleanup();
    return fail(RUVGNN_ERR_INVALID_INPUT,
                (flag & 1) ? "row_ptr is not monotone non-decreasing"
                           : "col_idx entry out of range (>= n_cols)");
  }
  // sparse level: list the rows that have neighbours when they are fewer than half of all rows
  if (n_nodes >= 1024 && n_cols == n_nodes && !(flags & RUVGNN_GRAPH_NO_ACTIVE_LIST)) {
    std::vector<uint64_t> hrp;
    const uint64_t *rp = row_ptr;
    if (dev_ptrs) {
      hrp.resize(n_nodes + 1);
      GTRY(cudaMemcpy(hrp.data(), g->row_ptr, (n_nodes + 1) * sizeof(uint64_t), cudaMemcpyDeviceToHost));
      rp = hrp.data();
    }
    uint64_t na = 0;
    for (uint64_t i = 0; i < n_nodes; ++i) na += rp[i + 1] > rp[i];
    if (na * 2 < n_nodes) {
      std::vector<uint32_t> idx(na ? na : 1);
      std::vector<uint64_t> rpa(na + 1);
      uint64_t k = 0;
      for (uint64_t i = 0; i < n_nodes; ++i)
        if (rp[i + 1] > rp[i]) {
          idx[k] = (uint32_t)i;
          rpa[k] = rp[i];
          ++k;
        }
      rpa[na] = n_edges;
      // rows between two active rows are empty, so the CSR offsets of consecutive active rows abut
      GTRY(cudaMalloc(&g->active_idx, (na + 4) * sizeof(uint32_t)));
      GTRY(cudaMalloc(&g->row_ptr_active, (na + 1) * sizeof(uint64_t)));
      GTRY(cudaMemcpy(g->active_idx, idx.data(), na * sizeof(uint32_t), cudaMemcpyHostToDevice));
      GTRY(cudaMemcpy(g->row_ptr_active, rpa.data(), (na + 1) * sizeof(uint64_t), cudaMemcpyHostToDevice));
      g->n_active = na;
    }
  }
#undef GTRY
  *out = g.release();
  return RUVGNN_OK;
  RUVGNN_ABI_CATCH
}

int32_t ruvgnn_graph_dims(const ruvgnn_graph *g, uint64_t *n_nodes, uint64_t *n_cols,
                          uint64_t *n_edges) {
  RUVGNN_ABI_TRY
  if (!g) return fail(RUVGNN_ERR_INVALID_INPUT, "graph is null");
  if (n_nodes) *n_nodes = g->n_nodes;
  if (n_cols) *n_cols = g->n_cols;
  if (n_edges) *n_edges = g->n_edges;
  return RUVGNN_OK;
  RUVGNN_ABI_CATCH
}

int32_t ruvgnn_graph_download(const ruvgnn_graph *g, uint64_t *row_ptr, uint32_t *col_idx,
                              float *edge_w) {
  RUVGNN_ABI_TRY
  if (!g) return fail(RUVGNN_ERR_INVALID_INPUT, "graph is null");
  DeviceGuard dg(g->device);
  if (row_ptr)
    RUVGNN_CUDA_TRY(cudaMemcpy(row_ptr, g->row_ptr, (g->n_nodes + 1) * sizeof(uint64_t),
                               cudaMemcpyDeviceToHost));
  if (col_idx && g->n_edges)
    RUVGNN_CUDA_TRY(
        cudaMemcpy(col_idx, g->col_idx, g->n_edges * sizeof(uint32_t), cudaMemcpyDeviceToHost));
  if (edge_w && g->n_edges)
    RUVGNN_CUDA_TRY(
        cudaMemcpy(edge_w, g->edge_w, g->n_edges * sizeof(float), cudaMemcpyDeviceToHost));
  return RUVGNN_OK;
  RUVGNN_ABI_CATCH
}

void ruvgnn_graph_destroy(ruvgnn_graph *g) {
  if (!g) return;
  DeviceGuard dg(g->device);
  if (g->owns) {
    if (g->row_ptr) cudaFree(g->row_ptr);
    if (g->col_idx) cudaFree(g->col_idx);
    if (g->edge_w) cudaFree(g->edge_w);
  }
  if (g->active_idx) cudaFree(g->active_idx);
  if (g->row_ptr_active) cudaFree(g->row_ptr_active);
  delete g;
}

int32_t ruvgnn_layer_forward(ruvgnn_ctx *c, const ruvgnn_layer *L, const ruvgnn_graph *G,
                             const float *x, float *out, uint32_t flags) {
  RUVGNN_ABI_TRY
  if (!c || !L || !G || !x || !out) return fail(RUVGNN_ERR_INVALID_INPUT, "null argument");
  if (L->device != c->device || G->device != c->device)
    return fail(RUVGNN_ERR_INVALID_INPUT, "layer / graph live on another device than ctx");
  if (G->n_cols != G->n_nodes)
    return fail(RUVGNN_ERR_INVALID_INPUT,
                "graph addresses halo rows; use ruvgnn_layer_aggregate_update");
  const uint64_t n = G->n_nodes;
  if (n == 0) return RUVGNN_OK;
  DeviceGuard dg(c->device);
  if (!dg.ok) return fail(RUVGNN_ERR_CUDA, "cudaSetDevice failed");
  const bool dev_ptrs = flags & RUVGNN_FWD_DEVICE_PTRS;
  const bool timing = flags & RUVGNN_FWD_TIMING;
  c->state_per_node = false;
  c->peer_dev = nullptr;
  const float *x_dev = x;
  float *out_dev = out;
  int32_t rc;
  c->timing_valid = false;
  if (!dev_ptrs) {
    if ((rc = ensure_stage(c, &c->x_stage, &c->x_stage_floats, (size_t)n * L->D))) return rc;
    if ((rc = ensure_stage(c, &c->out_stage, &c->out_stage_floats, (size_t)n * L->H))) return rc;
    x_dev = c->x_stage;
    out_dev = c->out_stage;
  } else if (x == out) {
    return fail(RUVGNN_ERR_INVALID_INPUT, "out must not alias x (Jacobi update)");
  }
  if ((rc = ensure_ws(c, L, n))) return rc;
  if (timing) {
    c->timing_valid = true;
    begin_timing_slot(c);
    cudaEventRecord(c->ev[0], c->stream);
  }
  if (dev_ptrs) {
    if (timing) cudaEventRecord(c->ev[1], c->stream);
    if ((rc = forward_device(c, L, G, x_dev, out_dev, timing))) return rc;
    if (timing) {
      cudaEventRecord(c->ev[4], c->stream);
      cudaEventRecord(c->ev[5], c->stream);
    }
  } else {
    // Host buffers: row-block pipeline.  H2D of block b+1 overlaps the projection of block b
    // (copy stream s_in); every row must be projected before any gather (neighbours are
    // arbitrary rows); then gather + update of block b overlaps the D2H of block b-1 (s_out).
    const uint64_t target = 131072;
    uint64_t nb = (n + target - 1) / target;
    if (nb > (uint64_t)ruvgnn_ctx::kMaxBlocks) nb = ruvgnn_ctx::kMaxBlocks;
    if (nb < 1) nb = 1;
    const uint64_t per = ((n + nb - 1) / nb + 127) / 128 * 128;
    RUVGNN_CUDA_TRY(cudaEventRecord(c->ev_start, c->stream));
    RUVGNN_CUDA_TRY(cudaStreamWaitEvent(c->s_in, c->ev_start, 0));
    RUVGNN_CUDA_TRY(cudaStreamWaitEvent(c->s_out, c->ev_start, 0));
    uint64_t used = 0;
    for (uint64_t b = 0; b < nb; ++b) {
      const uint64_t rb = b * per, re = rb + per < n ? rb + per : n;
      if (rb >= n) break;
      ++used;
      RUVGNN_CUDA_TRY(cudaMemcpyAsync(c->x_stage + rb * L->D, x + rb * L->D,
                                      (re - rb) * L->D * sizeof(float), cudaMemcpyHostToDevice,
                                      c->s_in));
      RUVGNN_CUDA_TRY(cudaEventRecord(c->ev_in[b], c->s_in));
    }
    for (uint64_t b = 0; b < used; ++b) {
      const uint64_t rb = b * per, re = rb + per < n ? rb + per : n;
      RUVGNN_CUDA_TRY(cudaStreamWaitEvent(c->stream, c->ev_in[b], 0));
      if (timing && b == used - 1) cudaEventRecord(c->ev[1], c->stream);
      if ((rc = do_project_rows(c, L, n, x_dev, rb, re))) return rc;
    }
    if (timing) {
      cudaEventRecord(c->ev[2], c->stream);
      cudaEventRecord(c->ev[6], c->stream);
      cudaEventRecord(c->ev[3], c->stream);
    }
    for (uint64_t b = 0; b < used; ++b) {
      const uint64_t rb = b * per, re = rb + per < n ? rb + per : n;
      if ((rc = do_aggregate_update(c, L, G, nullptr, rb, re, out_dev, false))) return rc;
      RUVGNN_CUDA_TRY(cudaEventRecord(c->ev_out[b], c->stream));
      RUVGNN_CUDA_TRY(cudaStreamWaitEvent(c->s_out, c->ev_out[b], 0));
      RUVGNN_CUDA_TRY(cudaMemcpyAsync(out + rb * L->H, c->out_stage + rb * L->H,
                                      (re - rb) * L->H * sizeof(float), cudaMemcpyDeviceToHost,
                                      c->s_out));
    }
    if (timing) cudaEventRecord(c->ev[4], c->stream);
    RUVGNN_CUDA_TRY(cudaEventRecord(c->ev_done, c->s_out));
    RUVGNN_CUDA_TRY(cudaStreamWaitEvent(c->stream, c->ev_done, 0));
    if (timing) cudaEventRecord(c->ev[5], c->stream);
  }
  if (!(flags & RUVGNN_FWD_ASYNC) || !dev_ptrs) RUVGNN_CUDA_TRY(cudaStreamSynchronize(c->stream));
  return RUVGNN_OK;
  RUVGNN_ABI_CATCH
}

int32_t ruvgnn_stack_forward(ruvgnn_ctx *c, const ruvgnn_layer *const *layers,
                             const ruvgnn_graph *const *graphs, uint32_t n_layers, const float *x,
                             float *out, uint32_t flags) {
  RUVGNN_ABI_TRY
  if (!c || !layers || !graphs || !x || !out || n_layers == 0)
    return fail(RUVGNN_ERR_INVALID_INPUT, "null argument");
  const uint64_t n = graphs[0] ? graphs[0]->n_nodes : 0;
  for (uint32_t l = 0; l < n_layers; ++l) {
    if (!layers[l] || !graphs[l]) return fail(RUVGNN_ERR_INVALID_INPUT, "null layer / graph");
    if (layers[l]->device != c->device || graphs[l]->device != c->device)
      return fail(RUVGNN_ERR_INVALID_INPUT, "layer / graph live on another device than ctx");
    if (graphs[l]->n_nodes != n || graphs[l]->n_cols != n)
      return fail(RUVGNN_ERR_DIMENSION_MISMATCH,
                  "Tensor dimension mismatch: every level graph must cover the same nodes");
    if (l && layers[l]->D != layers[l - 1]->H) {
      char buf[128];
      snprintf(buf, sizeof buf, "Tensor dimension mismatch: expected %u, got %u (layer %u input)",
               layers[l - 1]->H, layers[l]->D, l);
      return fail(RUVGNN_ERR_DIMENSION_MISMATCH, buf);
    }
  }
  if (n == 0) return RUVGNN_OK;
  DeviceGuard dg(c->device);
  if (!dg.ok) return fail(RUVGNN_ERR_CUDA, "cudaSetDevice failed");
  const bool dev_ptrs = flags & RUVGNN_FWD_DEVICE_PTRS;
  const bool timing = flags & RUVGNN_FWD_TIMING;
  c->timing_valid = false;
  c->state_per_node = false;
  c->peer_dev = nullptr;
  int32_t rc;
  size_t ws_need = 0;
  for (uint32_t l = 0; l < n_layers; ++l) {
    const size_t f = ws_floats_needed(c, layers[l], n);
    ws_need = f > ws_need ? f : ws_need;
  }
  if ((rc = ensure_ws_floats(c, ws_need))) return rc;
  const float *cur = x;
  if (!dev_ptrs) {
    if ((rc = ensure_stage(c, &c->x_stage, &c->x_stage_floats, (size_t)n * layers[0]->D))) return rc;
    RUVGNN_CUDA_TRY(cudaMemcpyAsync(c->x_stage, x, (size_t)n * layers[0]->D * sizeof(float),
                                    cudaMemcpyHostToDevice, c->stream));
    cur = c->x_stage;
    if ((rc = ensure_stage(c, &c->out_stage, &c->out_stage_floats,
                           (size_t)n * layers[n_layers - 1]->H)))
      return rc;
  } else if (x == out) {
    return fail(RUVGNN_ERR_INVALID_INPUT, "out must not alias x (Jacobi update)");
  }
  float *final_dev = dev_ptrs ? out : c->out_stage;
  for (uint32_t l = 0; l < n_layers; ++l) {
    float *dst = final_dev;
    if (l + 1 < n_layers) {
      const int b = (int)(l & 1);
      if ((rc = ensure_stage(c, &c->stack_buf[b], &c->stack_floats[b], (size_t)n * layers[l]->H)))
        return rc;
      dst = c->stack_buf[b];
    }
    if (timing) {  // one timing slot per layer: ruvgnn_ctx_last_timing then reports per-layer averages
      c->timing_valid = true;
      begin_timing_slot(c);
      cudaEventRecord(c->ev[0], c->stream);
      cudaEventRecord(c->ev[1], c->stream);
    }
    if ((rc = forward_device(c, layers[l], graphs[l], cur, dst, timing))) return rc;
    if (timing) {
      cudaEventRecord(c->ev[4], c->stream);
      cudaEventRecord(c->ev[5], c->stream);
    }
    cur = dst;
  }
  if (!dev_ptrs)
    RUVGNN_CUDA_TRY(cudaMemcpyAsync(out, c->out_stage,
                                    (size_t)n * layers[n_layers - 1]->H * sizeof(float),
                                    cudaMemcpyDeviceToHost, c->stream));
  if (!(flags & RUVGNN_FWD_ASYNC) || !dev_ptrs) RUVGNN_CUDA_TRY(cudaStreamSynchronize(c->stream));
  return RUVGNN_OK;
  RUVGNN_ABI_CATCH
}

int32_t ruvgnn_layer_forward_node(ruvgnn_ctx *c, const ruvgnn_layer *L, const float *x,
                                  uint32_t x_len, const float *nbrs, uint32_t k, const float *w,
                                  uint32_t n_w, float *out) {
  RUVGNN_ABI_TRY
  if (!c || !L || !x || !out) return fail(RUVGNN_ERR_INVALID_INPUT, "null argument");
  if (x_len != L->D) {
    char buf[96];
    snprintf(buf, sizeof buf, "Tensor dimension mismatch: expected %u, got %u", L->D, x_len);
    return fail(RUVGNN_ERR_DIMENSION_MISMATCH, buf);
  }
  if (k && !nbrs) return fail(RUVGNN_ERR_INVALID_INPUT, "nbrs is null");
  if (n_w && !w) return fail(RUVGNN_ERR_INVALID_INPUT, "edge weights are null");
  const uint32_t D = L->D, H = L->H;
  // rows: 0 = the node, 1..k = its neighbours (no out-edges: their outputs are discarded)
  std::vector<float> X((size_t)(k + 1) * D);
  memcpy(X.data(), x, sizeof(float) * D);
  if (k) memcpy(X.data() + D, nbrs, sizeof(float) * (size_t)k * D);
  std::vector<uint64_t> rp(k + 2, k);
  rp[0] = 0;
  std::vector<uint32_t> col(k ? k : 1);
  for (uint32_t j = 0; j < k; ++j) col[j] = j + 1;
  // aggregate_messages' normalisation incl. zip truncation (layer.rs:421-445), done on the
  // host in the reference's order so n_w != k behaves identically.
  std::vector<float> nw(k ? k : 1, 0.f);
  if (k && n_w) {
    float wsum = 0.f;
    for (uint32_t j = 0; j < n_w; ++j) wsum = wsum + w[j];
    const uint32_t lim = k < n_w ? k : n_w;
    for (uint32_t j = 0; j < lim; ++j) nw[j] = wsum > 0.f ? w[j] / wsum : 1.0f / (float)n_w;
  }
  ruvgnn_graph *g = nullptr;
  int32_t rc = ruvgnn_graph_create(c, k + 1, k + 1, rp.data(), col.data(), nw.data(),
                                   RUVGNN_GRAPH_PRENORMALIZED, &g);
  if (rc) return rc;
  std::vector<float> O((size_t)(k + 1) * H);
  rc = ruvgnn_layer_forward(c, L, g, X.data(), O.data(), 0);
  ruvgnn_graph_destroy(g);
  if (rc) return rc;
  memcpy(out, O.data(), sizeof(float) * H);
  return RUVGNN_OK;
  RUVGNN_ABI_CATCH
}

int32_t ruvgnn_layer_project(ruvgnn_ctx *c, const ruvgnn_layer *L, uint64_t n_local,
                             const float *x_dev) {
  RUVGNN_ABI_TRY
  if (!c || !L || !x_dev) return fail(RUVGNN_ERR_INVALID_INPUT, "null argument");
  DeviceGuard dg(c->device);
  if (!dg.ok) return fail(RUVGNN_ERR_CUDA, "cudaSetDevice failed");
  c->timing_valid = c->timing_on;
  c->state_per_node = false;
  c->peer_dev = nullptr;
  if (c->timing_on) {
    begin_timing_slot(c);
    cudaEventRecord(c->ev[0], c->stream);
    cudaEventRecord(c->ev[1], c->stream);
  }
  int32_t rc = do_project(c, L, n_local, x_dev);
  if (c->timing_on) cudaEventRecord(c->ev[2], c->stream);
  return rc;
  RUVGNN_ABI_CATCH
}

int32_t ruvgnn_layer_project_rows(ruvgnn_ctx *c, const ruvgnn_layer *L, uint64_t n_local,
                                  const float *x_dev, uint64_t row_begin, uint64_t row_end) {
  RUVGNN_ABI_TRY
  if (!c || !L || !x_dev) return fail(RUVGNN_ERR_INVALID_INPUT, "null argument");
  if (row_end > n_local || row_begin > row_end)
    return fail(RUVGNN_ERR_INVALID_INPUT, "row range out of bounds");
  DeviceGuard dg(c->device);
  if (!dg.ok) return fail(RUVGNN_ERR_CUDA, "cudaSetDevice failed");
  c->timing_valid = c->timing_on;
  c->state_per_node = false;
  c->peer_dev = nullptr;
  if (c->timing_on && row_begin == 0) {
    begin_timing_slot(c);
    cudaEventRecord(c->ev[0], c->stream);
    cudaEventRecord(c->ev[1], c->stream);
  }
  int32_t rc = do_project_rows(c, L, n_local, x_dev, row_begin, row_end);
  if (c->timing_on) cudaEventRecord(c->ev[2], c->stream);
  return rc;
  RUVGNN_ABI_CATCH
}

uint32_t ruvgnn_layer_halo_width(const ruvgnn_layer *L) {
  if (!L) return 0;
  return L->r1 ? L->H : 3 * L->H;
}

int32_t ruvgnn_halo_pack(ruvgnn_ctx *c, const ruvgnn_layer *L, const uint32_t *idx_dev,
                         uint64_t n_idx, float *send_dev) {
  RUVGNN_ABI_TRY
  if (!c || !L) return fail(RUVGNN_ERR_INVALID_INPUT, "null argument");
  if (n_idx == 0) return RUVGNN_OK;
  if (!idx_dev || !send_dev) return fail(RUVGNN_ERR_INVALID_INPUT, "null argument");
  if (!projection_ready(c, L, c->proj_rows, true))
    return fail(RUVGNN_ERR_INVALID_INPUT, "ruvgnn_layer_project has not run for this layer");
  DeviceGuard dg(c->device);
  Ws w = carve(c, L, c->proj_rows);
  RUVGNN_CUDA_TRY(launch_pack_rows(w.halo_src, w.ldm, w.halo_w, idx_dev, n_idx, send_dev,
                                   c->stream, &c->launches));
  return RUVGNN_OK;
  RUVGNN_ABI_CATCH
}

int32_t ruvgnn_layer_messages(ruvgnn_ctx *c, const ruvgnn_layer *L, uint64_t row_begin,
                              uint64_t n_rows, float *out_dev) {
  RUVGNN_ABI_TRY
  if (!c || !L) return fail(RUVGNN_ERR_INVALID_INPUT, "null argument");
  if (n_rows == 0) return RUVGNN_OK;
  if (!out_dev) return fail(RUVGNN_ERR_INVALID_INPUT, "null argument");
  if (!projection_ready(c, L, c->proj_rows, true))
    return fail(RUVGNN_ERR_INVALID_INPUT, "ruvgnn_layer_project has not run for this layer");
  if (row_begin + n_rows > c->proj_rows)
    return fail(RUVGNN_ERR_INVALID_INPUT, "row range out of bounds");
  DeviceGuard dg(c->device);
  if (!dg.ok) return fail(RUVGNN_ERR_CUDA, "cudaSetDevice failed");
  Ws w = carve(c, L, c->proj_rows);
  const size_t rowb = (size_t)L->H * sizeof(float);
  RUVGNN_CUDA_TRY(cudaMemcpy2DAsync(out_dev, rowb, w.m + row_begin * w.ldm,
                                    (size_t)w.ldm * sizeof(float), rowb, n_rows,
                                    cudaMemcpyDeviceToDevice, c->stream));
  return RUVGNN_OK;
  RUVGNN_ABI_CATCH
}

int32_t ruvgnn_layer_expand_halo(ruvgnn_ctx *c, const ruvgnn_layer *L, const float *m_rows,
                                 uint64_t n_rows, float *halo_dev) {
  RUVGNN_ABI_TRY
  if (!c || !L) return fail(RUVGNN_ERR_INVALID_INPUT, "null argument");
  if (n_rows == 0) return RUVGNN_OK;
  if (!m_rows || !halo_dev) return fail(RUVGNN_ERR_INVALID_INPUT, "null argument");
  DeviceGuard dg(c->device);
  if (!dg.ok) return fail(RUVGNN_ERR_CUDA, "cudaSetDevice failed");
  const uint32_t H = L->H;
  const size_t rowb = (size_t)H * sizeof(float);
  if (L->r1) {  // the halo row is the message row itself
    RUVGNN_CUDA_TRY(cudaMemcpyAsync(halo_dev, m_rows, rowb * n_rows, cudaMemcpyDeviceToDevice,
                                    c->stream));
    return RUVGNN_OK;
  }
  GemmArgs g;  // K | V = M [W_k; W_v]^T + b: the K, V column blocks of the Q|K|V projection
  g.M = (int64_t)n_rows;
  g.N = (int32_t)(2 * H);
  g.K1 = (int32_t)H;
  g.A1 = m_rows;
  g.lda1 = H;
  g.W = L->Wqkv + (size_t)H * H;
  g.bias = L->bqkv + H;
  g.out = halo_dev;
  g.ldo = 3 * (int64_t)H;
  int32_t rc = run_gemm(c, L, g, EPI_BIAS, TC_W_KV);
  if (rc) return rc;
  RUVGNN_CUDA_TRY(cudaMemcpy2DAsync(halo_dev + 2 * H, 3 * rowb, m_rows, rowb, rowb, n_rows,
                                    cudaMemcpyDeviceToDevice, c->stream));
  return RUVGNN_OK;
  RUVGNN_ABI_CATCH
}

int32_t ruvgnn_layer_aggregate_update(ruvgnn_ctx *c, const ruvgnn_layer *L, const ruvgnn_graph *G,
                                      const float *halo_dev, uint64_t row_begin, uint64_t row_end,
                                      float *out_dev) {
  RUVGNN_ABI_TRY
  if (!c || !L || !G || !out_dev) return fail(RUVGNN_ERR_INVALID_INPUT, "null argument");
  if (!projection_ready(c, L, G->n_nodes, true))
    return fail(RUVGNN_ERR_INVALID_INPUT,
                "ruvgnn_layer_project must run first (this layer, every row, n_local == graph rows)");
  if (G->n_cols > G->n_nodes && !halo_dev)
    return fail(RUVGNN_ERR_INVALID_INPUT, "graph addresses halo rows but halo_dev is null");
  if (row_end > G->n_nodes || row_begin > row_end)
    return fail(RUVGNN_ERR_INVALID_INPUT, "row range out of bounds");
  DeviceGuard dg(c->device);
  if (!dg.ok) return fail(RUVGNN_ERR_CUDA, "cudaSetDevice failed");
  c->timing_valid = c->timing_on;
  int32_t rc = do_aggregate_update(c, L, G, halo_dev, row_begin, row_end, out_dev);
  if (c->timing_on) {
    cudaEventRecord(c->ev[4], c->stream);
    cudaEventRecord(c->ev[5], c->stream);
  }
  return rc;
  RUVGNN_ABI_CATCH
}

}  // extern "C"

namespace {

struct DevBuf {  // RAII device scratch
  void *p = nullptr;
  ~DevBuf() {
    if (p) cudaFree(p);
  }
  cudaError_t alloc(size_t bytes) { return cudaMalloc(&p, bytes ? bytes : 16); }
  template <class T>
  T *as() const {
    return static_cast<T *>(p);
  }
};

// similarities + top-k for device-resident queries / candidates
int32_t search_device(ruvgnn_ctx *c, const float *q_dev, uint64_t nq, const float *e_dev,
                      uint64_t n, uint32_t dim, int k, float temperature, uint32_t *idx_dev,
                      float *w_dev) {
  DevBuf dots, qn, en, zero;
  RUVGNN_CUDA_TRY(dots.alloc(nq * n * sizeof(float)));
  RUVGNN_CUDA_TRY(qn.alloc(nq * sizeof(float)));
  RUVGNN_CUDA_TRY(en.alloc(n * sizeof(float)));
  RUVGNN_CUDA_TRY(zero.alloc(n * sizeof(float)));
  RUVGNN_CUDA_TRY(cudaMemsetAsync(zero.p, 0, n * sizeof(float), c->stream));
  RUVGNN_CUDA_TRY(launch_row_norms(q_dev, nq, dim, qn.as<float>(), c->stream, &c->launches));
  RUVGNN_CUDA_TRY(launch_row_norms(e_dev, n, dim, en.as<float>(), c->stream, &c->launches));
  GemmArgs g;  // dots[q][c] = <query q, candidate c>  (FP32 FFMA)
  g.M = (int64_t)nq;
  g.N = (int32_t)n;
  g.K1 = (int32_t)dim;
  g.A1 = q_dev;
  g.lda1 = dim;
  g.W = e_dev;
  g.bias = zero.as<float>();
  g.out = dots.as<float>();
  g.ldo = (int64_t)n;
  RUVGNN_CUDA_TRY(launch_gemm_simt(g, EPI_BIAS, c->stream, &c->launches));
  RUVGNN_CUDA_TRY(launch_topk_softmax(dots.as<float>(), qn.as<float>(), en.as<float>(), nq, n, k,
                                      temperature, idx_dev, w_dev, c->stream, &c->launches));
  RUVGNN_CUDA_TRY(cudaStreamSynchronize(c->stream));
  return RUVGNN_OK;
}

}  // namespace

extern "C" int32_t ruvgnn_differentiable_search(ruvgnn_ctx *c, const float *queries,
                                                uint64_t n_queries, const float *cands,
                                                uint64_t n_cand, uint32_t dim, uint32_t k,
                                                float temperature, uint32_t *idx_out, float *w_out,
                                                uint32_t *k_out, uint32_t flags) {
  RUVGNN_ABI_TRY
  if (!c || !k_out) return fail(RUVGNN_ERR_INVALID_INPUT, "null argument");
  *k_out = 0;
  if (n_cand == 0 || n_queries == 0) return RUVGNN_OK;  // search.rs:62-64: empty result
  if (!queries || !cands || !idx_out || !w_out || dim == 0)
    return fail(RUVGNN_ERR_INVALID_INPUT, "null argument");
  if (n_cand > 0x7fffffffull) return fail(RUVGNN_ERR_INVALID_INPUT, "too many candidates");
  uint32_t ke = k < n_cand ? k : (uint32_t)n_cand;  // search.rs:66
  if (ke == 0) return RUVGNN_OK;
  if (ke > 16) return fail(RUVGNN_ERR_INVALID_INPUT, "k > 16 is not supported");
  DeviceGuard dg(c->device);
  if (!dg.ok) return fail(RUVGNN_ERR_CUDA, "cudaSetDevice failed");
  const bool dev = flags & RUVGNN_FWD_DEVICE_PTRS;
  DevBuf qd, ed, id, wd;
  const float *q_dev = queries, *e_dev = cands;
  if (!dev) {
    RUVGNN_CUDA_TRY(qd.alloc(n_queries * dim * sizeof(float)));
    RUVGNN_CUDA_TRY(ed.alloc(n_cand * dim * sizeof(float)));
    RUVGNN_CUDA_TRY(cudaMemcpyAsync(qd.p, queries, n_queries * dim * sizeof(float),
                                    cudaMemcpyHostToDevice, c->stream));
    RUVGNN_CUDA_TRY(cudaMemcpyAsync(ed.p, cands, n_cand * dim * sizeof(float),
                                    cudaMemcpyHostToDevice, c->stream));
    q_dev = qd.as<float>();
    e_dev = ed.as<float>();
  }
  RUVGNN_CUDA_TRY(id.alloc(n_queries * ke * sizeof(uint32_t)));
  RUVGNN_CUDA_TRY(wd.alloc(n_queries * ke * sizeof(float)));
  int32_t rc = search_device(c, q_dev, n_queries, e_dev, n_cand, dim, (int)ke, temperature,
                             id.as<uint32_t>(), wd.as<float>());
  if (rc) return rc;
  const cudaMemcpyKind kind = dev ? cudaMemcpyDeviceToDevice : cudaMemcpyDeviceToHost;
  RUVGNN_CUDA_TRY(cudaMemcpy(idx_out, id.p, n_queries * ke * sizeof(uint32_t), kind));
  RUVGNN_CUDA_TRY(cudaMemcpy(w_out, wd.p, n_queries * ke * sizeof(float), kind));
  *k_out = ke;
  return RUVGNN_OK;
  RUVGNN_ABI_CATCH
}

extern "C" int32_t ruvgnn_hierarchical_forward(ruvgnn_ctx *c, const float *queries,
                                               uint64_t n_queries, uint32_t query_dim,
                                               const float *const *level_emb,
                                               const uint64_t *level_rows,
                                               const uint32_t *level_dim,
                                               const ruvgnn_layer *const *layers,
                                               uint32_t n_levels, float *out,
                                               uint64_t out_capacity_floats, uint32_t *out_dim) {
  RUVGNN_ABI_TRY
  if (!c || !queries || !out || !out_dim) return fail(RUVGNN_ERR_INVALID_INPUT, "null argument");
  if (n_levels && (!level_emb || !level_rows || !level_dim || !layers))
    return fail(RUVGNN_ERR_INVALID_INPUT, "null argument");
  // validate the chain first (the reference would panic inside ndarray on a width mismatch)
  uint32_t width = query_dim;
  for (uint32_t l = 0; l < n_levels; ++l) {
    if (level_rows[l] == 0) continue;  // search.rs:121-123
    if (!layers[l] || !level_emb[l]) return fail(RUVGNN_ERR_INVALID_INPUT, "null layer / level");
    if (layers[l]->device != c->device)
      return fail(RUVGNN_ERR_INVALID_INPUT, "layer lives on another device than ctx");
    if (level_dim[l] != width || layers[l]->D != width) {
      char buf[128];
      snprintf(buf, sizeof buf, "Tensor dimension mismatch: expected %u, got %u (level %u)", width,
               level_dim[l] != width ? level_dim[l] : layers[l]->D, l);
      return fail(RUVGNN_ERR_DIMENSION_MISMATCH, buf);
    }
    width = layers[l]->H;
  }
  *out_dim = width;
  if (n_queries > 0xffffffffffffffffull / (width ? width : 1) ||
      out_capacity_floats < n_queries * (uint64_t)width)
    return fail(RUVGNN_ERR_INVALID_INPUT, "out buffer is too small for n_queries x final width");
  if (n_queries == 0) return RUVGNN_OK;
  DeviceGuard dg(c->device);
  if (!dg.ok) return fail(RUVGNN_ERR_CUDA, "cudaSetDevice failed");
  c->timing_valid = false;
  c->state_per_node = false;
  c->peer_dev = nullptr;
  const uint64_t nq = n_queries;
  DevBuf cur, nxt;
  RUVGNN_CUDA_TRY(cur.alloc(nq * query_dim * sizeof(float)));
  RUVGNN_CUDA_TRY(cudaMemcpyAsync(cur.p, queries, nq * query_dim * sizeof(float),
                                  cudaMemcpyHostToDevice, c->stream));
  width = query_dim;
  for (uint32_t l = 0; l < n_levels; ++l) {
    const uint64_t n = level_rows[l];
    if (n == 0) continue;
    const ruvgnn_layer *L = layers[l];
    const int k = n < 5 ? (int)n : 5;  // search.rs:128
    DevBuf ed, idx, wts, xb, ob;
    RUVGNN_CUDA_TRY(ed.alloc(n * width * sizeof(float)));
    RUVGNN_CUDA_TRY(cudaMemcpyAsync(ed.p, level_emb[l], n * width * sizeof(float),
                                    cudaMemcpyHostToDevice, c->stream));
    RUVGNN_CUDA_TRY(idx.alloc(nq * k * sizeof(uint32_t)));
    RUVGNN_CUDA_TRY(wts.alloc(nq * k * sizeof(float)));
    int32_t rc = search_device(c, cur.as<float>(), nq, ed.as<float>(), n, width, k, 1.0f,
                               idx.as<uint32_t>(), wts.as<float>());
    if (rc) return rc;
    // batch = [combined queries ; selected rows]; query q owns edges to rows nq + q*k + j
    const uint64_t rows = nq * (uint64_t)(k + 1);
    RUVGNN_CUDA_TRY(xb.alloc(rows * width * sizeof(float)));
    RUVGNN_CUDA_TRY(launch_build_batch(cur.as<float>(), ed.as<float>(), idx.as<uint32_t>(),
                                       wts.as<float>(), nq, k, width, xb.as<float>(), c->stream,
                                       &c->launches));
    std::vector<uint64_t> rp(rows + 1);
    std::vector<uint32_t> ci(nq * k);
    for (uint64_t q = 0; q <= rows; ++q) rp[q] = q < nq ? q * k : nq * k;
    for (uint64_t e = 0; e < nq * k; ++e) ci[e] = (uint32_t)(nq + e);
    std::vector<float> ew(nq * k, 1.0f);
    ruvgnn_graph *G = nullptr;
    rc = ruvgnn_graph_create(c, rows, rows, rp.data(), ci.data(), ew.data(), 0, &G);
    if (rc) return rc;
    // the soft weights are the edge weights (search.rs:156): same [q][j] order as the edges
    cudaError_t ce = cudaMemcpyAsync(G->edge_w, wts.p, nq * k * sizeof(float),
                                     cudaMemcpyDeviceToDevice, c->stream);
    if (ce == cudaSuccess) ce = ob.alloc(nq * L->H * sizeof(float));
    if (ce != cudaSuccess) {
      ruvgnn_graph_destroy(G);
      return fail(RUVGNN_ERR_CUDA, cudaGetErrorString(ce));
    }
    rc = do_project(c, L, rows, xb.as<float>());
    if (!rc) rc = do_aggregate_update(c, L, G, nullptr, 0, nq, ob.as<float>());
    cudaStreamSynchronize(c->stream);
    ruvgnn_graph_destroy(G);
    if (rc) return rc;
    std::swap(cur.p, ob.p);
    width = L->H;
  }
  RUVGNN_CUDA_TRY(cudaMemcpyAsync(out, cur.p, nq * width * sizeof(float), cudaMemcpyDeviceToHost,
                                  c->stream));
  RUVGNN_CUDA_TRY(cudaStreamSynchronize(c->stream));
  return RUVGNN_OK;
  RUVGNN_ABI_CATCH
}
