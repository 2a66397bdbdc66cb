/*
 * ruvgnn_oracle.c -- CPU restatement of ruvector-gnn's RuvectorLayer forward.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under ruvector_b200/ may link, import or
 * execute this file; only tests/, __graft_entry__.smoke() and bench.py's
 * cpu_baseline / --impl reference legs use it, and there only as the checker
 * or as the timed CPU baseline -- never as the product path.
 *
 * PARITY UNPINNED BY THE REFERENCE: the reference ships no golden vectors for
 * this path and cannot be compiled here (no cargo/rustc;
 * ndarray/serde sources are not vendored), so oracle/_ref does not exist.
 * The restatement is pinned instead by (i) an independent NumPy twin
 * (oracle/oracle_np.py), (ii) hand-derived closed-form cases and (iii) every
 * property the reference's own tests assert (tests/test_oracle.py).
 *
 * Every function cites the reference lines it restates; paths are relative to
 * /root/reference/crates/ruvector-gnn/src/.  Arithmetic is f32 throughout with
 * the reference's operation order; compile with -ffp-contract=off so that no
 * FMA contraction happens (rustc never contracts).
 *
 * Third-party arithmetic: ndarray 0.16.1 (Cargo.lock:5131) performs the
 * mat-vec (`Array2::dot(&ArrayView1)`, layer.rs:39) and `mean()` (layer.rs:72).
 * Without the `blas` feature that is, per output row, ndarray's
 * numeric_util::unrolled_dot (8 partial sums, combined (p0+p4)+(p1+p5)+
 * (p2+p6)+(p3+p7), then the <8 tail) and unrolled_fold for the sum.  The crate
 * source is not in the tree, so the exact association is "as published",
 * restated here; it only moves results by a few ulp.
 */
#include <math.h>
#include <pthread.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

typedef struct {
  const float *w; /* row-major [out][in], as serialized (layer.rs:28-29) */
  const float *b; /* [out] */
} orc_linear;

/* Field order = serde declaration order of the reference structs
 * (layer.rs:315-333, 215-227, 86-93); identical to ruvgnn_weights. */
typedef struct {
  orc_linear w_msg, w_agg;
  orc_linear w_z, u_z, w_r, u_r, w_h, u_h;
  orc_linear q, k, v, out;
  const float *gamma, *beta;
  float eps;
} orc_weights;

typedef struct {
  uint32_t input_dim, hidden_dim, heads;
  float dropout;
  orc_weights w;
} orc_layer;

/* ndarray numeric_util::unrolled_dot (see header). */
static float unrolled_dot(const float *x, const float *y, size_t n) {
  float p0 = 0.f, p1 = 0.f, p2 = 0.f, p3 = 0.f, p4 = 0.f, p5 = 0.f, p6 = 0.f, p7 = 0.f;
  size_t i = 0;
  for (; i + 8 <= n; i += 8) {
    p0 = p0 + x[i + 0] * y[i + 0];
    p1 = p1 + x[i + 1] * y[i + 1];
    p2 = p2 + x[i + 2] * y[i + 2];
    p3 = p3 + x[i + 3] * y[i + 3];
    p4 = p4 + x[i + 4] * y[i + 4];
    p5 = p5 + x[i + 5] * y[i + 5];
    p6 = p6 + x[i + 6] * y[i + 6];
    p7 = p7 + x[i + 7] * y[i + 7];
  }
  float sum = 0.f;
  sum = sum + (p0 + p4);
  sum = sum + (p1 + p5);
  sum = sum + (p2 + p6);
  sum = sum + (p3 + p7);
  for (; i < n; ++i) sum = sum + x[i] * y[i];
  return sum;
}

/* ndarray numeric_util::unrolled_fold with Add, used by ArrayBase::sum(). */
static float unrolled_sum(const float *x, size_t n) {
  float p0 = 0.f, p1 = 0.f, p2 = 0.f, p3 = 0.f, p4 = 0.f, p5 = 0.f, p6 = 0.f, p7 = 0.f;
  size_t i = 0;
  for (; i + 8 <= n; i += 8) {
    p0 = p0 + x[i + 0];
    p1 = p1 + x[i + 1];
    p2 = p2 + x[i + 2];
    p3 = p3 + x[i + 3];
    p4 = p4 + x[i + 4];
    p5 = p5 + x[i + 5];
    p6 = p6 + x[i + 6];
    p7 = p7 + x[i + 7];
  }
  float acc = 0.f;
  acc = acc + (p0 + p4);
  acc = acc + (p1 + p5);
  acc = acc + (p2 + p6);
  acc = acc + (p3 + p7);
  for (; i < n; ++i) acc = acc + x[i];
  return acc;
}

/* Linear::forward, layer.rs:37-41: y = W.dot(x) + b. */
static void linear_fwd(const orc_linear *l, size_t in, size_t out, const float *x, float *y) {
  for (size_t o = 0; o < out; ++o) y[o] = unrolled_dot(l->w + o * in, x, in) + l->b[o];
}

/* LayerNorm::forward, layer.rs:68-81: biased variance, divide by sqrt(var+eps). */
static void layernorm_fwd(const float *gamma, const float *beta, float eps, size_t n,
                          const float *x, float *y) {
  float mean = n ? unrolled_sum(x, n) / (float)n : 0.f; /* x.mean().unwrap_or(0.0) :72 */
  float var = 0.f;
  for (size_t i = 0; i < n; ++i) { /* .map(|v| (v-mean).powi(2)).sum::<f32>() :73 */
    float d = x[i] - mean;
    var = var + d * d;
  }
  var = var / (float)n;
  float denom = sqrtf(var + eps);
  for (size_t i = 0; i < n; ++i) { /* :76,79 */
    float nrm = (x[i] - mean) / denom;
    y[i] = gamma[i] * nrm + beta[i];
  }
}

/* GRUCell::sigmoid, layer.rs:278-285 (branch form). */
static float sigmoid_ref(float x) {
  if (x > 0.f) return 1.f / (1.f + expf(-x));
  float ex = expf(x);
  return ex / (1.f + ex);
}

/* scaled_dot_product_attention for one head, layer.rs:174-210.
 * q: [dh]; k, v: row j at k + j*stride; out: [dh]. */
static void sdpa_head(const float *q, const float *k, const float *v, size_t stride, size_t n_keys,
                      size_t dh, float *scores, float *out) {
  float scale = sqrtf((float)dh); /* :184 */
  for (size_t j = 0; j < n_keys; ++j) { /* :187-193 sequential sum, then divide */
    const float *kj = k + j * stride;
    float dot = 0.f;
    for (size_t t = 0; t < dh; ++t) dot = dot + q[t] * kj[t];
    scores[j] = dot / scale;
  }
  float mx = -INFINITY; /* :196 */
  for (size_t j = 0; j < n_keys; ++j) mx = fmaxf(mx, scores[j]);
  float sum = 0.f; /* :197-198 */
  for (size_t j = 0; j < n_keys; ++j) {
    scores[j] = expf(scores[j] - mx);
    sum = sum + scores[j];
  }
  sum = fmaxf(sum, 1e-10f);
  for (size_t t = 0; t < dh; ++t) out[t] = 0.f; /* :202-207 */
  for (size_t j = 0; j < n_keys; ++j) {
    float wgt = scores[j] / sum; /* :199 */
    const float *vj = v + j * stride;
    for (size_t t = 0; t < dh; ++t) out[t] = out[t] + wgt * vj[t];
  }
}

/* Everything after the projections of one node are known.  Shared by the
 * faithful and the cached graph drivers so both are the same arithmetic.
 *   m      : node message W_msg x + b                       layer.rs:392
 *   nm     : K neighbour messages, row j at nm + j*H         :393-396
 *   kk, vv : K projected keys / values, row j at +j*H        :136-137
 *   ew,n_w : edge weights (zip-truncated like the reference) :421-445
 */
static void node_tail(const orc_layer *L, const float *m, const float *nm, const float *kk,
                      const float *vv, size_t K, const float *ew, size_t n_w, float *scratch,
                      float *out) {
  const size_t H = L->hidden_dim, h = L->heads, dh = H / h;
  const orc_weights *W = &L->w;
  float *q = scratch;           /* H */
  float *concat = q + H;        /* H */
  float *attn = concat + H;     /* H */
  float *agg = attn + H;        /* H */
  float *comb = agg + H;        /* H */
  float *a = comb + H;          /* H */
  float *t1 = a + H;            /* H */
  float *t2 = t1 + H;           /* H */
  float *z = t2 + H;            /* H */
  float *r = z + H;             /* H */
  float *rh = r + H;            /* H */
  float *ht = rh + H;           /* H */
  float *upd = ht + H;          /* H */
  float *scores = upd + H;      /* K */
  float *nw = scores + K;       /* n_w */

  /* MultiHeadAttention::forward :129-160 (keys non-empty here) */
  linear_fwd(&W->q, H, H, m, q);
  for (size_t g = 0; g < h; ++g) /* split_heads :163-171 = contiguous chunks */
    sdpa_head(q + g * dh, kk + g * dh, vv + g * dh, H, K, dh, scores, concat + g * dh);
  linear_fwd(&W->out, H, H, concat, attn);

  /* aggregate_messages :421-445 */
  for (size_t t = 0; t < H; ++t) agg[t] = 0.f;
  if (K > 0 && n_w > 0) {
    float wsum = 0.f;
    for (size_t j = 0; j < n_w; ++j) wsum = wsum + ew[j];
    if (wsum > 0.f)
      for (size_t j = 0; j < n_w; ++j) nw[j] = ew[j] / wsum;
    else
      for (size_t j = 0; j < n_w; ++j) nw[j] = 1.0f / (float)n_w;
    size_t lim = K < n_w ? K : n_w; /* zip truncation :438 */
    for (size_t j = 0; j < lim; ++j)
      for (size_t t = 0; t < H; ++t) agg[t] = agg[t] + nw[j] * nm[j * H + t];
  }

  /* :407-408 */
  for (size_t t = 0; t < H; ++t) comb[t] = attn[t] + agg[t];
  linear_fwd(&W->w_agg, H, H, comb, a);

  /* GRUCell::forward(input=a, hidden=m) :255-275 */
  linear_fwd(&W->w_z, H, H, a, t1);
  linear_fwd(&W->u_z, H, H, m, t2);
  for (size_t t = 0; t < H; ++t) z[t] = sigmoid_ref(t1[t] + t2[t]);
  linear_fwd(&W->w_r, H, H, a, t1);
  linear_fwd(&W->u_r, H, H, m, t2);
  for (size_t t = 0; t < H; ++t) r[t] = sigmoid_ref(t1[t] + t2[t]);
  for (size_t t = 0; t < H; ++t) rh[t] = r[t] * m[t];
  linear_fwd(&W->w_h, H, H, a, t1);
  linear_fwd(&W->u_h, H, H, rh, t2);
  for (size_t t = 0; t < H; ++t) ht[t] = tanhf(t1[t] + t2[t]);
  for (size_t t = 0; t < H; ++t) {
    float omz = 1.0f - z[t];
    upd[t] = omz * m[t] + z[t] * ht[t];
  }
  /* apply_dropout :448-451 (always a scale), LayerNorm :417 */
  float keep = 1.0f - L->dropout;
  for (size_t t = 0; t < H; ++t) upd[t] = upd[t] * keep;
  layernorm_fwd(W->gamma, W->beta, W->eps, H, upd, out);
}

static size_t tail_scratch_floats(size_t H, size_t K, size_t n_w) { return 13 * H + K + n_w + 8; }

/* RuvectorLayer::forward, layer.rs:379-418, the reference's per-node call:
 * every neighbour is re-projected through W_msg, W_k, W_v.
 * x: [D]; nbrs: K rows of D floats; ew: n_w edge weights; out: [H]. */
int orc_layer_forward_node(const orc_layer *L, const float *x, const float *nbrs, uint32_t K,
                           const float *ew, uint32_t n_w, float *out) {
  const size_t D = L->input_dim, H = L->hidden_dim;
  const orc_weights *W = &L->w;
  float *m = (float *)malloc(sizeof(float) * H);
  if (!m) return -1;
  linear_fwd(&W->w_msg, D, H, x, m);
  if (K == 0) { /* :385-389: LN(W_msg x + b); no GRU, no dropout scale */
    layernorm_fwd(W->gamma, W->beta, W->eps, H, m, out);
    free(m);
    return 0;
  }
  float *buf = (float *)malloc(sizeof(float) * (3 * (size_t)K * H + tail_scratch_floats(H, K, n_w)));
  if (!buf) {
    free(m);
    return -1;
  }
  float *nm = buf, *kk = nm + (size_t)K * H, *vv = kk + (size_t)K * H, *scratch = vv + (size_t)K * H;
  for (uint32_t j = 0; j < K; ++j) {
    linear_fwd(&W->w_msg, D, H, nbrs + (size_t)j * D, nm + (size_t)j * H); /* :393-396 */
    linear_fwd(&W->k, H, H, nm + (size_t)j * H, kk + (size_t)j * H);       /* :136 */
    linear_fwd(&W->v, H, H, nm + (size_t)j * H, vv + (size_t)j * H);       /* :137 */
  }
  node_tail(L, m, nm, kk, vv, K, ew, n_w, scratch, out);
  free(buf);
  free(m);
  return 0;
}

/* ---- graph-wide drivers: out[i] = forward(x[i], [x[j] for j in N(i)], w[i]), Jacobi ---- */

typedef struct {
  const orc_layer *L;
  uint64_t begin, end;
  const float *X;
  const uint64_t *row_ptr;
  const uint32_t *col_idx;
  const float *edge_w;
  float *out;
  /* cached variant only */
  const float *M, *Kp, *Vp;
  int rc;
} job_t;

/* The reference callers' loop (verified_training.rs:422-436, stage_iii.rs:143-168):
 * clone neighbour vectors, call forward.  Per-edge re-projection included. */
static void *faithful_worker(void *arg) {
  job_t *jb = (job_t *)arg;
  const size_t D = jb->L->input_dim, H = jb->L->hidden_dim;
  size_t cap = 0;
  float *nb = NULL;
  for (uint64_t i = jb->begin; i < jb->end; ++i) {
    uint64_t b = jb->row_ptr[i], e = jb->row_ptr[i + 1];
    size_t K = (size_t)(e - b);
    if (K > cap) {
      free(nb);
      cap = K;
      nb = (float *)malloc(sizeof(float) * cap * D);
      if (!nb) {
        jb->rc = -1;
        return NULL;
      }
    }
    for (size_t j = 0; j < K; ++j)
      memcpy(nb + j * D, jb->X + (size_t)jb->col_idx[b + j] * D, sizeof(float) * D);
    if (orc_layer_forward_node(jb->L, jb->X + i * D, nb, (uint32_t)K, jb->edge_w + b, (uint32_t)K,
                               jb->out + i * H)) {
      jb->rc = -1;
      break;
    }
  }
  free(nb);
  return NULL;
}

static void *cached_worker(void *arg) {
  job_t *jb = (job_t *)arg;
  const orc_layer *L = jb->L;
  const size_t H = L->hidden_dim;
  size_t cap = 0;
  float *buf = NULL;
  for (uint64_t i = jb->begin; i < jb->end; ++i) {
    uint64_t b = jb->row_ptr[i], e = jb->row_ptr[i + 1];
    size_t K = (size_t)(e - b);
    const float *m = jb->M + i * H;
    if (K == 0) {
      layernorm_fwd(L->w.gamma, L->w.beta, L->w.eps, H, m, jb->out + i * H);
      continue;
    }
    if (K > cap) {
      free(buf);
      cap = K;
      buf = (float *)malloc(sizeof(float) * (3 * cap * H + tail_scratch_floats(H, cap, cap)));
      if (!buf) {
        jb->rc = -1;
        return NULL;
      }
    }
    float *nm = buf, *kk = nm + K * H, *vv = kk + K * H, *scratch = vv + K * H;
    for (size_t j = 0; j < K; ++j) {
      size_t c = jb->col_idx[b + j];
      memcpy(nm + j * H, jb->M + c * H, sizeof(float) * H);
      memcpy(kk + j * H, jb->Kp + c * H, sizeof(float) * H);
      memcpy(vv + j * H, jb->Vp + c * H, sizeof(float) * H);
    }
    node_tail(L, m, nm, kk, vv, K, jb->edge_w + b, K, scratch, jb->out + i * H);
  }
  free(buf);
  return NULL;
}

typedef struct {
  const orc_layer *L;
  uint64_t begin, end;
  const float *X;
  float *M, *Kp, *Vp;
} proj_job_t;

static void *proj_worker(void *arg) {
  proj_job_t *p = (proj_job_t *)arg;
  const size_t D = p->L->input_dim, H = p->L->hidden_dim;
  for (uint64_t i = p->begin; i < p->end; ++i) {
    linear_fwd(&p->L->w.w_msg, D, H, p->X + i * D, p->M + i * H);
    linear_fwd(&p->L->w.k, H, H, p->M + i * H, p->Kp + i * H);
    linear_fwd(&p->L->w.v, H, H, p->M + i * H, p->Vp + i * H);
  }
  return NULL;
}

static int run_threads(void *(*fn)(void *), void *jobs, size_t job_sz, int n) {
  pthread_t *th = (pthread_t *)malloc(sizeof(pthread_t) * (size_t)n);
  if (!th) return -1;
  for (int t = 1; t < n; ++t) pthread_create(&th[t], NULL, fn, (char *)jobs + job_sz * (size_t)t);
  fn(jobs);
  for (int t = 1; t < n; ++t) pthread_join(th[t], NULL);
  free(th);
  return 0;
}

/* Faithful graph forward over nodes [node_begin, node_end) (a bounded sample
 * for the CPU baseline; the whole range for parity).  Writes out rows of the
 * same indices.  n_threads >= 1. */
int orc_forward_graph_faithful(const orc_layer *L, uint64_t n_nodes, const float *X,
                               const uint64_t *row_ptr, const uint32_t *col_idx,
                               const float *edge_w, uint64_t node_begin, uint64_t node_end,
                               float *out, int n_threads) {
  if (n_threads < 1) n_threads = 1;
  if (node_end > n_nodes) node_end = n_nodes;
  if (node_begin >= node_end) return 0;
  job_t *jobs = (job_t *)calloc((size_t)n_threads, sizeof(job_t));
  if (!jobs) return -1;
  uint64_t span = node_end - node_begin, per = (span + (uint64_t)n_threads - 1) / (uint64_t)n_threads;
  for (int t = 0; t < n_threads; ++t) {
    uint64_t b = node_begin + per * (uint64_t)t, e = b + per;
    if (b > node_end) b = node_end;
    if (e > node_end) e = node_end;
    jobs[t] = (job_t){L, b, e, X, row_ptr, col_idx, edge_w, out, NULL, NULL, NULL, 0};
  }
  run_threads(faithful_worker, jobs, sizeof(job_t), n_threads);
  int rc = 0;
  for (int t = 0; t < n_threads; ++t) rc |= jobs[t].rc;
  free(jobs);
  return rc;
}

/* Same values as the faithful driver, bit for bit (each projection is a pure
 * function of its input row), but every node is projected once.  Used by the
 * parity tests at sizes where the faithful loop would take minutes. */
int orc_forward_graph_cached(const orc_layer *L, uint64_t n_nodes, const float *X,
                             const uint64_t *row_ptr, const uint32_t *col_idx, const float *edge_w,
                             float *out, int n_threads) {
  if (n_threads < 1) n_threads = 1;
  const size_t H = L->hidden_dim;
  float *M = (float *)malloc(sizeof(float) * 3 * n_nodes * H + 16);
  if (!M) return -1;
  float *Kp = M + n_nodes * H, *Vp = Kp + n_nodes * H;
  uint64_t per = (n_nodes + (uint64_t)n_threads - 1) / (uint64_t)n_threads;
  proj_job_t *pj = (proj_job_t *)calloc((size_t)n_threads, sizeof(proj_job_t));
  job_t *jobs = (job_t *)calloc((size_t)n_threads, sizeof(job_t));
  if (!pj || !jobs) {
    free(M);
    free(pj);
    free(jobs);
    return -1;
  }
  for (int t = 0; t < n_threads; ++t) {
    uint64_t b = per * (uint64_t)t, e = b + per;
    if (b > n_nodes) b = n_nodes;
    if (e > n_nodes) e = n_nodes;
    pj[t] = (proj_job_t){L, b, e, X, M, Kp, Vp};
    jobs[t] = (job_t){L, b, e, X, row_ptr, col_idx, edge_w, out, M, Kp, Vp, 0};
  }
  run_threads(proj_worker, pj, sizeof(proj_job_t), n_threads);
  run_threads(cached_worker, jobs, sizeof(job_t), n_threads);
  int rc = 0;
  for (int t = 0; t < n_threads; ++t) rc |= jobs[t].rc;
  free(pj);
  free(jobs);
  free(M);
  return rc;
}

/* Exposed pieces so the tests can pin each stage separately. */
void orc_linear_forward(const float *w, const float *b, uint32_t in, uint32_t out, const float *x,
                        float *y) {
  orc_linear l = {w, b};
  linear_fwd(&l, in, out, x, y);
}
void orc_layernorm_forward(const float *gamma, const float *beta, float eps, uint32_t n,
                           const float *x, float *y) {
  layernorm_fwd(gamma, beta, eps, n, x, y);
}
float orc_sigmoid(float x) { return sigmoid_ref(x); }

/* ---- search.rs restatement (hierarchical / differentiable search), rows f2 ---- */

/* cosine_similarity, search.rs:4-26: f32 dot, f64 norms cast to f32. */
float orc_cosine_similarity(const float *a, const float *b, uint32_t n) {
  float dot = 0.f;
  double na = 0.0, nb = 0.0;
  for (uint32_t i = 0; i < n; ++i) dot = dot + a[i] * b[i];
  for (uint32_t i = 0; i < n; ++i) na += (double)a[i] * (double)a[i];
  for (uint32_t i = 0; i < n; ++i) nb += (double)b[i] * (double)b[i];
  float fa = (float)sqrt(na), fb = (float)sqrt(nb);
  if (fa == 0.f || fb == 0.f) return 0.f;
  return dot / (fa * fb);
}

/* softmax with temperature, search.rs:29-44. */
void orc_softmax_t(const float *v, uint32_t n, float temperature, float *out) {
  if (!n) return;
  float mx = -INFINITY;
  for (uint32_t i = 0; i < n; ++i) mx = fmaxf(mx, v[i]);
  float sum = 0.f;
  for (uint32_t i = 0; i < n; ++i) {
    out[i] = expf((v[i] - mx) / temperature);
    sum = sum + out[i];
  }
  sum = fmaxf(sum, 1e-10f);
  for (uint32_t i = 0; i < n; ++i) out[i] = out[i] / sum;
}

/* differentiable_search, search.rs:56-94.  Stable descending sort by weight
 * (Rust's sort_by is stable), top-k.  idx/wts: [k_out]; returns k_out. */
uint32_t orc_differentiable_search(const float *query, const float *cands, uint32_t n_cand,
                                   uint32_t dim, uint32_t k, float temperature, uint32_t *idx,
                                   float *wts) {
  if (!n_cand) return 0;
  if (k > n_cand) k = n_cand;
  float *sim = (float *)malloc(sizeof(float) * 2 * (size_t)n_cand);
  uint32_t *ord = (uint32_t *)malloc(sizeof(uint32_t) * (size_t)n_cand);
  float *sw = sim + n_cand;
  for (uint32_t i = 0; i < n_cand; ++i)
    sim[i] = orc_cosine_similarity(query, cands + (size_t)i * dim, dim);
  orc_softmax_t(sim, n_cand, temperature, sw);
  for (uint32_t i = 0; i < n_cand; ++i) ord[i] = i;
  /* stable selection of the k largest (insertion of the best remaining, ties keep index order) */
  for (uint32_t s = 0; s < k; ++s) {
    uint32_t best = s;
    for (uint32_t t = s + 1; t < n_cand; ++t)
      if (sw[ord[t]] > sw[ord[best]]) best = t;
    uint32_t keep = ord[best];
    for (uint32_t t = best; t > s; --t) ord[t] = ord[t - 1];
    ord[s] = keep;
  }
  for (uint32_t s = 0; s < k; ++s) {
    idx[s] = ord[s];
    wts[s] = sw[ord[s]];
  }
  free(sim);
  free(ord);
  return k;
}

/* hierarchical_forward, search.rs:105-162.  level l has n_level[l] rows of
 * dim_level[l] floats at emb[l]; layers[l] maps dim -> hidden.  cur/out sized
 * for the largest dim involved.  Returns the final length. */
uint32_t orc_hierarchical_forward(const float *query, uint32_t qlen, const float *const *emb,
                                  const uint32_t *n_level, const uint32_t *dim_level,
                                  const orc_layer *const *layers, uint32_t n_levels, float *out) {
  uint32_t len = qlen;
  float *cur = (float *)malloc(sizeof(float) * 4096);
  float *tmp = (float *)malloc(sizeof(float) * 4096 * 8);
  memcpy(cur, query, sizeof(float) * qlen);
  for (uint32_t l = 0; l < n_levels; ++l) {
    if (!n_level[l]) continue; /* :121-123 */
    uint32_t dim = dim_level[l], idx[5];
    float wts[5];
    uint32_t k = orc_differentiable_search(cur, emb[l], n_level[l], dim, 5, 1.0f, idx, wts);
    float *aggv = tmp, *comb = tmp + 4096, *nb = tmp + 2 * 4096;
    for (uint32_t i = 0; i < len; ++i) aggv[i] = 0.f;
    for (uint32_t s = 0; s < k; ++s) /* :133-140 */
      for (uint32_t i = 0; i < dim && i < len; ++i)
        aggv[i] = aggv[i] + wts[s] * emb[l][(size_t)idx[s] * dim + i];
    for (uint32_t i = 0; i < len; ++i) comb[i] = (cur[i] + aggv[i]) / 2.0f; /* :143-147 */
    for (uint32_t s = 0; s < k; ++s)
      memcpy(nb + (size_t)s * dim, emb[l] + (size_t)idx[s] * dim, sizeof(float) * dim);
    orc_layer_forward_node(layers[l], comb, nb, k, wts, k, cur); /* :158 */
    len = layers[l]->hidden_dim;
  }
  memcpy(out, cur, sizeof(float) * len);
  free(cur);
  free(tmp);
  return len;
}
