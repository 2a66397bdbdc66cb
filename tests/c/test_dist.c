/* Plain C caller of the partitioned forward: ONE process, one host thread per GPU
 * (ruvgnn_dist_create_all = ncclCommInitAll), no Python, no torch.  This is the shape a Rust host
 * (the owner of RuvectorLayer, crates/ruvector-gnn/src/layer.rs:315-333) would drive.
 *
 *   ./test_dist [n_gpus]     default 2; exits 77 when fewer devices are present
 *
 * A ring-with-chords graph over N nodes is split into equal ranges; every rank runs
 *   - the single-GPU forward of the WHOLE graph (reference bits),
 *   - ruvgnn_dist_forward (default = exact mode) -> must equal those bits on its own rows,
 *   - ruvgnn_dist_forward RUVGNN_DIST_OVERLAP    -> within 2e-5 row-relative of them,
 * through both data planes. */
#define _POSIX_C_SOURCE 200809L
#include <math.h>
#include <pthread.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "ruvgnn.h"

enum { D = 128, H = 128, HEADS = 4, N = 4096, DEG = 12, MAXW = 8 };

static float *g_x, *g_ref;       /* [N][D] inputs, [N][H] single-GPU outputs */
static uint64_t g_rp[N + 1];
static uint32_t g_col[N * DEG];
static float g_w[N * DEG];
static ruvgnn_weights g_W;
static int g_world;
static ruvgnn_ctx *g_ctx[MAXW];
static ruvgnn_dist *g_dist[MAXW];
static int g_fail[MAXW];
static pthread_barrier_t g_bar;

static float frand(uint32_t *s) {
  *s = *s * 1664525u + 1013904223u;
  return (float)((*s >> 8) & 0xffff) / 65536.0f - 0.5f;
}

#define CHECK(rank, cond)                                                              \
  do {                                                                                 \
    if (!(cond)) {                                                                     \
      printf("  rank %d CHECK FAILED %s:%d: %s (%s)\n", rank, __FILE__, __LINE__, #cond, \
             ruvgnn_last_error());                                                     \
      ++g_fail[rank];                                                                  \
    }                                                                                  \
  } while (0)

static double max_row_rel(const float *a, const float *b, int rows) {
  double worst = 0;
  for (int r = 0; r < rows; ++r) {
    double num = 0, den = 0;
    for (int c = 0; c < H; ++c) {
      const double d = (double)a[r * H + c] - b[r * H + c];
      num += d * d;
      den += (double)b[r * H + c] * b[r * H + c];
    }
    const double rel = sqrt(num) / fmax(sqrt(den), 1e-6);
    if (rel > worst) worst = rel;
  }
  return worst;
}

static void *rank_main(void *arg) {
  const int rank = (int)(size_t)arg;
  ruvgnn_ctx *ctx = g_ctx[rank];
  ruvgnn_dist *dist = g_dist[rank];
  ruvgnn_layer *layer = NULL;
  CHECK(rank, ruvgnn_layer_create(ctx, D, H, HEADS, 0.1f, &g_W, &layer) == RUVGNN_OK);
  /* reference bits: the whole graph on this GPU */
  if (rank == 0) {
    ruvgnn_graph *g = NULL;
    CHECK(rank, ruvgnn_graph_create(ctx, N, N, g_rp, g_col, g_w, 0, &g) == RUVGNN_OK);
    CHECK(rank, ruvgnn_layer_forward(ctx, layer, g, g_x, g_ref, 0) == RUVGNN_OK);
    ruvgnn_graph_destroy(g);
  }
  pthread_barrier_wait(&g_bar);

  uint64_t bounds[MAXW + 1];
  for (int p = 0; p <= g_world; ++p) bounds[p] = (uint64_t)N * p / g_world;
  const uint64_t lo = bounds[rank], hi = bounds[rank + 1], n = hi - lo;
  uint64_t *rp = malloc((n + 1) * sizeof *rp);
  for (uint64_t r = 0; r <= n; ++r) rp[r] = g_rp[lo + r] - g_rp[lo];
  float *out = malloc(n * H * sizeof *out);

  const uint32_t planes[2] = {RUVGNN_PART_FORCE_SPARSE, RUVGNN_PART_FORCE_DENSE};
  for (int pl = 0; pl < 2; ++pl) {
    ruvgnn_partition *part = NULL;
    CHECK(rank, ruvgnn_partition_create(dist, bounds, rp, g_col + g_rp[lo], g_w + g_rp[lo], planes[pl], &part) == RUVGNN_OK);
    if (!part) break;
    ruvgnn_partition_info info;
    CHECK(rank, ruvgnn_partition_stats(part, &info) == RUVGNN_OK);
    CHECK(rank, info.n_local == n && info.n_total == N && info.dense == (uint32_t)pl && info.n_halo > 0);
    /* exact mode: the bits of the single-GPU run */
    memset(out, 0, n * H * sizeof *out);
    CHECK(rank, ruvgnn_dist_forward(dist, part, layer, g_x + lo * D, out, RUVGNN_DIST_EXACT) == RUVGNN_OK);
    CHECK(rank, memcmp(out, g_ref + lo * H, n * H * sizeof *out) == 0);
    memset(out, 0, n * H * sizeof *out);
    CHECK(rank, ruvgnn_dist_forward(dist, part, layer, g_x + lo * D, out, 0) == RUVGNN_OK); /* the default IS the exact mode */
    CHECK(rank, memcmp(out, g_ref + lo * H, n * H * sizeof *out) == 0);
    /* overlapped multi-pass mode: same values up to f32 summation order */
    memset(out, 0, n * H * sizeof *out);
    CHECK(rank, ruvgnn_dist_forward(dist, part, layer, g_x + lo * D, out, RUVGNN_DIST_OVERLAP) == RUVGNN_OK);
    const double rel = max_row_rel(out, g_ref + lo * H, (int)n);
    CHECK(rank, rel <= 2e-5);
    ruvgnn_dist_stats st;
    CHECK(rank, ruvgnn_dist_last_stats(dist, &st) == RUVGNN_OK);
    CHECK(rank, st.passes >= 2 && st.exact == 0 && st.bytes_received > 0);
    if (rank == 0)
      printf("plane %s: halo rows %llu, passes %u, overlapped vs single GPU %.2e, exchange %.3f ms, exposed wait %.3f ms\n",
             pl ? "dense" : "sparse", (unsigned long long)info.n_halo, st.passes, rel, st.exchange_ms, st.exposed_wait_ms);
    ruvgnn_partition_destroy(part);
  }
  free(out);
  free(rp);
  ruvgnn_layer_destroy(layer);
  return NULL;
}

int main(int argc, char **argv) {
  g_world = argc > 1 ? atoi(argv[1]) : 2;
  if (g_world < 2 || g_world > MAXW || N % g_world) {
    printf("n_gpus must be in [2, %d] and divide %d\n", MAXW, N);
    return 2;
  }
  /* inputs */
  uint32_t seed = 12345;
  g_x = malloc((size_t)N * D * sizeof *g_x);
  g_ref = malloc((size_t)N * H * sizeof *g_ref);
  for (size_t i = 0; i < (size_t)N * D; ++i) g_x[i] = frand(&seed);
  for (int r = 0; r < N; ++r) { /* ring neighbours + far chords: every rank needs remote rows */
    g_rp[r] = (uint64_t)r * DEG;
    for (int k = 0; k < DEG; ++k) {
      const int off = k < DEG / 2 ? k + 1 : (k - DEG / 2 + 1) * (N / 7) + k;
      g_col[r * DEG + k] = (uint32_t)((r + off) % N);
      g_w[r * DEG + k] = 0.05f * (float)(k + 1);
    }
  }
  g_rp[N] = (uint64_t)N * DEG;
  static float wbuf[12][H * D], bbuf[12][H], gamma[H], beta[H];
  ruvgnn_linear *lin[12] = {&g_W.w_msg, &g_W.w_agg, &g_W.w_z, &g_W.u_z, &g_W.w_r, &g_W.u_r,
                            &g_W.w_h,   &g_W.u_h,   &g_W.q,   &g_W.k,   &g_W.v,   &g_W.out};
  for (int l = 0; l < 12; ++l) {
    for (int i = 0; i < H * D; ++i) wbuf[l][i] = 0.25f * frand(&seed);
    for (int i = 0; i < H; ++i) bbuf[l][i] = 0.1f * frand(&seed);
    lin[l]->w = wbuf[l];
    lin[l]->b = bbuf[l];
  }
  for (int i = 0; i < H; ++i) { gamma[i] = 1.0f + 0.5f * frand(&seed); beta[i] = 0.1f * frand(&seed); }
  g_W.gamma = gamma; g_W.beta = beta; g_W.eps = 1e-5f;

  for (int r = 0; r < g_world; ++r) {
    if (ruvgnn_ctx_create(r, NULL, &g_ctx[r]) != RUVGNN_OK) {
      printf("device %d unavailable: %s\nSKIPPED\n", r, ruvgnn_last_error());
      return 77;
    }
  }
  if (ruvgnn_dist_create_all(g_ctx, g_world, g_dist) != RUVGNN_OK) {
    printf("ruvgnn_dist_create_all: %s\nFAILED\n", ruvgnn_last_error());
    return 1;
  }
  pthread_barrier_init(&g_bar, NULL, (unsigned)g_world);
  pthread_t th[MAXW];
  for (int r = 0; r < g_world; ++r) pthread_create(&th[r], NULL, rank_main, (void *)(size_t)r);
  int failures = 0;
  for (int r = 0; r < g_world; ++r) {
    pthread_join(th[r], NULL);
    failures += g_fail[r];
  }
  for (int r = 0; r < g_world; ++r) {
    ruvgnn_dist_destroy(g_dist[r]);
    ruvgnn_ctx_destroy(g_ctx[r]);
  }
  printf(failures ? "FAILED (%d)\n" : "OK\n", failures);
  return failures ? 1 : 0;
}
