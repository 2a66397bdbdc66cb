/* Plain C99 caller of the C ABI (what a cgo / Rust `extern "C"` / JNI binding links against):
 * constructor validation with the reference's messages, the neighbour-list packer, JSON
 * normalisation and -- without a device -- the loud CUDA error of a compute entry.
 *   ./test_abi            host-only checks
 *   ./test_abi gpu        also one per-node forward (layer.rs:512-521 shape) */
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "ruvgnn.h"

static int failures = 0;
#define CHECK(cond)                                                        \
  do {                                                                     \
    if (!(cond)) {                                                         \
      printf("  CHECK FAILED %s:%d: %s\n", __FILE__, __LINE__, #cond);     \
      ++failures;                                                          \
    }                                                                      \
  } while (0)

/* hnsw_rs `Neighbour` (#[repr(C)], patches/hnsw_rs/src/hnsw.rs:97-106): d_id usize, distance f32, p_id (u8, i32) */
typedef struct {
  uint64_t d_id;
  float distance;
  uint8_t layer;
  int32_t rank;
} neighbour;

int main(int argc, char **argv) {
  const int gpu = argc > 1 && strcmp(argv[1], "gpu") == 0;

  /* RuvectorLayer::new validation (layer.rs:353-358, :101-106) */
  CHECK(ruvgnn_layer_validate(4, 8, 2, 0.1f) == RUVGNN_OK);
  CHECK(ruvgnn_layer_validate(4, 8, 2, 1.5f) == RUVGNN_ERR_LAYER_CONFIG);
  CHECK(strstr(ruvgnn_last_error(), "Dropout must be between 0.0 and 1.0, got 1.5") != NULL);
  CHECK(ruvgnn_layer_validate(4, 7, 3, 0.1f) == RUVGNN_ERR_LAYER_CONFIG);
  CHECK(strstr(ruvgnn_last_error(), "must be divisible by number of heads") != NULL);

  /* neighbour lists -> CSR, bit-exact and in stored order */
  neighbour l0[2] = {{2, 0.25f, 0, 0}, {1, 0.5f, 0, 1}};
  neighbour l2[1] = {{0, 0.125f, 0, 2}};
  const void *lists[3] = {l0, NULL, l2};
  uint64_t counts[3] = {2, 0, 1}, row_ptr[4];
  uint32_t col[3];
  float w[3];
  CHECK(ruvgnn_pack_neighbours(3, lists, counts, sizeof(neighbour), offsetof(neighbour, d_id),
                               offsetof(neighbour, distance), row_ptr, col, w) == RUVGNN_OK);
  CHECK(row_ptr[0] == 0 && row_ptr[1] == 2 && row_ptr[2] == 2 && row_ptr[3] == 3);
  CHECK(col[0] == 2 && col[1] == 1 && col[2] == 0 && w[0] == 0.25f && w[1] == 0.5f && w[2] == 0.125f);
  l2[0].d_id = 3; /* out of range */
  CHECK(ruvgnn_pack_neighbours(3, lists, counts, sizeof(neighbour), offsetof(neighbour, d_id),
                               offsetof(neighbour, distance), row_ptr, col, w) == RUVGNN_ERR_INVALID_INPUT);

  /* a truncated layer JSON is a deserialisation error, not a crash */
  const char *bad = "{\"w_msg\":{\"weights\":{\"v\":1,\"dim\":[2,2],\"data\":[1,2,3";
  size_t need = 0;
  CHECK(ruvgnn_layer_json_normalize(bad, strlen(bad), NULL, 0, &need) == RUVGNN_ERR_INVALID_INPUT);

  ruvgnn_ctx *ctx = NULL;
  const int32_t st = ruvgnn_ctx_create(0, NULL, &ctx);
  if (!gpu) {
    if (st != RUVGNN_OK) {
      CHECK(st == RUVGNN_ERR_CUDA && ctx == NULL);
      CHECK(strstr(ruvgnn_last_error(), "no CPU fallback") != NULL);
    }
  } else {
    CHECK(st == RUVGNN_OK && ctx != NULL);
    /* identity-free smoke: random-ish weights, output length = hidden_dim and finite (layer.rs:512-521) */
    enum { D = 4, H = 8 };
    static float wm[H * D], wh[H * H], bias[H], gamma[H], beta[H];
    for (int i = 0; i < H * D; ++i) wm[i] = 0.1f * (float)((i * 7) % 5 - 2);
    for (int i = 0; i < H * H; ++i) wh[i] = 0.05f * (float)((i * 11) % 7 - 3);
    for (int i = 0; i < H; ++i) { bias[i] = 0.01f * (float)i; gamma[i] = 1.f; beta[i] = 0.f; }
    ruvgnn_weights W;
    memset(&W, 0, sizeof W);
    ruvgnn_linear msg = {wm, bias}, hh = {wh, bias};
    W.w_msg = msg; W.w_agg = hh; W.w_z = hh; W.u_z = hh; W.w_r = hh; W.u_r = hh; W.w_h = hh; W.u_h = hh;
    W.q = hh; W.k = hh; W.v = hh; W.out = hh;
    W.gamma = gamma; W.beta = beta; W.eps = 1e-5f;
    ruvgnn_layer *layer = NULL;
    CHECK(ruvgnn_layer_create(ctx, D, H, 2, 0.1f, &W, &layer) == RUVGNN_OK);
    const float node[D] = {1.f, 2.f, 3.f, 4.f};
    const float nbrs[2 * D] = {0.5f, 1.f, 1.5f, 2.f, 2.f, 3.f, 4.f, 5.f};
    const float ew[2] = {0.3f, 0.7f};
    float out[H];
    CHECK(ruvgnn_layer_forward_node(ctx, layer, node, D, nbrs, 2, ew, 2, out) == RUVGNN_OK);
    float mean = 0.f;
    for (int i = 0; i < H; ++i) { CHECK(isfinite(out[i])); mean += out[i]; }
    CHECK(fabsf(mean / H) < 1e-5f); /* LayerNorm(gamma = 1, beta = 0) output is zero-mean */
    CHECK(ruvgnn_layer_forward_node(ctx, layer, node, D - 1, nbrs, 2, ew, 2, out) == RUVGNN_ERR_DIMENSION_MISMATCH);
    ruvgnn_layer_destroy(layer);
  }
  if (ctx) ruvgnn_ctx_destroy(ctx);
  printf(failures ? "FAILED (%d)\n" : "OK\n", failures);
  return failures ? 1 : 0;
}
