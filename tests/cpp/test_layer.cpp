// C++ host-side tests written to read like the reference's own unit tests
// (crates/ruvector-gnn/src/layer.rs:459-546, crates/ruvector-gnn-node/test/basic.test.js:54-73).
//   ./test_layer config   -> constructor / error tests only (no GPU needed)
//   ./test_layer all      -> also the forward tests (needs a B200)
#include <cmath>
#include <cstdio>
#include <cstring>
#include <string>

#include "ruvgnn.hpp"

using ruvgnn::GnnError;
using ruvgnn::RuvectorLayer;

static int failures = 0;
#define ASSERT(cond)                                                  \
  do {                                                                \
    if (!(cond)) {                                                    \
      std::printf("  ASSERT FAILED %s:%d: %s\n", __FILE__, __LINE__, #cond); \
      ++failures;                                                     \
    }                                                                 \
  } while (0)

static void test_ruvector_layer_invalid_dropout() {  // layer.rs:536-539
  bool err = false;
  try {
    RuvectorLayer layer(4, 8, 2, 1.5f);
  } catch (const GnnError &e) {
    err = e.kind() == GnnError::LayerConfig &&
          std::string(e.what()) == "Layer configuration error: Dropout must be between 0.0 and 1.0, got 1.5";
  }
  ASSERT(err);
}

static void test_ruvector_layer_invalid_heads() {  // layer.rs:541-545, :494-499
  bool err = false;
  try {
    RuvectorLayer layer(4, 7, 3, 0.1f);
  } catch (const GnnError &e) {
    err = e.kind() == GnnError::LayerConfig && std::strstr(e.what(), "divisible") != nullptr;
  }
  ASSERT(err);
}

static void test_xavier_init() {  // layer.rs:21-34, 59-64
  RuvectorLayer layer(128, 256, 4, 0.1f, 7);
  ASSERT(layer.w_msg.weights.size() == 256u * 128u && layer.w_msg.bias.size() == 256u);
  double s2 = 0;
  for (float w : layer.w_msg.weights) s2 += (double)w * w;
  const double sd = std::sqrt(s2 / layer.w_msg.weights.size());
  ASSERT(std::fabs(sd - std::sqrt(2.0 / (128 + 256))) < 2e-3);
  for (float b : layer.w_msg.bias) ASSERT(b == 0.f);
  ASSERT(layer.gamma[0] == 1.f && layer.beta[0] == 0.f && layer.eps == 1e-5f);
}

static void test_ruvector_layer() {  // layer.rs:512-521
  RuvectorLayer layer(4, 8, 2, 0.1f, 1);
  std::vector<float> node = {1.0f, 2.0f, 3.0f, 4.0f};
  std::vector<std::vector<float>> neighbors = {{0.5f, 1.0f, 1.5f, 2.0f}, {2.0f, 3.0f, 4.0f, 5.0f}};
  std::vector<float> weights = {0.3f, 0.7f};
  std::vector<float> output = layer.forward(node, neighbors, weights);
  ASSERT(output.size() == 8u);
  for (float v : output) ASSERT(std::isfinite(v));
}

static void test_ruvector_layer_no_neighbors() {  // layer.rs:523-533
  RuvectorLayer layer(4, 8, 2, 0.1f, 2);
  std::vector<float> output = layer.forward({1.0f, 2.0f, 3.0f, 4.0f}, {}, {});
  ASSERT(output.size() == 8u);
  float mean = 0;
  for (float v : output) mean += v;
  ASSERT(std::fabs(mean / 8) < 1e-5f);  // LayerNorm(gamma = 1, beta = 0) output is zero-mean (:472-480)
}

static void test_json_roundtrip_same_forward() {  // basic.test.js:54-73
  RuvectorLayer layer(4, 8, 2, 0.1f, 3);
  std::vector<float> node = {1.0f, 2.0f, 3.0f, 4.0f};
  std::vector<std::vector<float>> neighbors = {{0.5f, 1.0f, 1.5f, 2.0f}, {2.0f, 3.0f, 4.0f, 5.0f}};
  std::vector<float> weights = {0.3f, 0.7f};
  std::vector<float> a = layer.forward(node, neighbors, weights);
  RuvectorLayer again = RuvectorLayer::from_json(layer.to_json());
  std::vector<float> b = again.forward(node, neighbors, weights);
  ASSERT(a.size() == b.size());
  for (size_t i = 0; i < a.size(); ++i) ASSERT(std::fabs(a[i] - b[i]) < 1e-6f);
}

static void test_forward_graph_equals_per_node_calls() {  // verified_training.rs:422-436
  RuvectorLayer layer(8, 16, 4, 0.1f, 4);
  const uint64_t n = 5;
  std::vector<float> x(n * 8);
  for (size_t i = 0; i < x.size(); ++i) x[i] = std::sin(0.37f * (float)i);
  std::vector<uint64_t> row_ptr = {0, 2, 2, 5, 6, 8};
  std::vector<uint32_t> col = {1, 4, 0, 1, 3, 2, 0, 3};
  std::vector<float> w = {0.2f, 0.8f, 0.1f, 0.5f, 0.9f, 1.0f, 0.3f, 0.3f};
  std::vector<float> all = layer.forward_graph(x, n, row_ptr, col, w);
  for (uint64_t i = 0; i < n; ++i) {
    std::vector<std::vector<float>> nb;
    std::vector<float> ew;
    for (uint64_t e = row_ptr[i]; e < row_ptr[i + 1]; ++e) {
      nb.emplace_back(x.begin() + col[e] * 8, x.begin() + (col[e] + 1) * 8);
      ew.push_back(w[e]);
    }
    std::vector<float> one = layer.forward(std::vector<float>(x.begin() + i * 8, x.begin() + (i + 1) * 8), nb, ew);
    for (uint32_t t = 0; t < 16; ++t) ASSERT(std::fabs(one[t] - all[i * 16 + t]) < 1e-5f);
  }
  bool err = false;
  try {
    layer.forward({1.f, 2.f}, {}, {});
  } catch (const GnnError &e) {
    err = e.kind() == GnnError::DimensionMismatch;
  }
  ASSERT(err);
}

int main(int argc, char **argv) {
  const bool gpu = argc > 1 && std::string(argv[1]) == "all";
  test_ruvector_layer_invalid_dropout();
  test_ruvector_layer_invalid_heads();
  test_xavier_init();
  if (gpu) {
    test_ruvector_layer();
    test_ruvector_layer_no_neighbors();
    test_json_roundtrip_same_forward();
    test_forward_graph_equals_per_node_calls();
  } else {
    // without a device the compute entry must fail loudly, never fall back
    bool err = false;
    try {
      RuvectorLayer layer(4, 8, 2, 0.1f, 1);
      layer.forward({1.f, 2.f, 3.f, 4.f}, {}, {});
    } catch (const GnnError &e) {
      err = e.kind() == GnnError::Cuda;
    }
    int ndev = 0;
    (void)ndev;
    std::printf("no-GPU forward raised Cuda error: %s\n", err ? "yes" : "no (a GPU is present?)");
  }
  std::printf(failures ? "FAILED (%d)\n" : "OK\n", failures);
  return failures ? 1 : 0;
}
