// ruvgnn.hpp -- header-only C++ host-side mirror of ruvector_gnn::layer::RuvectorLayer over the
// C ABI in ruvgnn.h (the reference is Rust; no Rust toolchain exists in the build image, so the
// compiled-language host side is C++17).  Same names, argument meaning and error behaviour as
// crates/ruvector-gnn/src/layer.rs:315-457 and error.rs:9-60:
//
//   ruvgnn::RuvectorLayer layer(4, 8, 2, 0.1f);                 // throws GnnError (LayerConfig)
//   std::vector<float> out = layer.forward(node, neighbors, w);  // per node (layer.rs:379-384)
//   std::string json = layer.to_json();                          // serde shape
//   auto again = ruvgnn::RuvectorLayer::from_json(json);
//   std::vector<float> all = layer.forward_graph(x, n, row_ptr, col_idx, edge_w);  // batched
//
// All arithmetic runs in libruvgnn.so on the GPU; there is no CPU path.
#ifndef RUVGNN_HPP
#define RUVGNN_HPP

#include <cmath>
#include <cstdint>
#include <memory>
#include <random>
#include <stdexcept>
#include <string>
#include <vector>

#include "ruvgnn.h"

namespace ruvgnn {

// GnnError (error.rs:9-60): what() carries the reference's #[error(...)] text.
class GnnError : public std::runtime_error {
 public:
  enum Kind { DimensionMismatch = 1, InvalidShape = 2, LayerConfig = 3, InvalidInput = 7, Cuda = 100, Nccl = 101, Oom = 102, Other = -1 };
  GnnError(Kind k, const std::string &msg) : std::runtime_error(prefix(k) + msg), kind_(k) {}
  Kind kind() const { return kind_; }

 private:
  static std::string prefix(Kind k) {
    switch (k) {
      case InvalidShape: return "Invalid tensor shape: ";
      case LayerConfig: return "Layer configuration error: ";
      case InvalidInput: return "Invalid input: ";
      default: return "";
    }
  }
  Kind kind_;
};

inline void check(int32_t status) {
  if (status == RUVGNN_OK) return;
  const char *m = ruvgnn_last_error();
  GnnError::Kind k = GnnError::Other;
  switch (status) {
    case 1: k = GnnError::DimensionMismatch; break;
    case 2: k = GnnError::InvalidShape; break;
    case 3: k = GnnError::LayerConfig; break;
    case 7: k = GnnError::InvalidInput; break;
    case 100: k = GnnError::Cuda; break;
    case 101: k = GnnError::Nccl; break;
    case 102: k = GnnError::Oom; break;
  }
  throw GnnError(k, m ? m : "");
}

// One process-wide context per device, created on first use.
inline ruvgnn_ctx *default_context(int device = 0) {
  static ruvgnn_ctx *ctx[16] = {};
  if (device < 0 || device >= 16) throw GnnError(GnnError::InvalidInput, "device out of range");
  if (!ctx[device]) check(ruvgnn_ctx_create(device, nullptr, &ctx[device]));
  return ctx[device];
}

// One Linear (layer.rs:14-17): weights row-major [out][in], bias [out].
struct Linear {
  std::vector<float> weights, bias;
  uint32_t output_dim = 0, input_dim = 0;
  Linear() = default;
  // Linear::new (layer.rs:21-34): Xavier normal, std = sqrt(2 / (in + out)), zero bias
  Linear(uint32_t in, uint32_t out, std::mt19937 &rng) : weights((size_t)in * out), bias(out, 0.f), output_dim(out), input_dim(in) {
    std::normal_distribution<double> normal(0.0, std::sqrt(2.0f / (float)(in + out)));
    for (float &w : weights) w = (float)normal(rng);
  }
};

class RuvectorLayer {
 public:
  // RuvectorLayer::new(input_dim, hidden_dim, heads, dropout) (layer.rs:347-368).
  RuvectorLayer(uint32_t input_dim, uint32_t hidden_dim, uint32_t heads, float dropout, uint64_t seed = std::random_device{}())
      : input_dim_(input_dim), hidden_dim_(hidden_dim), heads_(heads), dropout_(dropout) {
    check(ruvgnn_layer_validate(input_dim, hidden_dim, heads, dropout));
    std::mt19937 rng((uint32_t)seed);
    w_msg = Linear(input_dim, hidden_dim, rng);
    w_agg = Linear(hidden_dim, hidden_dim, rng);
    for (Linear *l : {&w_z, &u_z, &w_r, &u_r, &w_h, &u_h, &q_linear, &k_linear, &v_linear, &out_linear})
      *l = Linear(hidden_dim, hidden_dim, rng);
    gamma.assign(hidden_dim, 1.f);  // LayerNorm::new (layer.rs:59-64)
    beta.assign(hidden_dim, 0.f);
  }
  RuvectorLayer(const RuvectorLayer &) = delete;
  RuvectorLayer &operator=(const RuvectorLayer &) = delete;
  RuvectorLayer(RuvectorLayer &&o) noexcept { *this = std::move(o); }
  RuvectorLayer &operator=(RuvectorLayer &&o) noexcept {
    release();
    input_dim_ = o.input_dim_; hidden_dim_ = o.hidden_dim_; heads_ = o.heads_; dropout_ = o.dropout_; eps = o.eps;
    w_msg = std::move(o.w_msg); w_agg = std::move(o.w_agg); w_z = std::move(o.w_z); u_z = std::move(o.u_z);
    w_r = std::move(o.w_r); u_r = std::move(o.u_r); w_h = std::move(o.w_h); u_h = std::move(o.u_h);
    q_linear = std::move(o.q_linear); k_linear = std::move(o.k_linear); v_linear = std::move(o.v_linear);
    out_linear = std::move(o.out_linear); gamma = std::move(o.gamma); beta = std::move(o.beta);
    handle_ = o.handle_; o.handle_ = nullptr; device_ = o.device_;
    return *this;
  }
  ~RuvectorLayer() { release(); }

  // serde_json::from_str::<RuvectorLayer> (ruvector-gnn-node/src/lib.rs:108-116)
  static RuvectorLayer from_json(const std::string &json, int device = 0) {
    RuvectorLayer l;
    l.device_ = device;
    check(ruvgnn_layer_from_json(default_context(device), json.data(), json.size(), &l.handle_));
    check(ruvgnn_layer_dims(l.handle_, &l.input_dim_, &l.hidden_dim_, &l.heads_, &l.dropout_));
    l.from_json_only_ = true;
    return l;
  }
  // serde_json::to_string(&layer) (lib.rs:97-105)
  std::string to_json() {
    size_t need = 0;
    check(ruvgnn_layer_to_json(device_layer(), nullptr, 0, &need));
    std::string s(need + 1, '\0');
    check(ruvgnn_layer_to_json(device_layer(), &s[0], need + 1, &need));
    s.resize(need);
    return s;
  }

  // forward(&self, node_embedding, neighbor_embeddings, edge_weights) -> Vec<f32> (layer.rs:379-384)
  std::vector<float> forward(const std::vector<float> &node_embedding,
                             const std::vector<std::vector<float>> &neighbor_embeddings,
                             const std::vector<float> &edge_weights) {
    std::vector<float> flat;
    flat.reserve(neighbor_embeddings.size() * node_embedding.size());
    for (const auto &n : neighbor_embeddings) {
      if (n.size() != node_embedding.size())
        throw GnnError(GnnError::DimensionMismatch, "Tensor dimension mismatch: expected " +
                                                        std::to_string(node_embedding.size()) + ", got " + std::to_string(n.size()));
      flat.insert(flat.end(), n.begin(), n.end());
    }
    std::vector<float> out(hidden_dim_);
    check(ruvgnn_layer_forward_node(default_context(device_), device_layer(), node_embedding.data(),
                                    (uint32_t)node_embedding.size(), flat.data(), (uint32_t)neighbor_embeddings.size(),
                                    edge_weights.data(), (uint32_t)edge_weights.size(), out.data()));
    return out;
  }

  // Batched: out[i] = forward(x[i], [x[j] for j in N(i)], w[i]) for all i (the callers' loop,
  // verified_training.rs:422-436).  x is n x input_dim row-major; CSR as packed by ruvgnn_pack_neighbours.
  std::vector<float> forward_graph(const std::vector<float> &x, uint64_t n, const std::vector<uint64_t> &row_ptr,
                                   const std::vector<uint32_t> &col_idx, const std::vector<float> &edge_w) {
    if (x.size() != n * input_dim_ || row_ptr.size() != n + 1)
      throw GnnError(GnnError::DimensionMismatch, "Tensor dimension mismatch: expected " + std::to_string(n * input_dim_) +
                                                      ", got " + std::to_string(x.size()));
    ruvgnn_graph *g = nullptr;
    check(ruvgnn_graph_create(default_context(device_), n, n, row_ptr.data(), col_idx.data(), edge_w.data(), 0, &g));
    std::vector<float> out(n * hidden_dim_);
    const int32_t st = ruvgnn_layer_forward(default_context(device_), device_layer(), g, x.data(), out.data(), 0);
    ruvgnn_graph_destroy(g);
    check(st);
    return out;
  }

  uint32_t input_dim() const { return input_dim_; }
  uint32_t hidden_dim() const { return hidden_dim_; }
  uint32_t heads() const { return heads_; }
  float dropout() const { return dropout_; }

  // weights (public here so tests can inject them; the reference keeps them private + serde)
  Linear w_msg, w_agg, w_z, u_z, w_r, u_r, w_h, u_h, q_linear, k_linear, v_linear, out_linear;
  std::vector<float> gamma, beta;
  float eps = 1e-5f;

 private:
  RuvectorLayer() = default;
  void release() {
    if (handle_) ruvgnn_layer_destroy(handle_);
    handle_ = nullptr;
  }
  ruvgnn_layer *device_layer() {
    if (!handle_) {
      if (from_json_only_) throw GnnError(GnnError::Other, "layer handle lost");
      ruvgnn_weights w{};
      auto lin = [](const Linear &l) { return ruvgnn_linear{l.weights.data(), l.bias.data()}; };
      w.w_msg = lin(w_msg); w.w_agg = lin(w_agg); w.w_z = lin(w_z); w.u_z = lin(u_z); w.w_r = lin(w_r);
      w.u_r = lin(u_r); w.w_h = lin(w_h); w.u_h = lin(u_h); w.q = lin(q_linear); w.k = lin(k_linear);
      w.v = lin(v_linear); w.out = lin(out_linear);
      w.gamma = gamma.data(); w.beta = beta.data(); w.eps = eps;
      check(ruvgnn_layer_create(default_context(device_), input_dim_, hidden_dim_, heads_, dropout_, &w, &handle_));
    }
    return handle_;
  }
  uint32_t input_dim_ = 0, hidden_dim_ = 0, heads_ = 0;
  float dropout_ = 0.f;
  ruvgnn_layer *handle_ = nullptr;
  int device_ = 0;
  bool from_json_only_ = false;
};

}  // namespace ruvgnn
#endif  // RUVGNN_HPP
