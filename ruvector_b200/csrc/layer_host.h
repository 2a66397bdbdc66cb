// Host-side image of one RuvectorLayer in the reference's own (logical) layout,
// plus its serde-JSON (de)serialisation.  Field order follows the serde
// declaration order of crates/ruvector-gnn/src/layer.rs:315-333.
#pragma once
#include <stdint.h>

#include <string>
#include <vector>

namespace ruvgnn {

enum LinearId : int {
  L_W_MSG = 0, L_W_AGG, L_W_Z, L_U_Z, L_W_R, L_U_R, L_W_H, L_U_H, L_Q, L_K, L_V, L_OUT, L_COUNT
};

struct HostLinear {
  uint32_t out = 0, in = 0;
  std::vector<float> w;  // row-major [out][in]
  std::vector<float> b;  // [out]
};

struct LayerHost {
  uint32_t input_dim = 0, hidden_dim = 0, heads = 0;
  float dropout = 0.f;
  float eps = 1e-5f;
  HostLinear lin[L_COUNT];
  std::vector<float> gamma, beta;
};

// Returns "" on success, otherwise a message (InvalidInput).
std::string parse_layer_json(const char *json, size_t len, LayerHost &out);
std::string emit_layer_json(const LayerHost &l);

}  // namespace ruvgnn
