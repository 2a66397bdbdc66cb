// tcgen05 (5th-gen tensor core) projection path: interface.
#pragma once
#include "common.cuh"

namespace ruvgnn {

// The six device weight matrices of a layer, in the order upload_layer builds them.
// TC_W_KV = rows [H, 3H) of the Q|K|V matrix: rebuilds K|V of received message rows (M-only halo)
// TC_W_F  = [W_agg W_o | W_agg] resp. [W_agg W_ov | W_agg]: first stage of the fused K-upd (H = 128 only;
//           rows == 0 elsewhere, such an entry is skipped)
enum TcWeightId : int {
  TC_W_MSG = 0, TC_W_QKV, TC_W_O, TC_W_AGG, TC_W_ZR, TC_W_H, TC_W_KV, TC_W_F, TC_W_COUNT
};

// Per matrix: W_hi (TF32-representable part) and W_lo = W - W_hi, both [rows][cols]
// row-major on the device, plus the TMA descriptors that stream them.
struct TcWeights {
  bool usable = false;
  float *blob = nullptr;
  float *hi[TC_W_COUNT] = {};
  float *lo[TC_W_COUNT] = {};
  uint32_t rows[TC_W_COUNT] = {};
  uint32_t cols[TC_W_COUNT] = {};
  void *tmaps = nullptr;  // host array of CUtensorMap, 2 per matrix (hi, lo)
};

// host_mats[i]: the (already permuted) host image of matrix i.
cudaError_t tc_weights_create(const float *const host_mats[TC_W_COUNT],
                              const uint32_t rows[TC_W_COUNT], const uint32_t cols[TC_W_COUNT],
                              TcWeights *out);
void tc_weights_destroy(TcWeights *w);

// Same contract as launch_gemm_simt; `which` selects the weight matrix (g.W is ignored).
// single_pass: one TF32 MMA per product (RUVGNN_MODE_TF32) instead of the 3xTF32 split.
cudaError_t launch_gemm_tc(const GemmArgs &g, EpiKind epi, const TcWeights &w, int which,
                           bool single_pass, int sm_count, cudaStream_t s, uint64_t *launches);

// Fused K-upd (gemm_upd_fused.cu): is the kernel usable with these weights, and its launcher.
namespace tc {
bool upd_fused_supported(const TcWeights &w, uint32_t H);
cudaError_t launch_upd_fused(const FusedUpdArgs &a, const TcWeights &w, bool single_pass, int sm_count,
                             cudaStream_t s, uint64_t *launches);
}  // namespace tc

}  // namespace ruvgnn
