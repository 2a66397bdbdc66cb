// tcgen05 projection path -- placeholder translation unit until the kernel lands:
// reports the path as unusable so ruvgnn_layer_set_mode() refuses tensor-core modes.
#include "gemm_tc.cuh"

namespace ruvgnn {

cudaError_t tc_weights_create(const float *const *, const uint32_t *, const uint32_t *,
                              TcWeights *out) {
  out->usable = false;
  return cudaSuccess;
}
void tc_weights_destroy(TcWeights *) {}
cudaError_t launch_gemm_tc(const GemmArgs &, EpiKind, const TcWeights &, int, bool, int,
                           cudaStream_t, uint64_t *) {
  return cudaErrorNotSupported;
}

}  // namespace ruvgnn
