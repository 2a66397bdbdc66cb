// Shared declarations for libruvgnn.so (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include <string>

#include "../../include/ruvgnn.h"

namespace ruvgnn {

// ---- error plumbing ------------------------------------------------------
void set_error(const std::string &msg);
int32_t fail(int32_t code, const std::string &msg);

#define RUVGNN_CUDA_TRY(expr)                                                             \
  do {                                                                                    \
    cudaError_t _e = (expr);                                                              \
    if (_e != cudaSuccess) {                                                              \
      return ::ruvgnn::fail(_e == cudaErrorMemoryAllocation ? RUVGNN_ERR_OOM              \
                                                            : RUVGNN_ERR_CUDA,            \
                            std::string(#expr) + ": " + cudaGetErrorString(_e));          \
    }                                                                                     \
  } while (0)

// ---- the internal column order ("phys") ----------------------------------
// Every N x H activation the gather kernel touches (Q, K, V, M, o, agg and
// everything downstream up to the LayerNorm) is stored with its columns
// permuted so that lane l of a warp owns the logical floats [4P*l, 4P*(l+1))
// -- i.e. whole heads map to contiguous lane groups -- while each warp-wide
// 128-bit load still reads 512 contiguous bytes:
//     logical float4 index f = l*P + i   <->   physical float4 index i*32 + l
// with P = H/128.  The permutation is folded into the weights once at layer
// creation, so it costs nothing at run time.  When H % 128 != 0 (or the head
// count does not divide 32) the order is the identity and the generic kernels run.
struct ColumnOrder {
  uint32_t H = 0, P = 0;  // P == 0: identity
  __host__ __device__ inline uint32_t phys(uint32_t c) const {
    if (P == 0) return c;
    uint32_t per_lane = 4 * P;
    uint32_t lane = c / per_lane, r = c % per_lane;
    return ((r >> 2) * 32 + lane) * 4 + (r & 3);
  }
};

inline bool fast_layout_ok(uint32_t H, uint32_t heads) {
  if (H % 128 != 0 || heads == 0 || heads > 32 || (32 % heads) != 0) return false;
  uint32_t P = H / 128;
  return P == 1 || P == 2 || P == 3 || P == 4 || P == 6 || P == 8;
}

// ---- GEMM: C[M x N] = [A1 | A2] (M x (K1+K2)) * W^T (W is [N x (K1+K2)]) + epilogue
enum EpiKind : int {
  EPI_BIAS = 0,      // out = acc + bias
  EPI_BIAS_ADD = 1,  // out = acc + bias + add
  EPI_GATES = 2,     // n <  H: z  = sigmoid(acc+bias)             -> z[row][n]
                     // n >= H: rh = sigmoid(acc+bias) * m[row][n-H] -> rh[row][n-H]
  EPI_GRU = 3        // ht = tanh(acc+bias); out = ((1-z)*m + z*ht) * keep
};

struct GemmArgs {
  int64_t M = 0;
  int32_t N = 0, K1 = 0, K2 = 0;
  const float *A1 = nullptr;
  int64_t lda1 = 0;
  const float *A2 = nullptr;
  int64_t lda2 = 0;
  const float *W = nullptr;  // [N][K1+K2], row stride K1+K2
  const float *bias = nullptr;
  float *out = nullptr;
  int64_t ldo = 0;
  const float *add = nullptr;  // EPI_BIAS_ADD
  int64_t ldadd = 0;
  const float *m = nullptr;  // EPI_GATES / EPI_GRU: hidden state
  int64_t ldm = 0;
  float *z = nullptr;  // EPI_GATES: written; EPI_GRU: read
  int64_t ldz = 0;
  float *rh = nullptr;  // EPI_GATES: written
  int64_t ldrh = 0;
  int32_t H = 0;
  float keep = 1.f;  // 1 - dropout
};

cudaError_t launch_gemm_simt(const GemmArgs &a, EpiKind epi, cudaStream_t s, uint64_t *launches);

// ---- gather / aggregate (K-agg) -------------------------------------------
// Peer-direct plane of the partitioned forward: col_idx holds GLOBAL node ids; a column inside
// [bounds[rank], bounds[rank+1]) is an own row, any other is read straight from its owner's
// projected table over NVLink (peer memory mapped with CUDA IPC): the transfer is the gather's own
// loads, there is no halo buffer and no separate collective to wait for.
constexpr int kMaxPeers = 16;
struct PeerTable {
  const float *base[kMaxPeers] = {};  // peer p's projected table (same layout and row stride as ours)
  uint64_t bounds[kMaxPeers + 1] = {};
  int world = 0, rank = 0;
};

struct GatherArgs {
  const float *P = nullptr;  // [n_local][ldp]: Q | K | V | M (phys order), each H floats
  int64_t ldp = 0;
  const float *halo = nullptr;  // [n_halo][3H]: K | V | M of remote rows
  uint64_t n_local = 0;
  const uint64_t *row_ptr = nullptr;
  const uint32_t *col_idx = nullptr;
  const float *edge_w = nullptr;
  float *oa = nullptr;  // [n_local][2H]: o | agg
  uint64_t row_begin = 0, row_end = 0;
  uint32_t H = 0, heads = 0;
  int prenormalized = 0;
  // Multi-pass gather (partitioned forward, dist.cu): col_idx / edge_w are the pass-ordered copy of
  // the edge arrays (launch_build_pass_csr; weights already normalised) and this launch handles
  // group `pass` of every row, edges row_ptr[row] + [seg[row][pass], seg[row][pass+1]).  first: the
  // state starts from zero; otherwise it is read back from oa (raw accumulators o | agg) and ml
  // ([row][2*heads]: running max, running sum per head).  last: oa receives the normalised o | agg
  // that K-upd reads; otherwise the raw state.
  int multipass = 0, first = 1, last = 1;
  const uint32_t *seg = nullptr;  // [rows][npass1]
  int npass1 = 0, pass = 0;
  float *ml = nullptr;
  const PeerTable *peer = nullptr;  // device pointer; non-null: peer-direct plane (col_idx = global ids)
};
cudaError_t launch_gather(const GatherArgs &a, bool fast, int sm_count, cudaStream_t s,
                          uint64_t *launches);

// Single-row (R = 1) formulation, H = 128: per edge only the 512-byte message row M_j is read.
//   scores  s_gj = (W_k[g]^T q_g) . m_j / sqrt(d_h)   (the q_g . b_k[g] term is constant in j and
//                                                     cancels in the softmax)
//   values  o_g  = W_v[g] (sum_j a_gj m_j) + b_v[g]   (sum_j a_gj = 1)
// so the kernel reads the folded queries Qt [n][heads*H], accumulates S_g = sum_j a_gj m_j per
// head and writes [S_0 | .. | S_{h-1} | agg]; W_k / W_v are folded into the GEMMs around it.
struct GatherR1Args {
  const float *M = nullptr;  // [n_local][ldm] message rows
  int64_t ldm = 0;
  const float *halo = nullptr;  // [n_halo][128] remote message rows
  uint64_t n_local = 0;
  const float *Qt = nullptr;  // [n_local][heads*128]
  const uint64_t *row_ptr = nullptr;
  const uint32_t *col_idx = nullptr;
  const float *edge_w = nullptr;
  float *sa = nullptr;  // [n_local][(heads+1)*128]: S_0..S_{h-1} | agg
  uint64_t row_begin = 0, row_end = 0;
  uint32_t heads = 0;
  int prenormalized = 0;
  // multi-pass gather, as in GatherArgs (state: sa raw S_g | agg, ml [row][2*heads])
  int multipass = 0, first = 1, last = 1;
  const uint32_t *seg = nullptr;
  int npass1 = 0, pass = 0;
  float *ml = nullptr;
  const PeerTable *peer = nullptr;  // as in GatherArgs
};
cudaError_t launch_gather_r1(const GatherR1Args &a, int sm_count, cudaStream_t s, uint64_t *launches);
// Pass-ordered copy of a CSR's edge arrays for the multi-pass gather: pass p = the columns in
// [bounds[p], bounds[p+1]) (bounds_dev: npass + 1 ascending device values covering every column);
// col_p / w_p: [n_edges] (w_p normalised by the aggregate_messages rule); seg: [n_rows][npass + 1].
cudaError_t launch_build_pass_csr(const uint64_t *row_ptr, const uint32_t *col, const float *w,
                                  uint64_t n_rows, const uint64_t *bounds_dev, int npass,
                                  int prenormalized, uint32_t *col_p, float *w_p, uint32_t *seg,
                                  cudaStream_t s, uint64_t *launches);

// ---- row ops ----------------------------------------------------------------
// out[row] = LayerNorm(deg(row) > 0 ? d[row] : m[row]), written in logical order.
struct LayerNormArgs {
  const float *d = nullptr;
  int64_t ldd = 0;
  const float *m = nullptr;
  int64_t ldm = 0;
  const uint64_t *row_ptr = nullptr;  // nullptr: every row takes d
  const float *gamma = nullptr, *beta = nullptr;  // phys order
  float eps = 1e-5f;
  float *out = nullptr;  // [rows][H] logical order
  uint64_t row_begin = 0, row_end = 0;
  uint32_t H = 0, P = 0;
  const uint32_t *scatter = nullptr;  // row r is written to out[scatter[r]] (compacted active rows)
};
cudaError_t launch_layernorm(const LayerNormArgs &a, cudaStream_t s, uint64_t *launches);
cudaError_t launch_pack_rows(const float *src, int64_t ld_src, uint32_t row_floats,
                             const uint32_t *idx, uint64_t n_idx, float *dst, cudaStream_t s,
                             uint64_t *launches);
cudaError_t launch_validate_csr(const uint64_t *row_ptr, const uint32_t *col_idx, uint64_t n_nodes,
                                uint64_t n_cols, uint64_t n_edges, int *d_flag, cudaStream_t s,
                                uint64_t *launches);

// ---- fused K-upd (gemm_upd_fused.cu, H = 128, tensor-core modes) ---------------------
// out[row] = LayerNorm(deg(row) > 0 ? GRU(Wf [S | agg] + bf, m) * keep : m) for rows
// [row_begin, row_begin + rows); every base pointer is indexed by the absolute row number.
struct FusedUpdArgs {
  int64_t rows = 0;
  uint64_t row_begin = 0;
  const float *A1 = nullptr;  // gather state [row][lda1]: S_0..S_{h-1} | agg (R = 1) or o | agg (R = 3)
  int64_t lda1 = 0;
  int32_t K1 = 0;  // (heads + 1) * 128 or 256
  const float *m = nullptr;  // message rows
  int64_t ldm = 0;
  const float *bf = nullptr, *bzr = nullptr, *bh = nullptr, *gamma = nullptr, *beta = nullptr;
  float eps = 1e-5f, keep = 1.f;
  const uint64_t *row_ptr = nullptr;  // nullptr: every row has neighbours
  const uint32_t *scatter = nullptr;  // row r is written to out[scatter[r]]
  float *out = nullptr;               // [.][128]
};

// ---- search.cu: batched differentiable_search / hierarchical_forward pieces ---------
cudaError_t launch_row_norms(const float *x, uint64_t rows, uint32_t dim, float *norms,
                             cudaStream_t s, uint64_t *launches);
cudaError_t launch_topk_softmax(float *dots, const float *qnorm, const float *enorm,
                                uint64_t n_queries, uint64_t n, int k, float temperature,
                                uint32_t *idx_out, float *w_out, cudaStream_t s,
                                uint64_t *launches);
cudaError_t launch_build_batch(const float *cur, const float *emb, const uint32_t *idx,
                               const float *w, uint64_t nq, int k, uint32_t dim, float *xb,
                               cudaStream_t s, uint64_t *launches);

}  // namespace ruvgnn
