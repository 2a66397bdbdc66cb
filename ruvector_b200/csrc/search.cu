// Device kernels for the batched differentiable_search / hierarchical_forward
// (crates/ruvector-gnn/src/search.rs:4-162), SURVEY.md row f2.
//
//   row_norms      cosine_similarity's norms: f64 accumulation, sqrt, cast to f32 (:10-19)
//   topk_softmax   per query: sim = dot / (|q| |e|) (0 when a norm is 0, :21-25), temperature
//                  softmax over ALL candidates with max subtraction and sum.max(1e-10)
//                  (:29-44), the k best by descending weight, ties in index order (stable
//                  sort, :79-87)
//   build_batch    aggregated = sum_k w_k e_k; combined = (cur + aggregated) / 2 (:133-147);
//                  the selected rows become the neighbour embeddings, the soft weights the
//                  edge weights (:150-158)
// The dot products themselves are one FP32 GEMM (queries x candidates^T, gemm_simt.cu).
// All HBM-bound streaming kernels.
#include "common.cuh"

namespace ruvgnn {
namespace {

constexpr int KMAX = 16;

__global__ void __launch_bounds__(256) row_norms_kernel(const float *x, uint64_t rows, uint32_t dim,
                                                        float *norms) {
  const int lane = threadIdx.x & 31;
  const uint64_t warp = ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const uint64_t nwarps = ((uint64_t)gridDim.x * blockDim.x) >> 5;
  for (uint64_t r = warp; r < rows; r += nwarps) {
    const float *p = x + r * dim;
    double s = 0.0;
    for (uint32_t t = lane; t < dim; t += 32) s += (double)p[t] * (double)p[t];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if (lane == 0) norms[r] = (float)sqrt(s);
  }
}

struct Best {
  float v[KMAX];
  uint32_t i[KMAX];
};

// (va, ia) ranks before (vb, ib)?  Larger similarity first, lower index on ties.
__device__ __forceinline__ bool before(float va, uint32_t ia, float vb, uint32_t ib) {
  return va > vb || (va == vb && ia < ib);
}

// One CTA per query.  dots: [n_queries][n] raw dot products, overwritten with the soft weights.
__global__ void __launch_bounds__(256) topk_softmax_kernel(float *dots, const float *qnorm,
                                                           const float *enorm, uint64_t n, int k,
                                                           float temperature, uint32_t *idx_out,
                                                           float *w_out) {
  __shared__ float s_v[256 * KMAX];
  __shared__ uint32_t s_i[256 * KMAX];
  __shared__ float s_red[256];
  const int tid = threadIdx.x;
  const uint64_t q = blockIdx.x;
  float *row = dots + q * n;
  const float nq = qnorm[q];

  // pass A: similarities and their maximum (softmax, search.rs:29-44)
  float mx = -INFINITY;
  for (uint64_t c = tid; c < n; c += 256) {
    const float ne = enorm[c];
    const float sim = (nq == 0.f || ne == 0.f) ? 0.f : row[c] / (nq * ne);
    row[c] = sim;
    mx = fmaxf(mx, sim);  // f32::max semantics: NaN is ignored
  }
  s_red[tid] = mx;
  __syncthreads();
  for (int s = 128; s > 0; s >>= 1) {
    if (tid < s) s_red[tid] = fmaxf(s_red[tid], s_red[tid + s]);
    __syncthreads();
  }
  mx = s_red[0];
  __syncthreads();
  // pass B: normaliser
  float sum = 0.f;
  for (uint64_t c = tid; c < n; c += 256) sum += expf((row[c] - mx) / temperature);
  s_red[tid] = sum;
  __syncthreads();
  for (int s = 128; s > 0; s >>= 1) {
    if (tid < s) s_red[tid] += s_red[tid + s];
    __syncthreads();
  }
  sum = fmaxf(s_red[0], 1e-10f);
  __syncthreads();
  // pass C: the k best by soft WEIGHT, ties in index order -- the reference sorts the weights
  // with a stable sort (:79-87), so similarities whose weights round equal keep index order.
  // A NaN weight (NaN / Inf in the embeddings) ranks lowest, so k slots are always filled.
  Best b;
  for (int j = 0; j < k; ++j) {
    b.v[j] = -INFINITY;
    b.i[j] = 0xffffffffu;
  }
  for (uint64_t c = tid; c < n; c += 256) {
    const float w = expf((row[c] - mx) / temperature) / sum;
    row[c] = w;
    const float key = (w == w) ? w : -INFINITY;
    if (before(key, (uint32_t)c, b.v[k - 1], b.i[k - 1])) {  // insertion into the sorted list
      int j = k - 1;
      while (j > 0 && before(key, (uint32_t)c, b.v[j - 1], b.i[j - 1])) {
        b.v[j] = b.v[j - 1];
        b.i[j] = b.i[j - 1];
        --j;
      }
      b.v[j] = key;
      b.i[j] = (uint32_t)c;
    }
  }
  for (int j = 0; j < k; ++j) {
    s_v[tid * KMAX + j] = b.v[j];
    s_i[tid * KMAX + j] = b.i[j];
  }
  __syncthreads();
  if (tid == 0) {
    // k-way merge of the 256 sorted per-thread lists
    int head[256];
    for (int t = 0; t < 256; ++t) head[t] = 0;
    for (int j = 0; j < k; ++j) {
      int bt = -1;
      float bv = -INFINITY;
      uint32_t bi = 0xffffffffu;
      for (int t = 0; t < 256; ++t) {
        if (head[t] >= k) continue;
        const float v = s_v[t * KMAX + head[t]];
        const uint32_t i = s_i[t * KMAX + head[t]];
        if (i == 0xffffffffu) continue;
        if (bt < 0 || before(v, i, bv, bi)) {
          bt = t;
          bv = v;
          bi = i;
        }
      }
      if (bt < 0) {  // cannot happen for k <= n; keep the outputs defined anyway
        idx_out[q * k + j] = 0;
        w_out[q * k + j] = 0.f;
        continue;
      }
      ++head[bt];
      idx_out[q * k + j] = bi;
      w_out[q * k + j] = row[bi];
    }
  }
}

// Rows [0, nq): combined queries; rows nq + q*k + j: the j-th selected candidate of query q.
__global__ void __launch_bounds__(256) build_batch_kernel(const float *cur, const float *emb,
                                                          const uint32_t *idx, const float *w,
                                                          uint64_t nq, int k, uint32_t dim,
                                                          float *xb) {
  const int lane = threadIdx.x & 31;
  const uint64_t warp = ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const uint64_t nwarps = ((uint64_t)gridDim.x * blockDim.x) >> 5;
  for (uint64_t q = warp; q < nq; q += nwarps) {
    for (uint32_t t = lane; t < dim; t += 32) {
      float agg = 0.f;
      for (int j = 0; j < k; ++j) {
        const float e = emb[(uint64_t)idx[q * k + j] * dim + t];
        agg = agg + w[q * k + j] * e;
        xb[(nq + q * k + j) * dim + t] = e;
      }
      xb[q * dim + t] = (cur[q * dim + t] + agg) / 2.0f;
    }
  }
}

unsigned blocks_for(uint64_t rows) {
  uint64_t b = (rows + 7) / 8;
  if (b > 148ull * 16) b = 148ull * 16;
  if (b == 0) b = 1;
  return (unsigned)b;
}

}  // namespace

cudaError_t launch_row_norms(const float *x, uint64_t rows, uint32_t dim, float *norms,
                             cudaStream_t s, uint64_t *launches) {
  if (rows == 0) return cudaSuccess;
  if (launches) ++*launches;
  row_norms_kernel<<<blocks_for(rows), 256, 0, s>>>(x, rows, dim, norms);
  return cudaGetLastError();
}

cudaError_t launch_topk_softmax(float *dots, const float *qnorm, const float *enorm,
                                uint64_t n_queries, uint64_t n, int k, float temperature,
                                uint32_t *idx_out, float *w_out, cudaStream_t s,
                                uint64_t *launches) {
  if (n_queries == 0 || n == 0) return cudaSuccess;
  if (k < 1 || k > KMAX) return cudaErrorInvalidValue;
  if (launches) ++*launches;
  topk_softmax_kernel<<<(unsigned)n_queries, 256, 0, s>>>(dots, qnorm, enorm, n, k, temperature,
                                                          idx_out, w_out);
  return cudaGetLastError();
}

cudaError_t launch_build_batch(const float *cur, const float *emb, const uint32_t *idx,
                               const float *w, uint64_t nq, int k, uint32_t dim, float *xb,
                               cudaStream_t s, uint64_t *launches) {
  if (nq == 0) return cudaSuccess;
  if (launches) ++*launches;
  build_batch_kernel<<<blocks_for(nq), 256, 0, s>>>(cur, emb, idx, w, nq, k, dim, xb);
  return cudaGetLastError();
}

}  // namespace ruvgnn
