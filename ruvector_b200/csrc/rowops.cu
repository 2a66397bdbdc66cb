// Row-wise kernels: LayerNorm with the no-neighbour row select, halo row
// packing, CSR validation.  All HBM-bound, one warp per row.
#include "common.cuh"

namespace ruvgnn {
namespace {

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// LayerNorm::forward (crates/ruvector-gnn/src/layer.rs:68-81): mean, biased
// variance, (x-mean)/sqrt(var+eps)*gamma+beta.  Rows with no neighbours take
// LN(W_msg x + b) = LN(m) (:385-389); the others LN(dropout-scaled GRU output)
// (:411-417).  Input columns are in phys order, output is written in logical order.
__global__ void __launch_bounds__(256) layernorm_kernel(LayerNormArgs a) {
  const int lane = threadIdx.x & 31;
  const uint64_t warp = ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const uint64_t nwarps = ((uint64_t)gridDim.x * blockDim.x) >> 5;
  const int H = (int)a.H;
  const float invH = 1.0f / (float)H;
  for (uint64_t row = a.row_begin + warp; row < a.row_end; row += nwarps) {
    bool iso = false;
    if (a.row_ptr) iso = a.row_ptr[row] == a.row_ptr[row + 1];
    const float *src = iso ? a.m + row * (uint64_t)a.ldm : a.d + row * (uint64_t)a.ldd;
    float *dst = a.out + (a.scatter ? (uint64_t)a.scatter[row] : row) * (uint64_t)H;
    float s = 0.f;
    for (int t = lane; t < H; t += 32) s += src[t];
    const float mean = warp_sum(s) * invH;
    float v = 0.f;
    for (int t = lane; t < H; t += 32) {
      const float d = src[t] - mean;
      v = fmaf(d, d, v);
    }
    const float var = warp_sum(v) * invH;
    const float denom = sqrtf(var + a.eps);
    if (a.P == 0) {
      for (int t = lane; t < H; t += 32)
        dst[t] = a.gamma[t] * ((src[t] - mean) / denom) + a.beta[t];
    } else {
      // phys float4 index i*32+lane holds logical float4 lane*P+i
      const int P = (int)a.P;
      const float4 *s4 = reinterpret_cast<const float4 *>(src);
      const float4 *g4 = reinterpret_cast<const float4 *>(a.gamma);
      const float4 *b4 = reinterpret_cast<const float4 *>(a.beta);
      float4 *d4 = reinterpret_cast<float4 *>(dst);
      for (int i = 0; i < P; ++i) {
        const float4 x = s4[i * 32 + lane], g = g4[i * 32 + lane], b = b4[i * 32 + lane];
        float4 o;
        o.x = g.x * ((x.x - mean) / denom) + b.x;
        o.y = g.y * ((x.y - mean) / denom) + b.y;
        o.z = g.z * ((x.z - mean) / denom) + b.z;
        o.w = g.w * ((x.w - mean) / denom) + b.w;
        d4[lane * P + i] = o;
      }
    }
  }
}

// Fast layout (H = 128 P, phys column order): the row is read once as P float4 per lane and stays
// in registers for the mean, the variance and the normalised store.
template <int P>
__global__ void __launch_bounds__(256) layernorm_fast_kernel(LayerNormArgs a) {
  constexpr int H = 128 * P;
  const int lane = threadIdx.x & 31;
  const uint64_t warp = ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const uint64_t nwarps = ((uint64_t)gridDim.x * blockDim.x) >> 5;
  const float invH = 1.0f / (float)H;
  float4 g[P], b[P];
#pragma unroll
  for (int i = 0; i < P; ++i) {
    g[i] = reinterpret_cast<const float4 *>(a.gamma)[i * 32 + lane];
    b[i] = reinterpret_cast<const float4 *>(a.beta)[i * 32 + lane];
  }
  for (uint64_t row = a.row_begin + warp; row < a.row_end; row += nwarps) {
    bool iso = false;
    if (a.row_ptr) iso = a.row_ptr[row] == a.row_ptr[row + 1];
    const float4 *s4 = reinterpret_cast<const float4 *>(iso ? a.m + row * (uint64_t)a.ldm
                                                            : a.d + row * (uint64_t)a.ldd);
    float4 x[P];
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < P; ++i) {
      x[i] = s4[i * 32 + lane];
      s += (x[i].x + x[i].y) + (x[i].z + x[i].w);
    }
    const float mean = warp_sum(s) * invH;
    float v = 0.f;
#pragma unroll
    for (int i = 0; i < P; ++i) {
      x[i].x -= mean; x[i].y -= mean; x[i].z -= mean; x[i].w -= mean;
      v = fmaf(x[i].x, x[i].x, fmaf(x[i].y, x[i].y, fmaf(x[i].z, x[i].z, fmaf(x[i].w, x[i].w, v))));
    }
    const float denom = sqrtf(warp_sum(v) * invH + a.eps);
    float4 *d4 = reinterpret_cast<float4 *>(a.out + (a.scatter ? (uint64_t)a.scatter[row] : row) * (uint64_t)H);
#pragma unroll
    for (int i = 0; i < P; ++i) {  // phys float4 index i*32+lane holds logical float4 lane*P+i
      float4 o;
      o.x = g[i].x * (x[i].x / denom) + b[i].x;
      o.y = g[i].y * (x[i].y / denom) + b[i].y;
      o.z = g[i].z * (x[i].z / denom) + b[i].z;
      o.w = g[i].w * (x[i].w / denom) + b[i].w;
      d4[lane * P + i] = o;
    }
  }
}

__global__ void __launch_bounds__(256) pack_rows_kernel(const float *src, int64_t ld_src,
                                                        uint32_t row_floats, const uint32_t *idx,
                                                        uint64_t n_idx, float *dst) {
  const int lane = threadIdx.x & 31;
  const uint64_t warp = ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const uint64_t nwarps = ((uint64_t)gridDim.x * blockDim.x) >> 5;
  const bool al = (((uint64_t)ld_src | row_floats) & 3) == 0 &&
                  ((reinterpret_cast<uintptr_t>(src) | reinterpret_cast<uintptr_t>(dst)) & 15) == 0;
  const uint32_t v4 = al ? row_floats / 4 : 0;
  for (uint64_t r = warp; r < n_idx; r += nwarps) {
    const float4 *s = reinterpret_cast<const float4 *>(src + (uint64_t)idx[r] * ld_src);
    float4 *d = reinterpret_cast<float4 *>(dst + r * (uint64_t)row_floats);
    for (uint32_t t = lane; t < v4; t += 32) d[t] = __ldg(s + t);
    for (uint32_t t = v4 * 4 + lane; t < row_floats; t += 32)
      dst[r * (uint64_t)row_floats + t] = src[(uint64_t)idx[r] * ld_src + t];
  }
}

// flag bit 0: row_ptr not monotone / wrong ends; bit 1: col_idx out of range.
__global__ void validate_csr_kernel(const uint64_t *row_ptr, const uint32_t *col_idx,
                                    uint64_t n_nodes, uint64_t n_cols, uint64_t n_edges,
                                    int *flag) {
  const uint64_t tid = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
  int bad = 0;
  for (uint64_t i = tid; i < n_nodes; i += stride)
    if (row_ptr[i] > row_ptr[i + 1]) bad |= 1;
  if (tid == 0 && (row_ptr[0] != 0 || row_ptr[n_nodes] != n_edges)) bad |= 1;
  for (uint64_t e = tid; e < n_edges; e += stride)
    if ((uint64_t)col_idx[e] >= n_cols) bad |= 2;
  if (bad) atomicOr(flag, bad);
}

unsigned grid_for(uint64_t rows, int warps_per_block) {
  uint64_t b = (rows + warps_per_block - 1) / warps_per_block;
  const uint64_t cap = 148ull * 16;
  if (b > cap) b = cap;
  if (b == 0) b = 1;
  return (unsigned)b;
}

}  // namespace

cudaError_t launch_layernorm(const LayerNormArgs &a, cudaStream_t s, uint64_t *launches) {
  if (a.row_end <= a.row_begin) return cudaSuccess;
  if (launches) ++*launches;
  const unsigned grid = grid_for(a.row_end - a.row_begin, 8);
  const bool al = ((a.ldd | a.ldm) & 3) == 0 && a.H == 128u * a.P;
  switch (al ? (int)a.P : 0) {
    case 1: layernorm_fast_kernel<1><<<grid, 256, 0, s>>>(a); break;
    case 2: layernorm_fast_kernel<2><<<grid, 256, 0, s>>>(a); break;
    case 3: layernorm_fast_kernel<3><<<grid, 256, 0, s>>>(a); break;
    case 4: layernorm_fast_kernel<4><<<grid, 256, 0, s>>>(a); break;
    case 6: layernorm_fast_kernel<6><<<grid, 256, 0, s>>>(a); break;
    case 8: layernorm_fast_kernel<8><<<grid, 256, 0, s>>>(a); break;
    default: layernorm_kernel<<<grid, 256, 0, s>>>(a);
  }
  return cudaGetLastError();
}

cudaError_t launch_pack_rows(const float *src, int64_t ld_src, uint32_t row_floats,
                             const uint32_t *idx, uint64_t n_idx, float *dst, cudaStream_t s,
                             uint64_t *launches) {
  if (n_idx == 0) return cudaSuccess;
  if (launches) ++*launches;
  pack_rows_kernel<<<grid_for(n_idx, 8), 256, 0, s>>>(src, ld_src, row_floats, idx, n_idx, dst);
  return cudaGetLastError();
}

cudaError_t launch_validate_csr(const uint64_t *row_ptr, const uint32_t *col_idx, uint64_t n_nodes,
                                uint64_t n_cols, uint64_t n_edges, int *d_flag, cudaStream_t s,
                                uint64_t *launches) {
  if (launches) ++*launches;
  validate_csr_kernel<<<148 * 8, 256, 0, s>>>(row_ptr, col_idx, n_nodes, n_cols, n_edges, d_flag);
  return cudaGetLastError();
}

}  // namespace ruvgnn
