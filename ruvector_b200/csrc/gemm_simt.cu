// FP32 FFMA GEMM with the layer's fused epilogues: the any-shape projection
// path (RUVGNN_MODE_FP32_SIMT) and the on-device reference the tcgen05 path is
// validated against.
//
//   C[M x N] = [A1 | A2] * W^T + epilogue,   W row-major [N][K1+K2]
//
// i.e. a batch of the reference's Linear::forward mat-vecs
// (crates/ruvector-gnn/src/layer.rs:37-41, y = W x + b) over all nodes; the
// K-concatenation expresses GRU sums like W_z a + U_z h (:257-258) as one GEMM.
// Tile 128x128x16, 256 threads, 8x8 micro-tile split in 4+4 halves so every
// shared-memory read is a conflict-free LDS.128.
#include "common.cuh"

namespace ruvgnn {
namespace {

constexpr int BM = 128, BN = 128, BK = 16, PAD = 4;

__device__ __forceinline__ float sigmoid_ref(float x) {
  // GRUCell::sigmoid (layer.rs:278-285): x > 0 ? 1/(1+exp(-x)) : exp(x)/(1+exp(x)).  Both
  // branches use exp(-|x|); written branch-free with one exp and one divide (the x <= 0 case
  // multiplies by the reciprocal instead of dividing: <= 1 ulp).
  const float t = expf(-fabsf(x));
  const float r = 1.0f / (1.0f + t);
  return x > 0.f ? r : t * r;
}

template <int EPI>
__device__ __forceinline__ void epilogue_store(const GemmArgs &g, int64_t row, int col, float v) {
  v += g.bias[col];
  if (EPI == EPI_BIAS) {
    g.out[row * g.ldo + col] = v;
  } else if (EPI == EPI_BIAS_ADD) {
    g.out[row * g.ldo + col] = v + g.add[row * g.ldadd + col];
  } else if (EPI == EPI_GATES) {
    const float s = sigmoid_ref(v);
    if (col < g.H)
      g.z[row * g.ldz + col] = s;
    else
      g.rh[row * g.ldrh + (col - g.H)] = s * g.m[row * g.ldm + (col - g.H)];
  } else {  // EPI_GRU  (layer.rs:266-274, 448-451)
    const float ht = tanhf(v);
    const float z = g.z[row * g.ldz + col], m = g.m[row * g.ldm + col];
    const float omz = 1.0f - z;
    g.out[row * g.ldo + col] = (omz * m + z * ht) * g.keep;
  }
}

template <int EPI>
__device__ __forceinline__ void epilogue_store4(const GemmArgs &g, int64_t row, int col,
                                                float4 v) {
  const float4 b = *reinterpret_cast<const float4 *>(g.bias + col);
  v.x += b.x; v.y += b.y; v.z += b.z; v.w += b.w;
  if (EPI == EPI_BIAS) {
    *reinterpret_cast<float4 *>(g.out + row * g.ldo + col) = v;
  } else if (EPI == EPI_BIAS_ADD) {
    const float4 a = *reinterpret_cast<const float4 *>(g.add + row * g.ldadd + col);
    v.x += a.x; v.y += a.y; v.z += a.z; v.w += a.w;
    *reinterpret_cast<float4 *>(g.out + row * g.ldo + col) = v;
  } else if (EPI == EPI_GATES) {
    float4 s = make_float4(sigmoid_ref(v.x), sigmoid_ref(v.y), sigmoid_ref(v.z), sigmoid_ref(v.w));
    if (col < g.H) {
      *reinterpret_cast<float4 *>(g.z + row * g.ldz + col) = s;
    } else {
      const int c = col - g.H;
      const float4 m = *reinterpret_cast<const float4 *>(g.m + row * g.ldm + c);
      s.x *= m.x; s.y *= m.y; s.z *= m.z; s.w *= m.w;
      *reinterpret_cast<float4 *>(g.rh + row * g.ldrh + c) = s;
    }
  } else {
    const float4 z = *reinterpret_cast<const float4 *>(g.z + row * g.ldz + col);
    const float4 m = *reinterpret_cast<const float4 *>(g.m + row * g.ldm + col);
    float4 o;
    o.x = ((1.0f - z.x) * m.x + z.x * tanhf(v.x)) * g.keep;
    o.y = ((1.0f - z.y) * m.y + z.y * tanhf(v.y)) * g.keep;
    o.z = ((1.0f - z.z) * m.z + z.z * tanhf(v.z)) * g.keep;
    o.w = ((1.0f - z.w) * m.w + z.w * tanhf(v.w)) * g.keep;
    *reinterpret_cast<float4 *>(g.out + row * g.ldo + col) = o;
  }
}

// ALIGNED: K1, K2 multiples of BK, N multiple of 4, every pointer/stride 16-byte aligned.
template <int EPI, bool ALIGNED>
__global__ void __launch_bounds__(256) gemm_nt_simt_kernel(GemmArgs g) {
  __shared__ __align__(16) float As[2][BK][BM + PAD];
  __shared__ __align__(16) float Bs[2][BK][BN + PAD];
  const int tid = threadIdx.x;
  const int tx = tid & 15, ty = tid >> 4;
  // 1-D grid, column tile fastest: the CTAs sharing an A row-tile are co-resident,
  // so A is read from HBM once and re-read from L2.
  const int ntn = (g.N + BN - 1) / BN;
  const int64_t row0 = (int64_t)(blockIdx.x / ntn) * BM;
  const int col0 = (int)(blockIdx.x % ntn) * BN;
  const int K = g.K1 + g.K2;
  const int ktiles = (K + BK - 1) / BK;

  // global->register staging: thread loads 2 float4 of A and 2 of B per k-tile
  const int lr = tid >> 2;        // 0..63 (+64 for the second)
  const int lk = (tid & 3) * 4;   // 0,4,8,12
  float4 ra[2], rb[2];

  auto load_tile = [&](int kt) {
    const int kbase = kt * BK;
#pragma unroll
    for (int r = 0; r < 2; ++r) {
      const int64_t arow = row0 + lr + r * 64;
      const int brow = col0 + lr + r * 64;
      const int k = kbase + lk;
      if (ALIGNED) {
        const float *ap = (k < g.K1) ? g.A1 + arow * g.lda1 + k : g.A2 + arow * g.lda2 + (k - g.K1);
        ra[r] = arow < g.M ? *reinterpret_cast<const float4 *>(ap) : make_float4(0, 0, 0, 0);
        rb[r] = brow < g.N ? *reinterpret_cast<const float4 *>(g.W + (int64_t)brow * K + k)
                           : make_float4(0, 0, 0, 0);
      } else {
        float av[4], bv[4];
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const int kk = k + e;
          av[e] = 0.f;
          bv[e] = 0.f;
          if (kk < K) {
            if (arow < g.M)
              av[e] = kk < g.K1 ? g.A1[arow * g.lda1 + kk] : g.A2[arow * g.lda2 + (kk - g.K1)];
            if (brow < g.N) bv[e] = g.W[(int64_t)brow * K + kk];
          }
        }
        ra[r] = make_float4(av[0], av[1], av[2], av[3]);
        rb[r] = make_float4(bv[0], bv[1], bv[2], bv[3]);
      }
    }
  };
  auto store_tile = [&](int buf) {
#pragma unroll
    for (int r = 0; r < 2; ++r) {
      const int m = lr + r * 64;
      As[buf][lk + 0][m] = ra[r].x;
      As[buf][lk + 1][m] = ra[r].y;
      As[buf][lk + 2][m] = ra[r].z;
      As[buf][lk + 3][m] = ra[r].w;
      Bs[buf][lk + 0][m] = rb[r].x;
      Bs[buf][lk + 1][m] = rb[r].y;
      Bs[buf][lk + 2][m] = rb[r].z;
      Bs[buf][lk + 3][m] = rb[r].w;
    }
  };

  float acc[8][8];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[i][j] = 0.f;

  load_tile(0);
  store_tile(0);
  __syncthreads();
  for (int kt = 0; kt < ktiles; ++kt) {
    const int buf = kt & 1;
    if (kt + 1 < ktiles) load_tile(kt + 1);
#pragma unroll
    for (int k = 0; k < BK; ++k) {
      const float4 a0 = *reinterpret_cast<const float4 *>(&As[buf][k][ty * 4]);
      const float4 a1 = *reinterpret_cast<const float4 *>(&As[buf][k][64 + ty * 4]);
      const float4 b0 = *reinterpret_cast<const float4 *>(&Bs[buf][k][tx * 4]);
      const float4 b1 = *reinterpret_cast<const float4 *>(&Bs[buf][k][64 + tx * 4]);
      const float a[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
      const float b[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
      for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
    }
    if (kt + 1 < ktiles) {
      store_tile(buf ^ 1);
      __syncthreads();
    }
  }

#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int64_t row = row0 + (i < 4 ? ty * 4 + i : 64 + ty * 4 + (i - 4));
    if (row >= g.M) continue;
#pragma unroll
    for (int jh = 0; jh < 2; ++jh) {
      const int col = col0 + jh * 64 + tx * 4;
      if (ALIGNED) {
        if (col < g.N)
          epilogue_store4<EPI>(g, row, col,
                               make_float4(acc[i][jh * 4 + 0], acc[i][jh * 4 + 1],
                                           acc[i][jh * 4 + 2], acc[i][jh * 4 + 3]));
      } else {
#pragma unroll
        for (int e = 0; e < 4; ++e)
          if (col + e < g.N) epilogue_store<EPI>(g, row, col + e, acc[i][jh * 4 + e]);
      }
    }
  }
}

inline bool al16(const void *p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }

template <int EPI>
cudaError_t launch_epi(const GemmArgs &g, cudaStream_t s) {
  bool aligned = (g.K1 % BK == 0) && (g.K2 % BK == 0) && (g.N % 4 == 0) && al16(g.A1) &&
                 al16(g.W) && al16(g.bias) && (g.lda1 % 4 == 0) && al16(g.out) &&
                 (g.ldo % 4 == 0);
  if (g.K2 > 0) aligned = aligned && al16(g.A2) && (g.lda2 % 4 == 0);
  if (EPI == EPI_BIAS_ADD) aligned = aligned && al16(g.add) && (g.ldadd % 4 == 0);
  if (EPI == EPI_GATES)
    aligned = aligned && (g.H % 4 == 0) && al16(g.z) && al16(g.rh) && al16(g.m) &&
              (g.ldz % 4 == 0) && (g.ldrh % 4 == 0) && (g.ldm % 4 == 0);
  if (EPI == EPI_GRU)
    aligned = aligned && al16(g.z) && al16(g.m) && (g.ldz % 4 == 0) && (g.ldm % 4 == 0);
  const int64_t nblocks = (int64_t)((g.N + BN - 1) / BN) * ((g.M + BM - 1) / BM);
  if (nblocks > 2147483647LL) return cudaErrorInvalidValue;
  dim3 grid((unsigned)nblocks);
  if (aligned)
    gemm_nt_simt_kernel<EPI, true><<<grid, 256, 0, s>>>(g);
  else
    gemm_nt_simt_kernel<EPI, false><<<grid, 256, 0, s>>>(g);
  return cudaGetLastError();
}

}  // namespace

cudaError_t launch_gemm_simt(const GemmArgs &a, EpiKind epi, cudaStream_t s, uint64_t *launches) {
  if (a.M <= 0 || a.N <= 0) return cudaSuccess;
  if (launches) ++*launches;
  switch (epi) {
    case EPI_BIAS: return launch_epi<EPI_BIAS>(a, s);
    case EPI_BIAS_ADD: return launch_epi<EPI_BIAS_ADD>(a, s);
    case EPI_GATES: return launch_epi<EPI_GATES>(a, s);
    case EPI_GRU: return launch_epi<EPI_GRU>(a, s);
  }
  return cudaErrorInvalidValue;
}

}  // namespace ruvgnn
