// tcgen05 projection GEMMs for sm_100a (RUVGNN_MODE_FP32_TC / RUVGNN_MODE_TF32).
//
//   C[M x N] = [A1 | A2] * W^T + fused epilogue      (same contract as gemm_simt.cu)
//
// Two kernels share the warp roles below: this file's (A and B operands in shared memory) and
// gemm_tc3.cu's (A operand in TMEM); launch_gemm_tc picks per weight matrix.  The cta_group::2
// CTA-pair variants of round 1 never beat them (per-k-block cross-SM handshakes) and were removed.
//
// replaces the batched Linear::forward mat-vecs of crates/ruvector-gnn/src/layer.rs:37-41
// (w_msg :392, q/k/v :135-137, out :159, w_agg :408, the six GRU linears :255-275).
//
// FP32 parity on tensor cores: tcgen05 has no FP32 MMA, kind::tf32 keeps 11 significant
// bits of each input.  Every product is therefore error-compensated (3xTF32):
//     a = a_hi + a_lo,  w = w_hi + w_lo   (hi = TF32-rounded, lo = exact remainder)
//     a*w ~= a_lo*w_hi + a_hi*w_lo + a_hi*w_hi        (FP32 accumulate in TMEM)
// dropping only a_lo*w_lo (<= 2^-22 relative).  Weights are split once at layer creation
// and streamed by TMA; activation tiles are split by the producer warps on their way from
// global memory into 128B-swizzled shared memory.
//
// CTA = 10 warps, persistent over output tiles (BLOCK_M = 128 rows x BLOCK_N <= 256 cols):
//   warp 0      TMA producer: W_hi / W_lo k-blocks (32 floats = one 128B swizzle row)
//   warp 1      MMA issuer (one elected lane), owns the TMEM allocation
//   warps 2-5   A producers: LDG.128 -> hi/lo split -> STS into the swizzled A_hi / A_lo tiles
//   warps 6-9   epilogue: tcgen05.ld the accumulator, bias / gates / GRU, global stores
// Accumulators are double-buffered in TMEM (2 x BLOCK_N columns) so the epilogue of tile i
// overlaps the MMAs of tile i+1.
#include <cuda.h>
#include <stdlib.h>
#include <string.h>

#include <vector>

#include "gemm_tc.cuh"
#include "tc_common.cuh"

namespace ruvgnn {
namespace {

// BK = floats per k-block: 32 (128-byte rows, SWIZZLE_128B) or 16 (64-byte rows, SWIZZLE_64B).
using namespace tc;

struct SmemLayout {
  uint8_t *a_hi[MAX_STAGES], *a_lo[MAX_STAGES], *b_hi[MAX_STAGES], *b_lo[MAX_STAGES];
  uint64_t *full_a, *full_b, *empty, *tmem_full, *tmem_empty;
  uint32_t *tmem_ptr;
};

template <int EPI, int BK, bool COAL>
__global__ void __launch_bounds__(NUM_THREADS, 1)
gemm_tc_kernel(const __grid_constant__ CUtensorMap map_hi, const __grid_constant__ CUtensorMap map_lo,
               const TcParams p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  const GemmArgs &g = p.g;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  constexpr int BLOCK_K = BK;
  constexpr int A_TILE_BYTES = BLOCK_M * BK * 4;
  const int stages = p.stages, BN = p.block_n;
  const int b_tile_bytes = BN * BLOCK_K * 4;

  // carve shared memory (tiles 1024-byte aligned for the 128B swizzle)
  uint8_t *base = reinterpret_cast<uint8_t *>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) &
                                              ~(uintptr_t)1023);
  SmemLayout sm;
  for (int s = 0; s < stages; ++s) {
    uint8_t *st = base + (size_t)s * (2 * A_TILE_BYTES + 2 * b_tile_bytes);
    sm.a_hi[s] = st;
    sm.a_lo[s] = st + A_TILE_BYTES;
    sm.b_hi[s] = st + 2 * A_TILE_BYTES;
    sm.b_lo[s] = st + 2 * A_TILE_BYTES + b_tile_bytes;
  }
  uint64_t *bars =
      reinterpret_cast<uint64_t *>(base + (size_t)stages * (2 * A_TILE_BYTES + 2 * b_tile_bytes));
  sm.full_a = bars;
  sm.full_b = bars + MAX_STAGES;
  sm.empty = bars + 2 * MAX_STAGES;
  sm.tmem_full = bars + 3 * MAX_STAGES;
  sm.tmem_empty = bars + 3 * MAX_STAGES + 2;
  sm.tmem_ptr = reinterpret_cast<uint32_t *>(bars + 3 * MAX_STAGES + 4);
  float *stile = reinterpret_cast<float *>(bars + 3 * MAX_STAGES + 6) + (warp >= 6 ? warp - 6 : 0) * EPI_TILE_FLOATS;
  static_assert(BK == 16 || BK == 32, "k-block is one 64- or 128-byte swizzle row");

  const uint32_t tmem_cols = (2 * BN <= 32) ? 32 : (2 * BN <= 64) ? 64 : (2 * BN <= 128) ? 128
                             : (2 * BN <= 256) ? 256 : 512;
  if (warp == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&map_hi)) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&map_lo)) : "memory");
    for (int s = 0; s < stages; ++s) {
      mbar_init(&sm.full_a[s], 4);  // one arrive per A-producer warp
      mbar_init(&sm.full_b[s], 1);  // TMA producer's arrive.expect_tx
      mbar_init(&sm.empty[s], 1);   // tcgen05.commit
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&sm.tmem_full[i], 1);   // tcgen05.commit
      mbar_init(&sm.tmem_empty[i], 4);  // one arrive per epilogue warp
    }
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc(sm.tmem_ptr, tmem_cols);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *sm.tmem_ptr;

  const int K = g.K1 + g.K2;
  const int num_kb = K / BLOCK_K;
  const int num_tiles = p.num_m_tiles * p.num_n_tiles;

  if (warp == 0) {
    // ===== TMA producer: weight k-blocks =====
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        const int n0 = (tile % p.num_n_tiles) * BN;
        for (int kb = 0; kb < num_kb; ++kb) {
          mbar_wait(&sm.empty[stage], phase ^ 1);
          const uint32_t bytes = p.single_pass ? b_tile_bytes : 2 * b_tile_bytes;
          if (p.debug & 2) {
            mbar_arrive(&sm.full_b[stage]);
          } else {
            mbar_arrive_expect_tx(&sm.full_b[stage], bytes);
            tma_load_2d(sm.b_hi[stage], &map_hi, &sm.full_b[stage], kb * BLOCK_K, n0);
            if (!p.single_pass)
              tma_load_2d(sm.b_lo[stage], &map_lo, &sm.full_b[stage], kb * BLOCK_K, n0);
          }
          if (++stage == stages) {
            stage = 0;
            phase ^= 1;
          }
        }
      }
    }
  } else if (warp == 1) {
    // ===== MMA issuer =====
    if (lane == 0) {
      const uint32_t idesc = make_idesc_tf32(BN);
      int stage = 0;
      uint32_t phase = 0;
      int acc = 0;
      uint32_t acc_phase = 0;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        mbar_wait(&sm.tmem_empty[acc], acc_phase ^ 1);  // epilogue has drained this accumulator
        tc_fence_after();
        const uint32_t d = tmem_base + (uint32_t)(acc * BN);
        for (int kb = 0; kb < num_kb; ++kb) {
          mbar_wait(&sm.full_a[stage], phase);
          mbar_wait(&sm.full_b[stage], phase);
          tc_fence_after();
          const uint64_t ah = make_desc<BK>(smem_u32(sm.a_hi[stage]));
          const uint64_t al = make_desc<BK>(smem_u32(sm.a_lo[stage]));
          const uint64_t bh = make_desc<BK>(smem_u32(sm.b_hi[stage]));
          const uint64_t bl = make_desc<BK>(smem_u32(sm.b_lo[stage]));
#pragma unroll
          for (int kk = 0; kk < BLOCK_K / 8; ++kk) {
            const uint64_t off = (uint64_t)((kk * 32) >> 4);  // 8 tf32 = 32 bytes along K
            const uint32_t first = (kb | kk) ? 1u : 0u;
            if (p.debug & 4) {
            } else if (p.single_pass) {
              umma_tf32(d, ah + off, bh + off, idesc, first);
            } else {
              umma_tf32(d, al + off, bh + off, idesc, first);
              umma_tf32(d, ah + off, bl + off, idesc, 1u);
              umma_tf32(d, ah + off, bh + off, idesc, 1u);
            }
          }
          umma_commit(&sm.empty[stage]);  // frees the smem stage once these MMAs have read it
          if (kb == num_kb - 1) umma_commit(&sm.tmem_full[acc]);
          if (++stage == stages) {
            stage = 0;
            phase ^= 1;
          }
        }
        if (++acc == 2) {
          acc = 0;
          acc_phase ^= 1;
        }
      }
    }
  } else if (warp < 6) {
    // ===== A producers: global -> (hi, lo) -> swizzled smem =====
    // Register double buffer: the loads of k-block i+1 are in flight while k-block i is
    // split and stored, so the producers' global-load latency overlaps their own stores
    // and the wait for a free stage.
    constexpr int CH = BK / 4;          // 16-byte chunks per row
    constexpr int RPP = 128 / CH;       // rows covered per pass of the 128 producer threads
    const int t = threadIdx.x - 64;     // 0..127
    const int chunk = t % CH;
    const int rsub = t / CH;
    int stage = 0;
    uint32_t phase = 0;
    float4 cur[CH], nxt[CH];
    auto issue = [&](float4 (&v)[CH], int tile, int kb) {
      const int64_t m0 = (int64_t)(tile / p.num_n_tiles) * BLOCK_M;
      const int k0 = kb * BLOCK_K;
      const float *src;
      int64_t ld;
      if (k0 < g.K1) {
        src = g.A1 + k0;
        ld = g.lda1;
      } else {
        src = g.A2 + (k0 - g.K1);
        ld = g.lda2;
      }
#pragma unroll
      for (int i = 0; i < CH; ++i) {
        const int64_t row = m0 + i * RPP + rsub;
        v[i] = row < g.M ? __ldg(reinterpret_cast<const float4 *>(src + row * ld) + chunk)
                         : make_float4(0.f, 0.f, 0.f, 0.f);
      }
    };
    // row t of the k-block `dist` k-blocks ahead (one 128-byte line per producer thread), towards L2
    auto prefetch_ahead = [&](int tile, int kb) {
      int tt = tile, k = kb + p.pf_dist;
      while (k >= num_kb) {
        k -= num_kb;
        tt += gridDim.x;
      }
      if (tt >= num_tiles) return;
      const int64_t row = (int64_t)(tt / p.num_n_tiles) * BLOCK_M + t;
      if (row >= g.M) return;
      const int k0 = k * BLOCK_K;
      prefetch_l2(k0 < g.K1 ? g.A1 + row * g.lda1 + k0 : g.A2 + row * g.lda2 + (k0 - g.K1));
    };
    int tile = blockIdx.x, kb = 0;
    if (p.pf_dist > 0 && tile < num_tiles && !(p.debug & 1))
      for (int k = 1 - p.pf_dist; k < 0; ++k) prefetch_ahead(tile, k);
    if (tile < num_tiles && !(p.debug & 1)) issue(nxt, tile, 0);
    if (p.debug & 1)
      for (int i = 0; i < CH; ++i) nxt[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    while (tile < num_tiles) {
#pragma unroll
      for (int i = 0; i < CH; ++i) cur[i] = nxt[i];
      int ntile = tile, nkb = kb + 1;
      if (nkb == num_kb) {
        nkb = 0;
        ntile = tile + gridDim.x;
      }
      if (ntile < num_tiles && !(p.debug & 1)) issue(nxt, ntile, nkb);
      if (p.pf_dist > 0 && !(p.debug & 1)) prefetch_ahead(tile, kb);
      mbar_wait(&sm.empty[stage], phase ^ 1);
      uint8_t *hi = sm.a_hi[stage], *lo = sm.a_lo[stage];
#pragma unroll
      for (int i = 0; i < CH; ++i) {
        if (p.debug & 1) break;
        const uint32_t off = swz_off<BK>(i * RPP + rsub, chunk);
        float4 h;
        h.x = __uint_as_float((__float_as_uint(cur[i].x) + 0x1000u) & 0xFFFFE000u);
        h.y = __uint_as_float((__float_as_uint(cur[i].y) + 0x1000u) & 0xFFFFE000u);
        h.z = __uint_as_float((__float_as_uint(cur[i].z) + 0x1000u) & 0xFFFFE000u);
        h.w = __uint_as_float((__float_as_uint(cur[i].w) + 0x1000u) & 0xFFFFE000u);
        *reinterpret_cast<float4 *>(hi + off) = h;
        if (!p.single_pass) {
          float4 l;
          l.x = cur[i].x - h.x;
          l.y = cur[i].y - h.y;
          l.z = cur[i].z - h.z;
          l.w = cur[i].w - h.w;
          *reinterpret_cast<float4 *>(lo + off) = l;
        }
      }
      fence_proxy_async();  // generic-proxy stores -> visible to the tensor core (async proxy)
      __syncwarp();
      if (lane == 0) mbar_arrive(&sm.full_a[stage]);
      if (++stage == stages) {
        stage = 0;
        phase ^= 1;
      }
      tile = ntile;
      kb = nkb;
    }
  } else {
    // ===== epilogue: TMEM -> registers -> fused math -> global =====
    const int quad = warp & 3;  // TMEM lane quadrant this warp may access
    const int row_in_tile = quad * 32 + lane;
    int acc = 0;
    uint32_t acc_phase = 0;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
      const int64_t row = (int64_t)(tile / p.num_n_tiles) * BLOCK_M + row_in_tile;
      const int n0 = (tile % p.num_n_tiles) * BN;
      mbar_wait(&sm.tmem_full[acc], acc_phase);
      tc_fence_after();
      const uint32_t taddr = tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)(acc * BN);
      const int64_t row0 = row - lane;  // first row of this warp's 32-row slab
      for (int c = 0; c < BN; c += 32) {
        float v[32];
        tmem_ld32(taddr + (uint32_t)c, v);
        if (row0 >= g.M || (p.debug & 8)) continue;
        if (COAL) {
          epilogue_chunk<EPI>(g, row0, n0 + c, v, stile, lane);
        } else if (row < g.M) {
          epilogue_chunk_rowwise<EPI>(g, row, n0 + c, v);
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&sm.tmem_empty[acc]);
      if (++acc == 2) {
        acc = 0;
        acc_phase ^= 1;
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, tmem_cols);
  }
}

// ---- host side -----------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *,
                                  const cuuint64_t *, const cuuint64_t *, const cuuint32_t *,
                                  const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn get_encode() {
  static EncodeTiledFn fn = nullptr;
  if (fn) return fn;
  void *p = nullptr;
  cudaDriverEntryPointQueryResult qres;
  if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) !=
          cudaSuccess ||
      qres != cudaDriverEntryPointSuccess)
    return nullptr;
  fn = reinterpret_cast<EncodeTiledFn>(p);
  return fn;
}

int pick_block_n(uint32_t n) {
  static int cap = 0;  // RUVGNN_TC_BN_MAX: experiments with narrower column tiles
  if (!cap) {
    const char *e = getenv("RUVGNN_TC_BN_MAX");
    cap = e ? atoi(e) : 256;
    if (cap < 32 || cap > 256) cap = 256;
  }
  for (int bn = cap - cap % 32; bn >= 32; bn -= 32)  // epilogue walks 32-column chunks
    if (n % bn == 0) return bn;
  return 0;
}

// k-block size of the 1-CTA kernel: 32 floats (128-byte rows) by default; RUVGNN_TC_BK=16
// selects 64-byte rows (more, smaller stages) -- measured slower, kept for experiments.
int tc_block_k() {
  static int bk = 0;
  if (!bk) {
    const char *e = getenv("RUVGNN_TC_BK");
    bk = (e && atoi(e) == 16) ? 16 : 32;
  }
  return bk;
}

template <int EPI, int BK, bool COAL>
cudaError_t launch_epi_bk(const CUtensorMap &hi, const CUtensorMap &lo, const TcParams &p, int grid,
                          size_t smem, cudaStream_t s) {
  // function attributes are per device: remember which devices have been configured
  static bool configured_dev[64] = {};
  int dev = 0;
  cudaGetDevice(&dev);
  bool &configured = configured_dev[dev & 63];
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(gemm_tc_kernel<EPI, BK, COAL>,
                                         cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    if (e != cudaSuccess) return e;
    configured = true;
  }
  gemm_tc_kernel<EPI, BK, COAL><<<grid, NUM_THREADS, smem, s>>>(hi, lo, p);
  return cudaGetLastError();
}
template <int EPI>
cudaError_t launch_epi(const CUtensorMap &hi, const CUtensorMap &lo, const TcParams &p, int bk,
                       int grid, size_t smem, cudaStream_t s) {
  const bool coal = (tc_epi_mask() >> EPI) & 1;
  if (bk == 32)
    return coal ? launch_epi_bk<EPI, 32, true>(hi, lo, p, grid, smem, s)
                : launch_epi_bk<EPI, 32, false>(hi, lo, p, grid, smem, s);
  return coal ? launch_epi_bk<EPI, 16, true>(hi, lo, p, grid, smem, s)
              : launch_epi_bk<EPI, 16, false>(hi, lo, p, grid, smem, s);
}

}  // namespace

namespace tc {
int tc_pf_dist() {
  static int d = -1;
  if (d < 0) {
    const char *e = getenv("RUVGNN_TC_PF");
    d = e ? atoi(e) : 4;
    if (d < 0 || d > 64) d = 4;
  }
  return d;
}
int tc_epi_mask() {
  static int mask = -1;
  if (mask < 0) {
    const char *e = getenv("RUVGNN_TC_EPI_MASK");
    mask = e ? atoi(e) : 15;  // all kinds coalesced (measured: 59.7 ms/layer vs 61.8 row-wise)
  }
  return mask;
}
}  // namespace tc

cudaError_t tc_weights_create(const float *const host_mats[TC_W_COUNT],
                              const uint32_t rows[TC_W_COUNT], const uint32_t cols[TC_W_COUNT],
                              TcWeights *out) {
  out->usable = false;
  EncodeTiledFn encode = get_encode();
  if (!encode) return cudaSuccess;  // driver too old for TMA descriptors: path stays unusable
  size_t total = 0;
  for (int i = 0; i < TC_W_COUNT; ++i) {
    if (rows[i] == 0) continue;  // matrix not present for this layer shape
    if (cols[i] % 32 != 0 || pick_block_n(rows[i]) == 0) return cudaSuccess;
    total += 2 * (size_t)rows[i] * cols[i];
  }
  // the K-concatenated GEMMs switch A source at column H: H itself must be a k-block multiple
  if (rows[TC_W_O] % 32 != 0) return cudaSuccess;
  std::vector<float> blob(total);
  size_t off = 0;
  size_t offs_hi[TC_W_COUNT], offs_lo[TC_W_COUNT];
  for (int i = 0; i < TC_W_COUNT; ++i) {
    const size_t n = (size_t)rows[i] * cols[i];  // 0 for an absent matrix
    offs_hi[i] = off;
    offs_lo[i] = off + n;
    for (size_t j = 0; j < n; ++j) {
      const float w = host_mats[i][j];
      uint32_t b;
      memcpy(&b, &w, 4);
      b = (b + 0x1000u) & 0xFFFFE000u;  // round to TF32 (10 explicit mantissa bits)
      float h;
      memcpy(&h, &b, 4);
      blob[off + j] = h;
      blob[off + n + j] = w - h;  // exact in FP32
    }
    off += 2 * n;
  }
  cudaError_t e = cudaMalloc(&out->blob, total * sizeof(float));
  if (e != cudaSuccess) return e;
  e = cudaMemcpy(out->blob, blob.data(), total * sizeof(float), cudaMemcpyHostToDevice);
  if (e != cudaSuccess) return e;
  TcMaps *maps = new TcMaps();
  maps->block_k = tc_block_k();
  for (int i = 0; i < TC_W_COUNT; ++i) {
    out->hi[i] = out->blob + offs_hi[i];
    out->lo[i] = out->blob + offs_lo[i];
    out->rows[i] = rows[i];
    out->cols[i] = cols[i];
    maps->block_n[i] = 0;
    maps->block_n3[i] = 0;
    maps->ts_ok[i] = false;
    if (rows[i] == 0) continue;
    maps->block_n[i] = pick_block_n(rows[i]);
    const cuuint64_t dims[2] = {cols[i], rows[i]};
    const cuuint64_t strides[1] = {(cuuint64_t)cols[i] * sizeof(float)};
    const cuuint32_t box[2] = {(cuuint32_t)maps->block_k, (cuuint32_t)maps->block_n[i]};
    const cuuint32_t estr[2] = {1, 1};
    for (int part = 0; part < 2; ++part) {
      CUresult r = encode(part ? &maps->lo[i] : &maps->hi[i], CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2,
                          part ? out->lo[i] : out->hi[i], dims, strides, box, estr,
                          CU_TENSOR_MAP_INTERLEAVE_NONE,
                          maps->block_k == 32 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B,
                          CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
      if (r != CUDA_SUCCESS) {
        delete maps;
        cudaFree(out->blob);
        out->blob = nullptr;
        return cudaErrorInvalidValue;
      }
    }
    // A-in-TMEM kernel: column tiles of at most 192 so two A stages fit beside the accumulators
    maps->ts_ok[i] = false;
    maps->block_n3[i] = pick_block_n_ts(rows[i]);
    if (maps->block_n3[i] > 0) {
      const cuuint32_t box3[2] = {32u, (cuuint32_t)maps->block_n3[i]};
      CUresult r1 = encode(&maps->hi3[i], CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, out->hi[i], dims,
                           strides, box3, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                           CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                           CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
      CUresult r2 = encode(&maps->lo3[i], CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, out->lo[i], dims,
                           strides, box3, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                           CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                           CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
      maps->ts_ok[i] = r1 == CUDA_SUCCESS && r2 == CUDA_SUCCESS;
    }
  }
  out->tmaps = maps;
  out->usable = true;
  return cudaSuccess;
}

void tc_weights_destroy(TcWeights *w) {
  if (w->blob) cudaFree(w->blob);
  w->blob = nullptr;
  delete static_cast<TcMaps *>(w->tmaps);
  w->tmaps = nullptr;
  w->usable = false;
}

cudaError_t launch_gemm_tc(const GemmArgs &g, EpiKind epi, const TcWeights &w, int which,
                           bool single_pass, int sm_count, cudaStream_t s, uint64_t *launches) {
  if (g.M <= 0 || g.N <= 0) return cudaSuccess;
  const TcMaps *maps = static_cast<const TcMaps *>(w.tmaps);
  if (!w.usable || !maps) return cudaErrorNotSupported;
  if ((uint32_t)g.N != w.rows[which] || (uint32_t)(g.K1 + g.K2) != w.cols[which] ||
      g.K1 % 32 != 0 || g.K2 % 32 != 0)
    return cudaErrorInvalidValue;
  {
    // A operand from TMEM (gemm_tc3.cu) vs from shared memory (this file).  Measured on B200:
    // the TMEM kernel wins or ties whenever both use the same column-tile width (all GEMMs at
    // H = 384: 18.4 vs 19.6 ms per layer), the shared-memory kernel wins where it can use
    // 256-wide tiles (H = 768: TMEM holds at most 2 x 192 accumulator columns + the A ring).
    // RUVGNN_TC_TS = 0 never / 1 always / unset: that rule.
    static int ts = -2;
    if (ts == -2) {
      const char *e = getenv("RUVGNN_TC_TS");
      ts = e ? atoi(e) : -1;
    }
    const bool use_ts = ts == 1 || (ts == -1 && maps->block_n3[which] == maps->block_n[which]);
    if (use_ts && maps->ts_ok[which]) {
      if (launches) ++*launches;
      return launch_gemm_tc_ts(g, epi, *maps, which, single_pass, sm_count, s);
    }
  }
  const int BLOCK_K = maps->block_k;
  const int A_TILE_BYTES = BLOCK_M * BLOCK_K * 4;
  TcParams p;
  p.g = g;
  p.block_n = maps->block_n[which];
  p.single_pass = single_pass ? 1 : 0;
  {
    static int dbg = -1;
    if (dbg < 0) {
      const char *e = getenv("RUVGNN_TC_DEBUG");
      dbg = e ? atoi(e) : 0;
    }
    p.debug = dbg;
  }
  p.pf_dist = tc_pf_dist();
  p.num_m_tiles = (int)((g.M + BLOCK_M - 1) / BLOCK_M);
  p.num_n_tiles = g.N / p.block_n;
  const size_t stage_bytes = 2 * (size_t)A_TILE_BYTES + 2 * (size_t)p.block_n * BLOCK_K * 4;
  const size_t fixed = 1024 /*align*/ + 256 /*barriers*/ + 4 * EPI_TILE_FLOATS * sizeof(float);
  int stages = (int)((227 * 1024 - fixed) / stage_bytes);
  if (stages > MAX_STAGES) stages = MAX_STAGES;
  if (stages < 2) return cudaErrorInvalidValue;
  p.stages = stages;
  const size_t smem = stages * stage_bytes + fixed;
  const int64_t tiles = (int64_t)p.num_m_tiles * p.num_n_tiles;
  const int grid = (int)(tiles < sm_count ? tiles : sm_count);
  if (launches) ++*launches;
  switch (epi) {
    case EPI_BIAS:
      return launch_epi<EPI_BIAS>(maps->hi[which], maps->lo[which], p, BLOCK_K, grid, smem, s);
    case EPI_BIAS_ADD:
      return launch_epi<EPI_BIAS_ADD>(maps->hi[which], maps->lo[which], p, BLOCK_K, grid, smem, s);
    case EPI_GATES:
      return launch_epi<EPI_GATES>(maps->hi[which], maps->lo[which], p, BLOCK_K, grid, smem, s);
    case EPI_GRU:
      return launch_epi<EPI_GRU>(maps->hi[which], maps->lo[which], p, BLOCK_K, grid, smem, s);
  }
  return cudaErrorInvalidValue;
}

}  // namespace ruvgnn
