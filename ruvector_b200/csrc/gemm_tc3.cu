// A-in-TMEM variant of the tcgen05 3xTF32 GEMM (tcgen05.mma "TS" form).
//
// Why: with both operands in shared memory the three MMAs of one 8-wide k-step read
// 3 x (4 KB A + 8 KB B) from shared memory while TMA writes the next weight tile and the
// producer warps store the split A tile -- more than the 128 B/cycle an SM's shared memory
// moves.  Here the activation tile never touches shared memory:
//   warps 2-5   A producers: each thread owns ONE row of the 128-row tile, loads its 32-float
//               k-block (4 x 256-bit LDG), splits hi/lo in registers and writes both with
//               tcgen05.st.32x32b.x32 into a ring of TMEM column blocks
//   warp 1      MMA issuer: tcgen05.mma ... [d_tmem], [a_tmem], b_desc  (A from TMEM, B from smem)
//   warp 0      TMA producer for W_hi / W_lo (shared memory now holds only weight stages)
//   warps 6-9   epilogue (same fused, smem-transposed epilogue as the SS kernel)
// TMEM columns: [0, 2*BN) double-buffered accumulators, then a_stages x 64 columns
// (hi | lo) of A; BN <= 192 keeps 2 A stages inside the 512 columns.
#include <stdlib.h>

#include "tc_common.cuh"

namespace ruvgnn {
namespace tc {
namespace {

constexpr int BK = 32;

struct Tc3Params {
  TcParams p;
  int a_stages, b_stages;
};

template <int EPI>
__global__ void __launch_bounds__(NUM_THREADS, 1)
gemm_tc_ts_kernel(const __grid_constant__ CUtensorMap map_hi,
                  const __grid_constant__ CUtensorMap map_lo, const Tc3Params q) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  const TcParams &p = q.p;
  const GemmArgs &g = p.g;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int BN = p.block_n, SA = q.a_stages, SB = q.b_stages;
  const int b_tile_bytes = BN * BK * 4;
  const size_t stage_bytes = 2 * (size_t)b_tile_bytes;

  uint8_t *base = reinterpret_cast<uint8_t *>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) &
                                              ~(uintptr_t)1023);
  auto b_hi = [&](int s) { return base + s * stage_bytes; };
  auto b_lo = [&](int s) { return base + s * stage_bytes + b_tile_bytes; };
  uint64_t *bars = reinterpret_cast<uint64_t *>(base + (size_t)SB * stage_bytes);
  uint64_t *full_b = bars, *empty_b = bars + MAX_STAGES;
  uint64_t *full_a = bars + 2 * MAX_STAGES, *empty_a = bars + 3 * MAX_STAGES;
  uint64_t *tmem_full = bars + 4 * MAX_STAGES, *tmem_empty = bars + 4 * MAX_STAGES + 2;
  uint32_t *tmem_ptr = reinterpret_cast<uint32_t *>(bars + 4 * MAX_STAGES + 4);
  float *stile = reinterpret_cast<float *>(bars + 4 * MAX_STAGES + 6) +
                 (warp >= 6 ? warp - 6 : 0) * EPI_TILE_FLOATS;

  const uint32_t need_cols = 2 * BN + 64 * SA;
  const uint32_t tmem_cols = need_cols <= 32 ? 32 : need_cols <= 64 ? 64 : need_cols <= 128 ? 128
                             : need_cols <= 256 ? 256 : 512;
  if (warp == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&map_hi)) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&map_lo)) : "memory");
    for (int s = 0; s < SB; ++s) {
      mbar_init(&full_b[s], 1);
      mbar_init(&empty_b[s], 1);
    }
    for (int s = 0; s < SA; ++s) {
      mbar_init(&full_a[s], 4);  // one arrive per producer warp
      mbar_init(&empty_a[s], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&tmem_full[i], 1);
      mbar_init(&tmem_empty[i], 4);
    }
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc(tmem_ptr, tmem_cols);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;
  const uint32_t a_col0 = (uint32_t)(2 * BN);  // first A column

  const int K = g.K1 + g.K2;
  const int num_kb = K / BK;
  const int num_tiles = p.num_m_tiles * p.num_n_tiles;

  if (warp == 0) {
    // ===== TMA producer: weight k-blocks =====
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        const int n0 = (tile % p.num_n_tiles) * BN;
        for (int kb = 0; kb < num_kb; ++kb) {
          mbar_wait(&empty_b[stage], phase ^ 1);
          mbar_arrive_expect_tx(&full_b[stage], p.single_pass ? b_tile_bytes : 2 * b_tile_bytes);
          tma_load_2d(b_hi(stage), &map_hi, &full_b[stage], kb * BK, n0);
          if (!p.single_pass) tma_load_2d(b_lo(stage), &map_lo, &full_b[stage], kb * BK, n0);
          if (++stage == SB) {
            stage = 0;
            phase ^= 1;
          }
        }
      }
    }
  } else if (warp == 1) {
    // ===== MMA issuer: A from TMEM, B from shared memory =====
    if (lane == 0) {
      const uint32_t idesc = make_idesc_tf32(BN);
      int sa = 0, sb = 0, acc = 0;
      uint32_t pa = 0, pb = 0, acc_phase = 0;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        mbar_wait(&tmem_empty[acc], acc_phase ^ 1);
        tc_fence_after();
        const uint32_t d = tmem_base + (uint32_t)(acc * BN);
        for (int kb = 0; kb < num_kb; ++kb) {
          mbar_wait(&full_a[sa], pa);
          mbar_wait(&full_b[sb], pb);
          tc_fence_after();
          const uint32_t ah = tmem_base + a_col0 + (uint32_t)(sa * 64);
          const uint32_t al = ah + 32;
          const uint64_t bh = make_desc<BK>(smem_u32(b_hi(sb)));
          const uint64_t bl = make_desc<BK>(smem_u32(b_lo(sb)));
#pragma unroll
          for (int kk = 0; kk < BK / 8; ++kk) {
            const uint64_t off = (uint64_t)((kk * 32) >> 4);
            const uint32_t first = (kb | kk) ? 1u : 0u;
            if (p.single_pass) {
              umma_tf32_ts(d, ah + kk * 8, bh + off, idesc, first);
            } else {
              umma_tf32_ts(d, al + kk * 8, bh + off, idesc, first);
              umma_tf32_ts(d, ah + kk * 8, bl + off, idesc, 1u);
              umma_tf32_ts(d, ah + kk * 8, bh + off, idesc, 1u);
            }
          }
          umma_commit(&empty_a[sa]);
          umma_commit(&empty_b[sb]);
          if (kb == num_kb - 1) umma_commit(&tmem_full[acc]);
          if (++sa == SA) {
            sa = 0;
            pa ^= 1;
          }
          if (++sb == SB) {
            sb = 0;
            pb ^= 1;
          }
        }
        if (++acc == 2) {
          acc = 0;
          acc_phase ^= 1;
        }
      }
    }
  } else if (warp < 6) {
    // ===== A producers: one row per thread -> (hi, lo) -> TMEM =====
    const int quad = warp & 3;  // TMEM lane quadrant of this warp
    const int row_in_tile = quad * 32 + lane;
    const uint32_t lane_base = (uint32_t)(quad * 32) << 16;
    int stage = 0;
    uint32_t phase = 0;
    float4 cur[8], nxt[8];
    auto issue = [&](float4 (&v)[8], int tile, int kb) {
      const int64_t row = (int64_t)(tile / p.num_n_tiles) * BLOCK_M + row_in_tile;
      const int k0 = kb * BK;
      const float *src;
      if (k0 < g.K1)
        src = g.A1 + row * g.lda1 + k0;
      else
        src = g.A2 + row * g.lda2 + (k0 - g.K1);
      if (row < g.M) {
        // 256-bit loads: a thread reads its row's 128 contiguous bytes in 4 instructions, so the
        // warp's 32 rows cost 4 x 32 L1 wavefronts per k-block instead of 8 x 32.
#pragma unroll
        for (int i = 0; i < 4; ++i)
          asm volatile("ld.global.nc.v8.f32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                       : "=f"(v[2 * i].x), "=f"(v[2 * i].y), "=f"(v[2 * i].z), "=f"(v[2 * i].w),
                         "=f"(v[2 * i + 1].x), "=f"(v[2 * i + 1].y), "=f"(v[2 * i + 1].z),
                         "=f"(v[2 * i + 1].w)
                       : "l"(src + 8 * i));
      } else {
#pragma unroll
        for (int i = 0; i < 8; ++i) v[i] = make_float4(0.f, 0.f, 0.f, 0.f);
      }
    };
    // this thread's 128-byte slice of the k-block `dist` k-blocks ahead, pulled towards L2
    auto prefetch_ahead = [&](int tile, int kb) {
      int t = tile, k = kb + p.pf_dist;
      while (k >= num_kb) {
        k -= num_kb;
        t += gridDim.x;
      }
      if (t >= num_tiles) return;
      const int64_t row = (int64_t)(t / p.num_n_tiles) * BLOCK_M + row_in_tile;
      if (row >= g.M) return;
      const int k0 = k * BK;
      prefetch_l2(k0 < g.K1 ? g.A1 + row * g.lda1 + k0 : g.A2 + row * g.lda2 + (k0 - g.K1));
    };
    int tile = blockIdx.x, kb = 0;
    if (tile < num_tiles) issue(nxt, tile, 0);
    if (p.pf_dist > 0 && tile < num_tiles)
      for (int k = 1 - p.pf_dist; k < 0; ++k) prefetch_ahead(tile, k);  // k-blocks 1 .. dist-1 of the first tile
    while (tile < num_tiles) {
#pragma unroll
      for (int i = 0; i < 8; ++i) cur[i] = nxt[i];
      int ntile = tile, nkb = kb + 1;
      if (nkb == num_kb) {
        nkb = 0;
        ntile = tile + gridDim.x;
      }
      if (ntile < num_tiles) issue(nxt, ntile, nkb);
      if (p.pf_dist > 0) prefetch_ahead(tile, kb);
      mbar_wait(&empty_a[stage], phase ^ 1);
      tc_fence_after();
      const uint32_t ah = tmem_base + lane_base + a_col0 + (uint32_t)(stage * 64);
      uint32_t r[32];
      const float *cf = reinterpret_cast<const float *>(cur);
#pragma unroll
      for (int i = 0; i < 32; ++i) r[i] = (__float_as_uint(cf[i]) + 0x1000u) & 0xFFFFE000u;
      tmem_st32(ah, r);
      if (!p.single_pass) {
#pragma unroll
        for (int i = 0; i < 32; ++i) r[i] = __float_as_uint(cf[i] - __uint_as_float(r[i]));
        tmem_st32(ah + 32, r);
      }
      tmem_st_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&full_a[stage]);
      if (++stage == SA) {
        stage = 0;
        phase ^= 1;
      }
      tile = ntile;
      kb = nkb;
    }
  } else {
    // ===== epilogue =====
    const int quad = warp & 3;
    const int row_in_tile = quad * 32 + lane;
    int acc = 0;
    uint32_t acc_phase = 0;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
      const int64_t row = (int64_t)(tile / p.num_n_tiles) * BLOCK_M + row_in_tile;
      const int n0 = (tile % p.num_n_tiles) * BN;
      mbar_wait(&tmem_full[acc], acc_phase);
      tc_fence_after();
      const uint32_t taddr = tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)(acc * BN);
      const int64_t row0 = row - lane;
      for (int c = 0; c < BN; c += 32) {
        float v[32];
        tmem_ld32(taddr + (uint32_t)c, v);
        if (row0 >= g.M) continue;
        epilogue_chunk<EPI>(g, row0, n0 + c, v, stile, lane);
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tmem_empty[acc]);
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

template <int EPI>
cudaError_t launch_ts_epi(const CUtensorMap &hi, const CUtensorMap &lo, const Tc3Params &q, int grid,
                          size_t smem, cudaStream_t s) {
  // function attributes are per device: remember which devices have been configured
  static bool configured_dev[64] = {};
  int dev = 0;
  cudaGetDevice(&dev);
  bool &configured = configured_dev[dev & 63];
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(gemm_tc_ts_kernel<EPI>,
                                         cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    if (e != cudaSuccess) return e;
    configured = true;
  }
  gemm_tc_ts_kernel<EPI><<<grid, NUM_THREADS, smem, s>>>(hi, lo, q);
  return cudaGetLastError();
}

}  // namespace

int pick_block_n_ts(uint32_t n) {
  for (int bn = 192; bn >= 32; bn -= 32)
    if (n % bn == 0) return bn;
  return 0;
}

cudaError_t launch_gemm_tc_ts(const GemmArgs &g, EpiKind epi, const TcMaps &maps, int which,
                              bool single_pass, int sm_count, cudaStream_t s) {
  Tc3Params q;
  TcParams &p = q.p;
  p.g = g;
  p.block_n = maps.block_n3[which];
  p.single_pass = single_pass ? 1 : 0;
  p.debug = 0;
  p.pf_dist = tc_pf_dist();
  p.num_m_tiles = (int)((g.M + BLOCK_M - 1) / BLOCK_M);
  p.num_n_tiles = g.N / p.block_n;
  p.stages = 0;
  q.a_stages = (512 - 2 * p.block_n) / 64;
  if (q.a_stages > 4) q.a_stages = 4;
  if (q.a_stages < 2) return cudaErrorInvalidValue;
  const size_t stage_bytes = 2 * (size_t)p.block_n * BK * 4;
  const size_t fixed = 1024 + 384 + 4 * EPI_TILE_FLOATS * sizeof(float);
  int sb = (int)((227 * 1024 - fixed) / stage_bytes);
  if (sb > MAX_STAGES) sb = MAX_STAGES;
  if (sb < 2) return cudaErrorInvalidValue;
  q.b_stages = sb;
  const size_t smem = sb * stage_bytes + fixed;
  const int64_t tiles = (int64_t)p.num_m_tiles * p.num_n_tiles;
  const int grid = (int)(tiles < sm_count ? tiles : sm_count);
  switch (epi) {
    case EPI_BIAS: return launch_ts_epi<EPI_BIAS>(maps.hi3[which], maps.lo3[which], q, grid, smem, s);
    case EPI_BIAS_ADD:
      return launch_ts_epi<EPI_BIAS_ADD>(maps.hi3[which], maps.lo3[which], q, grid, smem, s);
    case EPI_GATES:
      return launch_ts_epi<EPI_GATES>(maps.hi3[which], maps.lo3[which], q, grid, smem, s);
    case EPI_GRU: return launch_ts_epi<EPI_GRU>(maps.hi3[which], maps.lo3[which], q, grid, smem, s);
  }
  return cudaErrorInvalidValue;
}

}  // namespace tc
}  // namespace ruvgnn
