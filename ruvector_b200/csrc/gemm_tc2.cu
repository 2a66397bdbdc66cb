// CTA-pair (cta_group::2) variant of the tcgen05 3xTF32 GEMM.
//
// The 1-CTA kernel is bound by operand traffic from L2 into each SM (a 128-row tile pulls
// the whole W_hi/W_lo k-block, 64 KB per 1536 MMA cycles).  Here two CTAs of a cluster own
// 256 consecutive rows of the same column tile: each CTA loads and splits its own 128 A
// rows but only HALF of the weight k-block (BLOCK_N/2 rows); one tcgen05.mma.cta_group::2
// (M = 256) issued by the leader CTA reads both halves.  Weight bytes per SM halve and the
// smaller stage (64 KB) leaves room for 3 stages.
//
// Synchronisation (all mbarriers live at identical offsets in both CTAs):
//   full_a[s]   leader's, 8 arrivals: the 4 A-producer warps of each CTA (remote arrive)
//   full_b[s]   leader's, 1 arrival (leader's expect_tx) + TMA bytes from BOTH CTAs
//   empty[s]    each CTA's own, signalled by tcgen05.commit ... multicast::cluster
//   tmem_full   each CTA's own, multicast commit;  tmem_empty: leader's, 8 epilogue warps
#include <stdlib.h>

#include "tc_common.cuh"

namespace ruvgnn {
namespace tc {
namespace {

constexpr uint32_t kPeerMask = 0xFEFFFFFFu;  // clears the CTA-rank bit: address in the even CTA

__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// arrive on the LEADER CTA's copy of `bar` (works from either CTA of the pair)
__device__ __forceinline__ void mbar_arrive_leader(uint64_t *bar) {
  asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(smem_u32(bar) &
                                                                                  kPeerMask)
               : "memory");
}
__device__ __forceinline__ void tma_load_2d_pair(void *dst, const CUtensorMap *map, uint64_t *bar,
                                                 int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes "
      "[%0], [%1, {%3, %4}], [%2];" ::"r"(smem_u32(dst)),
      "l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(bar) & kPeerMask), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void tmem_alloc_pair(uint32_t *dst_smem, uint32_t cols) {
  asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(
                   smem_u32(dst_smem)),
               "r"(cols)
               : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc_pair(uint32_t addr, uint32_t cols) {
  asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(addr), "r"(cols)
               : "memory");
}
__device__ __forceinline__ void umma_tf32_pair(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc,
                                               uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::tf32 [%0], %1, %2, %3, p;\n\t"
      "}" ::"r"(tmem_d),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// tcgen05.commit to the barrier at this offset in both CTAs of the pair
__device__ __forceinline__ void umma_commit_pair(uint64_t *bar) {
  asm volatile(
      "tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 "
      "[%0], %1;" ::"r"(smem_u32(bar)),
      "h"((uint16_t)3)
      : "memory");
}
// cute::UMMA::InstrDescriptor for the pair: M = 256 (m_dim = 16)
__device__ __forceinline__ uint32_t make_idesc_tf32_pair(int n) {
  return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(n >> 3) << 17) | (16u << 24);
}

constexpr int BK = 32;
constexpr int A_TILE_BYTES = BLOCK_M * BK * 4;

template <int EPI, bool COAL>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(NUM_THREADS, 1)
gemm_tc_pair_kernel(const __grid_constant__ CUtensorMap map_hi,
                    const __grid_constant__ CUtensorMap map_lo, const TcParams p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  const GemmArgs &g = p.g;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const bool leader = rank == 0;
  const int stages = p.stages, BN = p.block_n, HN = BN / 2;
  const int b_tile_bytes = HN * BK * 4;  // this CTA's half of one weight k-block copy
  const size_t stage_bytes = 2 * (size_t)A_TILE_BYTES + 2 * (size_t)b_tile_bytes;

  uint8_t *base = reinterpret_cast<uint8_t *>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) &
                                              ~(uintptr_t)1023);
  auto a_hi = [&](int s) { return base + s * stage_bytes; };
  auto a_lo = [&](int s) { return base + s * stage_bytes + A_TILE_BYTES; };
  auto b_hi = [&](int s) { return base + s * stage_bytes + 2 * A_TILE_BYTES; };
  auto b_lo = [&](int s) { return base + s * stage_bytes + 2 * A_TILE_BYTES + b_tile_bytes; };
  uint64_t *bars = reinterpret_cast<uint64_t *>(base + (size_t)stages * stage_bytes);
  uint64_t *full_a = bars, *full_b = bars + MAX_STAGES, *empty = bars + 2 * MAX_STAGES;
  uint64_t *tmem_full = bars + 3 * MAX_STAGES, *tmem_empty = bars + 3 * MAX_STAGES + 2;
  uint32_t *tmem_ptr = reinterpret_cast<uint32_t *>(bars + 3 * MAX_STAGES + 4);
  float *stile = reinterpret_cast<float *>(bars + 3 * MAX_STAGES + 6) + (warp >= 6 ? warp - 6 : 0) * EPI_TILE_FLOATS;

  const uint32_t tmem_cols = (2 * BN <= 32) ? 32 : (2 * BN <= 64) ? 64 : (2 * BN <= 128) ? 128
                             : (2 * BN <= 256) ? 256 : 512;
  if (warp == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&map_hi)) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&map_lo)) : "memory");
    for (int s = 0; s < stages; ++s) {
      mbar_init(&full_a[s], 8);
      mbar_init(&full_b[s], 1);
      mbar_init(&empty[s], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&tmem_full[i], 1);
      mbar_init(&tmem_empty[i], 8);
    }
    fence_barrier_init();
  }
  cluster_sync_all();  // both CTAs' barriers exist before any remote arrive / multicast commit
  if (warp == 1) tmem_alloc_pair(tmem_ptr, tmem_cols);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;

  const int K = g.K1 + g.K2;
  const int num_kb = K / BK;
  const int num_tiles = p.num_m_tiles * p.num_n_tiles;  // m tiles are 256 rows here
  const int pair = blockIdx.x >> 1, num_pairs = gridDim.x >> 1;

  if (warp == 0) {
    // ===== TMA producer: this CTA's half of the weight k-block =====
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int tile = pair; tile < num_tiles; tile += num_pairs) {
        const int n0 = (tile % p.num_n_tiles) * BN + (int)rank * HN;
        for (int kb = 0; kb < num_kb; ++kb) {
          mbar_wait(&empty[stage], phase ^ 1);
          if (p.debug & 2) {
            if (leader) mbar_arrive(&full_b[stage]);
          } else {
            if (leader) {
              const uint32_t bytes = (p.single_pass ? 1u : 2u) * 2u * (uint32_t)b_tile_bytes;
              mbar_arrive_expect_tx(&full_b[stage], bytes);
            }
            tma_load_2d_pair(b_hi(stage), &map_hi, &full_b[stage], kb * BK, n0);
            if (!p.single_pass) tma_load_2d_pair(b_lo(stage), &map_lo, &full_b[stage], kb * BK, n0);
          }
          if (++stage == stages) {
            stage = 0;
            phase ^= 1;
          }
        }
      }
    }
  } else if (warp == 1) {
    // ===== MMA issuer: leader CTA only =====
    if (leader && lane == 0) {
      const uint32_t idesc = make_idesc_tf32_pair(BN);
      int stage = 0;
      uint32_t phase = 0;
      int acc = 0;
      uint32_t acc_phase = 0;
      for (int tile = pair; tile < num_tiles; tile += num_pairs) {
        mbar_wait(&tmem_empty[acc], acc_phase ^ 1);
        tc_fence_after();
        const uint32_t d = tmem_base + (uint32_t)(acc * BN);
        for (int kb = 0; kb < num_kb; ++kb) {
          mbar_wait(&full_a[stage], phase);
          mbar_wait(&full_b[stage], phase);
          tc_fence_after();
          const uint64_t ah = make_desc<BK>(smem_u32(a_hi(stage)));
          const uint64_t al = make_desc<BK>(smem_u32(a_lo(stage)));
          const uint64_t bh = make_desc<BK>(smem_u32(b_hi(stage)));
          const uint64_t bl = make_desc<BK>(smem_u32(b_lo(stage)));
#pragma unroll
          for (int kk = 0; kk < BK / 8; ++kk) {
            const uint64_t off = (uint64_t)((kk * 32) >> 4);
            const uint32_t first = (kb | kk) ? 1u : 0u;
            if (p.debug & 4) {
            } else if (p.single_pass) {
              umma_tf32_pair(d, ah + off, bh + off, idesc, first);
            } else {
              umma_tf32_pair(d, al + off, bh + off, idesc, first);
              umma_tf32_pair(d, ah + off, bl + off, idesc, 1u);
              umma_tf32_pair(d, ah + off, bh + off, idesc, 1u);
            }
          }
          umma_commit_pair(&empty[stage]);
          if (kb == num_kb - 1) umma_commit_pair(&tmem_full[acc]);
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
    // ===== A producers: this CTA's 128 rows -> (hi, lo) -> swizzled smem =====
    constexpr int CH = BK / 4, RPP = 128 / CH;
    const int t = threadIdx.x - 64;
    const int chunk = t % CH, rsub = t / CH;
    int stage = 0;
    uint32_t phase = 0;
    float4 cur[CH], nxt[CH];
    auto issue = [&](float4 (&v)[CH], int tile, int kb) {
      const int64_t m0 = (int64_t)(tile / p.num_n_tiles) * (2 * BLOCK_M) + (int64_t)rank * BLOCK_M;
      const int k0 = kb * BK;
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
    int tile = pair, kb = 0;
    if (tile < num_tiles && !(p.debug & 1)) issue(nxt, tile, 0);
    if (p.debug & 1)
      for (int i = 0; i < CH; ++i) nxt[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    while (tile < num_tiles) {
#pragma unroll
      for (int i = 0; i < CH; ++i) cur[i] = nxt[i];
      int ntile = tile, nkb = kb + 1;
      if (nkb == num_kb) {
        nkb = 0;
        ntile = tile + num_pairs;
      }
      if (ntile < num_tiles && !(p.debug & 1)) issue(nxt, ntile, nkb);
      mbar_wait(&empty[stage], phase ^ 1);
      uint8_t *hi = a_hi(stage), *lo = a_lo(stage);
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
      fence_proxy_async();  // generic stores -> async proxy (tensor core)
      __syncwarp();
      if (lane == 0) mbar_arrive_leader(&full_a[stage]);
      if (++stage == stages) {
        stage = 0;
        phase ^= 1;
      }
      tile = ntile;
      kb = nkb;
    }
  } else {
    // ===== epilogue: this CTA's 128 accumulator rows =====
    const int quad = warp & 3;
    const int row_in_tile = quad * 32 + lane;
    int acc = 0;
    uint32_t acc_phase = 0;
    for (int tile = pair; tile < num_tiles; tile += num_pairs) {
      const int64_t row = (int64_t)(tile / p.num_n_tiles) * (2 * BLOCK_M) + (int64_t)rank * BLOCK_M +
                          row_in_tile;
      const int n0 = (tile % p.num_n_tiles) * BN;
      mbar_wait(&tmem_full[acc], acc_phase);
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
      if (lane == 0) mbar_arrive_leader(&tmem_empty[acc]);
      if (++acc == 2) {
        acc = 0;
        acc_phase ^= 1;
      }
    }
  }

  // nobody may leave while the peer can still read this CTA's smem / signal its barriers
  tc_fence_before();
  cluster_sync_all();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc_pair(tmem_base, tmem_cols);
  }
}

template <int EPI>
cudaError_t launch_pair_epi(const CUtensorMap &hi, const CUtensorMap &lo, const TcParams &p, int grid,
                            size_t smem, cudaStream_t s) {
  // function attributes are per device: remember which devices have been configured
  static bool configured_dev[64] = {};
  int dev = 0;
  cudaGetDevice(&dev);
  bool &configured = configured_dev[dev & 63];
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(gemm_tc_pair_kernel<EPI, true>,
                                         cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    if (e == cudaSuccess)
      e = cudaFuncSetAttribute(gemm_tc_pair_kernel<EPI, false>,
                               cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    if (e != cudaSuccess) return e;
    configured = true;
  }
  if ((tc_epi_mask() >> EPI) & 1)
    gemm_tc_pair_kernel<EPI, true><<<grid, NUM_THREADS, smem, s>>>(hi, lo, p);
  else
    gemm_tc_pair_kernel<EPI, false><<<grid, NUM_THREADS, smem, s>>>(hi, lo, p);
  return cudaGetLastError();
}

}  // namespace

cudaError_t launch_gemm_tc_pair(const GemmArgs &g, EpiKind epi, const TcMaps &maps, int which,
                                bool single_pass, int sm_count, cudaStream_t s) {
  TcParams p;
  p.g = g;
  p.block_n = maps.block_n[which];
  p.single_pass = single_pass ? 1 : 0;
  {
    static int dbg = -1;
    if (dbg < 0) {
      const char *e = getenv("RUVGNN_TC_DEBUG");
      dbg = e ? atoi(e) : 0;
    }
    p.debug = dbg;
  }
  p.num_m_tiles = (int)((g.M + 2 * BLOCK_M - 1) / (2 * BLOCK_M));
  p.num_n_tiles = g.N / p.block_n;
  const size_t stage_bytes = 2 * (size_t)A_TILE_BYTES + 2 * (size_t)(p.block_n / 2) * BK * 4;
  const size_t fixed = 1024 /*align*/ + 256 /*barriers*/ + 4 * EPI_TILE_FLOATS * sizeof(float);
  int stages = (int)((227 * 1024 - fixed) / stage_bytes);
  if (stages > MAX_STAGES) stages = MAX_STAGES;
  if (stages < 2) return cudaErrorInvalidValue;
  p.stages = stages;
  const size_t smem = stages * stage_bytes + fixed;
  const int64_t tiles = (int64_t)p.num_m_tiles * p.num_n_tiles;
  int pairs = sm_count / 2;
  if (tiles < pairs) pairs = (int)tiles;
  const int grid = 2 * pairs;
  switch (epi) {
    case EPI_BIAS: return launch_pair_epi<EPI_BIAS>(maps.hi2[which], maps.lo2[which], p, grid, smem, s);
    case EPI_BIAS_ADD:
      return launch_pair_epi<EPI_BIAS_ADD>(maps.hi2[which], maps.lo2[which], p, grid, smem, s);
    case EPI_GATES:
      return launch_pair_epi<EPI_GATES>(maps.hi2[which], maps.lo2[which], p, grid, smem, s);
    case EPI_GRU: return launch_pair_epi<EPI_GRU>(maps.hi2[which], maps.lo2[which], p, grid, smem, s);
  }
  return cudaErrorInvalidValue;
}

}  // namespace tc
}  // namespace ruvgnn
