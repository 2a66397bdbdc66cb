// Fused K-upd for H = 128 (the config 1 / 3 / 5 layer shape): one tcgen05 kernel per row block
// instead of four GEMMs and a LayerNorm launch.
//
//   a   = Wf [S | agg] + bf              Wf = [W_agg W_ov | W_agg] folded at layer creation
//                                        (layer.rs:156-159 out-proj, :407 combine, :408 w_agg)
//   z   = sigmoid(W_z a + U_z m + b)     (GRUCell::forward, layer.rs:255-275)
//   rh  = sigmoid(W_r a + U_r m + b) * m
//   d   = ((1 - z) m + z tanh(W_h a + U_h rh + b)) * (1 - p)      (:448-451 dropout scale)
//   out = LayerNorm(deg > 0 ? d : m)     (:68-81, no-neighbour rows :385-389)
//
// A 128-row tile goes through all five stages inside one CTA; nothing between the gather state
// and the output row touches HBM: per node the kernel reads (heads + 1) H floats of gather state
// and the message row, and writes the output row once.
//
// How the chain stays on chip: tcgen05.mma accepts its A operand from TMEM (TS form) in the very
// layout the accumulator has -- TMEM lane = tile row, column = k.  An epilogue thread owns one row:
// it reads the accumulator (tcgen05.ld), applies bias / gate, splits the value into its TF32 part
// and the exact remainder (3xTF32, gemm_tc.cu) and writes both back with tcgen05.st as the next
// stage's A operand.  TMEM columns (512):
//   X [  0,128)  accumulator of a, then z, then r
//   Y [128,256)  ring of two streamed message-row k-blocks (hi | lo, 32 + 32 columns each) during the gate
//                GEMMs; accumulator of the candidate GEMM once that ring has drained
//   P [256,384)  resident A operand, TF32 part:  a, later rh  } during stage 1, before a exists: ring of
//   Q [384,512)  resident A operand, remainder               } four streamed gather-state k-blocks
// z (later d) of the tile lives in shared memory, one private row per epilogue thread.
// Stage 1 of tile n+1 touches X and P | Q only, so it runs under the last epilogue (candidate
// activation, LayerNorm, output store) of tile n, which reads Y.
//
// Warp roles (as gemm_tc3.cu): warp 0 TMA weight producer, warp 1 MMA issuer, warps 2-5 A
// producers (one tile row per thread), warps 6-9 epilogue.  Per tile the MMA warp runs
//   S1 (K1/32 k-blocks, A streamed) | E1 | Z (a resident, m streamed) | Ez | R | C1 (a resident)
//   | Er | C2 (rh resident) | E3
// (the gate math of Ez runs under R, that of Er under C1: only the accumulator read-out and the hi | lo
// split of an operand stand between two dependent GEMMs)
// with the E phases on the epilogue warps.  Weight k-blocks keep streaming through the phases.  In Z
// and R the resident and the streamed k-blocks alternate (a0 m0 a1 m1 ..), so a ring slot has two
// k-blocks of MMA time to be refilled.
#include <stdio.h>
#include <stdlib.h>

#include "tc_common.cuh"

namespace ruvgnn {
namespace tc {
namespace {

constexpr int BK = 32;
constexpr int HF = 128;  // hidden size this kernel is written for
constexpr uint32_t COL_X = 0, COL_Y = 128, COL_P = 256, COL_Q = 384;
constexpr int S_STAGES = 4;  // stage-1 ring (in P | Q)
constexpr int M_STAGES = 2;  // message-row ring (in Y)
constexpr int B_STAGES = 4;
constexpr int B_TILE_BYTES = HF * BK * 4;  // one 128-row x 32-float weight k-block
constexpr int B_STAGE_BYTES = 2 * B_TILE_BYTES;
constexpr int ZLD = 132;  // floats per private z/d row: 16-byte aligned, conflict-free float4 access
constexpr int NVEC = 6;   // bf | bz | br | bh | gamma | beta, 128 floats each

struct FusedParams {
  FusedUpdArgs a;
  int single_pass;
  int pf_dist;
  int num_tiles;
  int debug;  // timing experiments only (RUVGNN_FUSED_DEBUG): 1 = no A loads, 2 = no weight TMA, 4 = no MMAs,
              // 8 = no activation math / message loads in the epilogues.  Results are garbage.
              // 16 = phase timeline of CTA 0's third tile (clock64) into `trace`, printed by the launcher.
  unsigned long long *trace;
  int gate;  // RUVGNN_FUSED_GATE: the MMA warp waits for the previous tile's ring_free before it completes stage 1
};

__device__ __forceinline__ void trace_at(const FusedParams &q, bool on, int slot) {
  if (on) q.trace[slot] = (unsigned long long)clock64();
}

__device__ __forceinline__ uint32_t tf32_bits(float x) {
  return (__float_as_uint(x) + 0x1000u) & 0xFFFFE000u;
}
__device__ __forceinline__ void ld_row32(const float *src, float4 (&v)[8]) {
#pragma unroll
  for (int i = 0; i < 4; ++i)
    asm volatile("ld.global.nc.v8.f32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                 : "=f"(v[2 * i].x), "=f"(v[2 * i].y), "=f"(v[2 * i].z), "=f"(v[2 * i].w),
                   "=f"(v[2 * i + 1].x), "=f"(v[2 * i + 1].y), "=f"(v[2 * i + 1].z),
                   "=f"(v[2 * i + 1].w)
                 : "l"(src + 8 * i));
}

__global__ void __launch_bounds__(NUM_THREADS, 1)
upd_fused_kernel(const __grid_constant__ CUtensorMap mf_hi, const __grid_constant__ CUtensorMap mf_lo,
                 const __grid_constant__ CUtensorMap mzr_hi, const __grid_constant__ CUtensorMap mzr_lo,
                 const __grid_constant__ CUtensorMap mh_hi, const __grid_constant__ CUtensorMap mh_lo,
                 const FusedParams q) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  const FusedUpdArgs &a = q.a;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const bool single = q.single_pass != 0;

  uint8_t *base = reinterpret_cast<uint8_t *>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) &
                                              ~(uintptr_t)1023);
  auto b_hi = [&](int s) { return base + (size_t)s * B_STAGE_BYTES; };
  auto b_lo = [&](int s) { return base + (size_t)s * B_STAGE_BYTES + B_TILE_BYTES; };
  float *zrows = reinterpret_cast<float *>(base + (size_t)B_STAGES * B_STAGE_BYTES);
  float *vecs = zrows + BLOCK_M * ZLD;
  const float *s_bf = vecs, *s_bz = vecs + HF, *s_br = vecs + 2 * HF, *s_bh = vecs + 3 * HF;
  const float *s_gamma = vecs + 4 * HF, *s_beta = vecs + 5 * HF;
  uint64_t *bars = reinterpret_cast<uint64_t *>(vecs + NVEC * HF);
  uint64_t *full_b = bars, *empty_b = bars + B_STAGES;
  uint64_t *full_s = bars + 2 * B_STAGES, *empty_s = full_s + S_STAGES;
  uint64_t *full_m = empty_s + S_STAGES, *empty_m = full_m + M_STAGES;
  uint64_t *acc_bar = empty_m + M_STAGES;  // MMA -> epilogue: an accumulator phase is complete
  uint64_t *epi_bar = acc_bar + 1;         // epilogue -> MMA: the phase's results are in place
  uint64_t *ring_free = acc_bar + 2;       // epilogue -> A producers: Y holds no accumulator any more
  uint64_t *drain_bar = acc_bar + 3;       // MMA -> MMA: the MMAs that read the Y ring as A have completed
  uint64_t *pq_free = acc_bar + 4;         // MMA -> A producers: the last MMAs reading rh from P | Q have completed
  uint64_t *r_done = acc_bar + 5;          // MMA -> epilogue: the R GEMM is complete (C1 may still be running)
  uint32_t *tmem_ptr = reinterpret_cast<uint32_t *>(acc_bar + 6);

  for (int i = threadIdx.x; i < HF; i += NUM_THREADS) {
    vecs[i] = a.bf[i];
    vecs[HF + i] = a.bzr[i];
    vecs[2 * HF + i] = a.bzr[HF + i];
    vecs[3 * HF + i] = a.bh[i];
    vecs[4 * HF + i] = a.gamma[i];
    vecs[5 * HF + i] = a.beta[i];
  }
  if (warp == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&mf_hi)) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&mf_lo)) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&mzr_hi)) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&mzr_lo)) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&mh_hi)) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&mh_lo)) : "memory");
    for (int s = 0; s < B_STAGES; ++s) {
      mbar_init(&full_b[s], 1);
      mbar_init(&empty_b[s], 1);
    }
    for (int s = 0; s < S_STAGES; ++s) {
      mbar_init(&full_s[s], 4);  // one arrive per producer warp
      mbar_init(&empty_s[s], 1);
    }
    for (int s = 0; s < M_STAGES; ++s) {
      mbar_init(&full_m[s], 4);
      mbar_init(&empty_m[s], 1);
    }
    mbar_init(acc_bar, 1);
    mbar_init(epi_bar, 4);  // one arrive per epilogue warp
    mbar_init(ring_free, 4);
    mbar_init(drain_bar, 1);
    mbar_init(pq_free, 1);
    mbar_init(r_done, 1);
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc(tmem_ptr, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;

  const int NK1 = a.K1 / BK;   // streamed k-blocks of stage 1
  const int NA = NK1 + 8;      // A-ring k-blocks per tile: stage 1, then m for Z, m again for R
  const int NB = NK1 + 24;     // weight k-blocks per tile: Wf, Wzr rows [0,128), rows [128,256), Wh
  const int num_tiles = q.num_tiles;

  if (warp == 0) {
    // ===== TMA producer: the tile's weight k-blocks in the order the MMA warp consumes them =====
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        for (int j = 0; j < NB; ++j) {
          const CUtensorMap *mh, *ml;
          int c0, c1 = 0;
          if (j < NK1) {
            mh = &mf_hi; ml = &mf_lo; c0 = j * BK;
          } else if (j < NK1 + 16) {
            const int jj = j - NK1, t = jj & 7;  // a0 m0 a1 m1 a2 m2 a3 m3: k-blocks 0 4 1 5 2 6 3 7
            mh = &mzr_hi; ml = &mzr_lo; c0 = ((t >> 1) + (t & 1) * 4) * BK; c1 = (jj >> 3) * HF;
          } else {
            mh = &mh_hi; ml = &mh_lo; c0 = (j - NK1 - 16) * BK;
          }
          mbar_wait(&empty_b[stage], phase ^ 1);
          if (q.debug & 2) {
            mbar_arrive(&full_b[stage]);
          } else {
            mbar_arrive_expect_tx(&full_b[stage], single ? B_TILE_BYTES : 2 * B_TILE_BYTES);
            tma_load_2d(b_hi(stage), mh, &full_b[stage], c0, c1);
            if (!single) tma_load_2d(b_lo(stage), ml, &full_b[stage], c0, c1);
          }
          if (++stage == B_STAGES) {
            stage = 0;
            phase ^= 1;
          }
        }
      }
    }
  } else if (warp == 1) {
    // ===== MMA issuer =====
    if (lane == 0) {
      const uint32_t idesc = make_idesc_tf32(HF);
      const uint32_t dX = tmem_base + COL_X, dY = tmem_base + COL_Y;
      const uint32_t tP = tmem_base + COL_P, tQ = tmem_base + COL_Q;
      int ss = 0, sm = 0, sb = 0;
      uint32_t ps = 0, pm = 0, pb = 0, pe = 0, pd = 0;
      uint32_t tiles_issued = 0;
      // one weight k-block against one A k-block (hi at ah, remainder at al)
      auto kblock = [&](uint32_t d, uint32_t ah, uint32_t al, bool first_kb) {
        mbar_wait(&full_b[sb], pb);
        tc_fence_after();
        const uint64_t bh = make_desc<BK>(smem_u32(b_hi(sb)));
        const uint64_t bl = make_desc<BK>(smem_u32(b_lo(sb)));
#pragma unroll
        for (int kk = 0; kk < BK / 8; ++kk) {
          const uint64_t off = (uint64_t)((kk * 32) >> 4);
          const uint32_t acc = (first_kb && kk == 0) ? 0u : 1u;
          if (q.debug & 4) {
          } else if (single) {
            umma_tf32_ts(d, ah + kk * 8, bh + off, idesc, acc);
          } else {
            umma_tf32_ts(d, al + kk * 8, bh + off, idesc, acc);
            umma_tf32_ts(d, ah + kk * 8, bl + off, idesc, 1u);
            umma_tf32_ts(d, ah + kk * 8, bh + off, idesc, 1u);
          }
        }
        umma_commit(&empty_b[sb]);
        if (++sb == B_STAGES) {
          sb = 0;
          pb ^= 1;
        }
      };
      auto s_kblock = [&](uint32_t d, bool first_kb) {  // next slot of the stage-1 ring
        mbar_wait(&full_s[ss], ps);
        const uint32_t ah = tmem_base + COL_P + (uint32_t)(ss * 64);
        kblock(d, ah, ah + 32, first_kb);
        umma_commit(&empty_s[ss]);
        if (++ss == S_STAGES) {
          ss = 0;
          ps ^= 1;
        }
      };
      auto m_kblock = [&](uint32_t d) {  // next slot of the message-row ring
        mbar_wait(&full_m[sm], pm);
        const uint32_t ah = tmem_base + COL_Y + (uint32_t)(sm * 64);
        kblock(d, ah, ah + 32, false);
        umma_commit(&empty_m[sm]);
        if (++sm == M_STAGES) {
          sm = 0;
          pm ^= 1;
        }
      };
      auto wait_epilogue = [&]() {
        mbar_wait(epi_bar, pe);
        pe ^= 1;
        tc_fence_after();
      };
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        const bool tr = (q.debug & 16) && tile == 2 * (int)gridDim.x;
        trace_at(q, tr, 0);
        for (int kb = 0; kb < NK1; ++kb) s_kblock(dX, kb == 0);  // S1: X = [S | agg] Wf^T
        // Stage 1 of this tile ran under the previous tile's last epilogue.  acc_bar is a two-phase (parity) barrier:
        // its next completion must not happen before every epilogue warp has OBSERVED the previous one (C2 of the
        // previous tile), or a warp that stalled for a whole stage 1 (~19 us) would wait on the wrong phase
        // (tests/test_fused_protocol.py finds that schedule).  ring_free of the previous tile is arrived by all four
        // epilogue warps after they passed that wait.  Opt-in (RUVGNN_FUSED_GATE=1) this round: written after the
        // round's GPU budget was spent, so the default instruction stream is the validated one; a spinning warp
        // does not stall that long on this hardware, and every mbarrier wait traps after 4 s rather than hanging.
        if (q.gate && tiles_issued > 0) mbar_wait(ring_free, (tiles_issued - 1) & 1);
        ++tiles_issued;
        umma_commit(acc_bar);
        trace_at(q, tr, 1);
        for (int half = 0; half < 2; ++half) {  // Z, then R: X = [a | m] Wzr[half]^T
          wait_epilogue();                      // E1: a in P | Q (half 0);  Ez: z taken out of X (half 1)
          trace_at(q, tr, 2 + 2 * half);
          for (int kb = 0; kb < 4; ++kb) {
            kblock(dX, tP + kb * 32, tQ + kb * 32, kb == 0);
            m_kblock(dX);
          }
          umma_commit(half == 0 ? acc_bar : r_done);
          trace_at(q, tr, 3 + 2 * half);
        }
        // C1: Y = a Wh[:, 0:128]^T.  Y was the message-row ring until here; the producers hold until
        // ring_free, and the MMAs that read it as their A operand must have completed before another MMA
        // writes it as its accumulator (only MMAs on the SAME accumulator are ordered by the pipeline).
        umma_commit(drain_bar);
        mbar_wait(drain_bar, pd);
        pd ^= 1;
        tc_fence_after();
        trace_at(q, tr, 6);
        for (int kb = 0; kb < 4; ++kb) kblock(dY, tP + kb * 32, tQ + kb * 32, kb == 0);
        umma_commit(acc_bar);  // r complete in X, a no longer needed
        trace_at(q, tr, 7);
        wait_epilogue();       // Er: rh in P | Q
        trace_at(q, tr, 8);
        for (int kb = 0; kb < 4; ++kb) kblock(dY, tP + kb * 32, tQ + kb * 32, false);  // C2
        umma_commit(acc_bar);
        trace_at(q, tr, 9);
        umma_commit(pq_free);  // the next tile's gather-state k-blocks may stream into P | Q
      }
    }
  } else if (warp < 6) {
    // ===== A producers: one tile row per thread -> (hi, lo) -> TMEM ring =====
    const int quad = warp & 3;
    const int row_in_tile = quad * 32 + lane;
    const uint32_t lane_base = (uint32_t)(quad * 32) << 16;
    int ss = 0, sm = 0;
    uint32_t ps = 0, pm = 0;
    uint32_t tiles_done = 0;
    float4 cur[8], nxt[8];
    auto src_of = [&](int tile, int i) -> const float * {
      const int64_t r = (int64_t)tile * BLOCK_M + row_in_tile;
      if (r >= a.rows) return nullptr;
      const int64_t row = (int64_t)a.row_begin + r;
      return i < NK1 ? a.A1 + row * a.lda1 + (int64_t)i * BK
                     : a.m + row * a.ldm + (int64_t)((i - NK1) & 3) * BK;
    };
    auto issue = [&](float4 (&v)[8], int tile, int i) {
      const float *src = src_of(tile, i);
      if (src && !(q.debug & 1)) {
        ld_row32(src, v);
      } else {
#pragma unroll
        for (int k = 0; k < 8; ++k) v[k] = make_float4(0.f, 0.f, 0.f, 0.f);
      }
    };
    auto prefetch_ahead = [&](int tile, int i) {
      int t = tile, k = i + q.pf_dist;
      while (k >= NA) {
        k -= NA;
        t += gridDim.x;
      }
      if (t >= num_tiles || k >= NK1 + 4) return;  // the second pass over m hits L2 anyway
      const float *src = src_of(t, k);
      if (src) prefetch_l2(src);
    };
    int tile = blockIdx.x, i = 0;
    if (tile < num_tiles) issue(nxt, tile, 0);
    if (q.pf_dist > 0 && tile < num_tiles)
      for (int k = 1 - q.pf_dist; k < 0; ++k) prefetch_ahead(tile, k);
    while (tile < num_tiles) {
#pragma unroll
      for (int k = 0; k < 8; ++k) cur[k] = nxt[k];
      int ntile = tile, ni = i + 1;
      if (ni == NA) {
        ni = 0;
        ntile = tile + gridDim.x;
      }
      if (ntile < num_tiles) issue(nxt, ntile, ni);
      if (q.pf_dist > 0) prefetch_ahead(tile, i);
      // Both rings live in columns that hold something else at the end of the previous tile: P | Q the
      // candidate GEMM's rh operand (until its last MMAs have completed), Y its accumulator (until the last
      // epilogue has read it).
      uint64_t *full;
      uint32_t col;
      if (i < NK1) {
        mbar_wait(&empty_s[ss], ps ^ 1);
        if (i == 0 && tiles_done > 0) mbar_wait(pq_free, (tiles_done - 1) & 1);
        full = &full_s[ss];
        col = COL_P + (uint32_t)(ss * 64);
        if (++ss == S_STAGES) {
          ss = 0;
          ps ^= 1;
        }
      } else {
        mbar_wait(&empty_m[sm], pm ^ 1);
        if (i == NK1 && tiles_done > 0) mbar_wait(ring_free, (tiles_done - 1) & 1);
        full = &full_m[sm];
        col = COL_Y + (uint32_t)(sm * 64);
        if (++sm == M_STAGES) {
          sm = 0;
          pm ^= 1;
        }
      }
      tc_fence_after();
      const uint32_t ah = tmem_base + lane_base + col;
      uint32_t r[32];
      const float *cf = reinterpret_cast<const float *>(cur);
#pragma unroll
      for (int k = 0; k < 32; ++k) r[k] = tf32_bits(cf[k]);
      tmem_st32(ah, r);
      if (!single) {
#pragma unroll
        for (int k = 0; k < 32; ++k) r[k] = __float_as_uint(cf[k] - __uint_as_float(r[k]));
        tmem_st32(ah + 32, r);
      }
      tmem_st_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(full);
      if (ni == 0) ++tiles_done;
      tile = ntile;
      i = ni;
    }
  } else {
    // ===== epilogue warps: the phases between the GEMMs, one tile row per thread =====
    const int quad = warp & 3;
    const int row_in_tile = quad * 32 + lane;
    const uint32_t lane_base = (uint32_t)(quad * 32) << 16;
    const uint32_t tX = tmem_base + lane_base + COL_X, tY = tmem_base + lane_base + COL_Y;
    const uint32_t tP = tmem_base + lane_base + COL_P, tQ = tmem_base + lane_base + COL_Q;
    float *zrow = zrows + row_in_tile * ZLD;
    uint32_t pacc = 0, pr = 0;
    auto wait_acc = [&]() {
      mbar_wait(acc_bar, pacc);
      pacc ^= 1;
      tc_fence_after();
    };
    auto signal = [&](uint64_t *bar) {
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(bar);
    };
    // v -> A operand columns [c, c + 32) of P | Q
    auto store_operand = [&](int c, const float (&v)[32]) {
      uint32_t r[32];
#pragma unroll
      for (int j = 0; j < 32; ++j) r[j] = tf32_bits(v[j]);
      tmem_st32(tP + (uint32_t)c, r);
      if (!single) {
#pragma unroll
        for (int j = 0; j < 32; ++j) r[j] = __float_as_uint(v[j] - __uint_as_float(r[j]));
        tmem_st32(tQ + (uint32_t)c, r);
      }
    };
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
      const int64_t r_local = (int64_t)tile * BLOCK_M + row_in_tile;
      const bool valid = r_local < a.rows;
      const int64_t row = (int64_t)a.row_begin + r_local;
      const float *mrow = a.m + (valid ? row : (int64_t)a.row_begin) * a.ldm;
      const bool iso = valid && a.row_ptr && a.row_ptr[row] == a.row_ptr[row + 1];

      const bool tr = (q.debug & 16) && tile == 2 * (int)gridDim.x && warp == 6 && lane == 0;
      // ---- E1: a = acc + bf  ->  P | Q
      wait_acc();
      trace_at(q, tr, 16);
#pragma unroll 1
      for (int c = 0; c < HF; c += 32) {
        float v[32];
        tmem_ld32(tX + (uint32_t)c, v);
#pragma unroll
        for (int j = 0; j < 32; ++j) v[j] += s_bf[c + j];
        store_operand(c, v);
      }
      tmem_st_wait();
      signal(epi_bar);
      trace_at(q, tr, 17);

      // ---- Ez: z = sigmoid(acc + b_z)  ->  private shared row.  The accumulator is copied out first and X handed
      // back at once, so the R GEMM runs under the sigmoids (they work on the copy, in place).
      wait_acc();
      trace_at(q, tr, 18);
#pragma unroll 1
      for (int c = 0; c < HF; c += 32) {
        float v[32];
        tmem_ld32(tX + (uint32_t)c, v);
#pragma unroll
        for (int j = 0; j < 8; ++j)
          *reinterpret_cast<float4 *>(zrow + c + 4 * j) = make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
      }
      signal(epi_bar);
      trace_at(q, tr, 19);
      if (!(q.debug & 8)) {
#pragma unroll 4
        for (int j = 0; j < HF / 4; ++j) {
          float4 z = *reinterpret_cast<const float4 *>(zrow + 4 * j);
          z.x = sigmoid_ref(z.x + s_bz[4 * j + 0]);
          z.y = sigmoid_ref(z.y + s_bz[4 * j + 1]);
          z.z = sigmoid_ref(z.z + s_bz[4 * j + 2]);
          z.w = sigmoid_ref(z.w + s_bz[4 * j + 3]);
          *reinterpret_cast<float4 *>(zrow + 4 * j) = z;
        }
      }

      // ---- Er: rh = sigmoid(acc + b_r) * m  ->  P | Q.  r is complete as soon as the R GEMM is (r_done), but
      // P | Q still hold a until C1 has read it: the gate math runs under C1 and parks rh in X (the thread's own
      // TMEM row, free until the next tile); once C1 has completed only the split into P | Q remains.
      mbar_wait(r_done, pr);
      pr ^= 1;
      tc_fence_after();
      trace_at(q, tr, 20);
      if (!(q.debug & 8)) {
#pragma unroll 1
        for (int c = 0; c < HF; c += 32) {
          float4 mm[8];
          ld_row32(mrow + c, mm);
          float v[32];
          tmem_ld32(tX + (uint32_t)c, v);
          const float *mf = reinterpret_cast<const float *>(mm);
          uint32_t r[32];
#pragma unroll
          for (int j = 0; j < 32; ++j) r[j] = __float_as_uint(sigmoid_ref(v[j] + s_br[c + j]) * mf[j]);
          tmem_st32(tX + (uint32_t)c, r);
        }
        tmem_st_wait();
      }
      wait_acc();  // C1 complete: a is dead
#pragma unroll 1
      for (int c = 0; c < HF; c += 32) {
        float v[32];
        tmem_ld32(tX + (uint32_t)c, v);
        store_operand(c, v);
      }
      tmem_st_wait();
      signal(epi_bar);
      trace_at(q, tr, 21);

      // ---- E3: d = ((1 - z) m + z tanh(acc + b_h)) * keep, LayerNorm, store
      wait_acc();
      trace_at(q, tr, 22);
#pragma unroll 1
      for (int c = 0; c < HF; c += 32) {
        float v[32];
        if (q.debug & 8) {
          tmem_ld32(tY + (uint32_t)c, v);
#pragma unroll
          for (int j = 0; j < 8; ++j)
            *reinterpret_cast<float4 *>(zrow + c + 4 * j) = make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
          continue;
        }
        float4 mm[8];
        ld_row32(mrow + c, mm);
        tmem_ld32(tY + (uint32_t)c, v);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const float4 z = *reinterpret_cast<const float4 *>(zrow + c + 4 * j);
          const float4 m = mm[j];
          float4 d;
          d.x = ((1.0f - z.x) * m.x + z.x * tanh_ref(v[4 * j + 0] + s_bh[c + 4 * j + 0])) * a.keep;
          d.y = ((1.0f - z.y) * m.y + z.y * tanh_ref(v[4 * j + 1] + s_bh[c + 4 * j + 1])) * a.keep;
          d.z = ((1.0f - z.z) * m.z + z.z * tanh_ref(v[4 * j + 2] + s_bh[c + 4 * j + 2])) * a.keep;
          d.w = ((1.0f - z.w) * m.w + z.w * tanh_ref(v[4 * j + 3] + s_bh[c + 4 * j + 3])) * a.keep;
          if (iso) d = m;  // no neighbours: LN(W_msg x + b), layer.rs:385-389
          *reinterpret_cast<float4 *>(zrow + c + 4 * j) = d;
        }
      }
      signal(ring_free);  // Y has been read: the producers may stream the next tile's message rows into it
      trace_at(q, tr, 23);
      // LayerNorm.  Every thread normalises its own row in place; the sums follow the association of
      // layernorm_fast_kernel<1> (rowops.cu) -- lane partials (x0 + x1) + (x2 + x3), then the xor-butterfly
      // tree 16, 8, 4, 2, 1 -- so the rows this kernel writes carry the bits that kernel would give them
      // (sparse levels mix the two), without a shuffle chain per row.
      {
        float p[32];
#pragma unroll
        for (int j = 0; j < 32; ++j) {
          const float4 x = *reinterpret_cast<const float4 *>(zrow + 4 * j);
          p[j] = (x.x + x.y) + (x.z + x.w);
        }
#pragma unroll
        for (int w = 16; w > 0; w >>= 1)
#pragma unroll
          for (int j = 0; j < w; ++j) p[j] = p[j] + p[j + w];
        const float mean = p[0] * (1.0f / HF);
#pragma unroll
        for (int j = 0; j < 32; ++j) {
          float4 x = *reinterpret_cast<const float4 *>(zrow + 4 * j);
          x.x -= mean; x.y -= mean; x.z -= mean; x.w -= mean;
          p[j] = fmaf(x.x, x.x, fmaf(x.y, x.y, fmaf(x.z, x.z, fmaf(x.w, x.w, 0.f))));
        }
#pragma unroll
        for (int w = 16; w > 0; w >>= 1)
#pragma unroll
          for (int j = 0; j < w; ++j) p[j] = p[j] + p[j + w];
        const float denom = sqrtf(p[0] * (1.0f / HF) + a.eps);
#pragma unroll 4
        for (int j = 0; j < 32; ++j) {
          float4 x = *reinterpret_cast<const float4 *>(zrow + 4 * j);
          const float4 g = *reinterpret_cast<const float4 *>(s_gamma + 4 * j);
          const float4 b = *reinterpret_cast<const float4 *>(s_beta + 4 * j);
          x.x -= mean; x.y -= mean; x.z -= mean; x.w -= mean;
          float4 o4;
          o4.x = g.x * (x.x / denom) + b.x;
          o4.y = g.y * (x.y / denom) + b.y;
          o4.z = g.z * (x.z / denom) + b.z;
          o4.w = g.w * (x.w / denom) + b.w;
          *reinterpret_cast<float4 *>(zrow + 4 * j) = o4;
        }
      }
      const uint64_t dst_row = valid ? (a.scatter ? (uint64_t)a.scatter[row] : (uint64_t)row) : 0;
      __syncwarp();  // the slab's rows are visible to the whole warp
      // copy-out: one warp instruction moves one row of the slab (512 contiguous bytes)
      const float *slab = zrows + (quad * 32) * ZLD;
      const int64_t rows_left = a.rows - ((int64_t)tile * BLOCK_M + quad * 32);
      const int nrows = rows_left > 32 ? 32 : (rows_left > 0 ? (int)rows_left : 0);
#pragma unroll 4
      for (int r = 0; r < nrows; ++r) {
        const uint64_t dr = __shfl_sync(0xffffffffu, dst_row, r);
        const float4 x = *reinterpret_cast<const float4 *>(slab + r * ZLD + 4 * lane);
        *reinterpret_cast<float4 *>(a.out + dr * (uint64_t)HF + 4 * lane) = x;
      }
      trace_at(q, tr, 24);
      __syncwarp();  // before the next tile's Ez overwrites the slab
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

constexpr size_t kFusedSmem = 1024 + (size_t)B_STAGES * B_STAGE_BYTES +
                              (size_t)BLOCK_M * ZLD * sizeof(float) + NVEC * HF * sizeof(float) + 256;
static_assert(kFusedSmem <= 227 * 1024, "fused K-upd shared memory");

}  // namespace

bool upd_fused_supported(const TcWeights &w, uint32_t H) {
  const TcMaps *maps = static_cast<const TcMaps *>(w.tmaps);
  if (!w.usable || !maps || H != (uint32_t)HF) return false;
  const int ids[3] = {TC_W_F, TC_W_ZR, TC_W_H};
  for (int i : ids)
    if (!maps->ts_ok[i] || maps->block_n3[i] != HF) return false;
  return w.rows[TC_W_F] == (uint32_t)HF && w.cols[TC_W_F] % BK == 0 && w.rows[TC_W_ZR] == 2u * HF &&
         w.cols[TC_W_ZR] == 2u * HF && w.rows[TC_W_H] == (uint32_t)HF && w.cols[TC_W_H] == 2u * HF;
}

cudaError_t launch_upd_fused(const FusedUpdArgs &a, const TcWeights &w, bool single_pass, int sm_count,
                             cudaStream_t s, uint64_t *launches) {
  if (a.rows <= 0) return cudaSuccess;
  if (!upd_fused_supported(w, HF) || (uint32_t)a.K1 != w.cols[TC_W_F]) return cudaErrorInvalidValue;
  // the row-per-thread 256-bit loads and the 128-bit row stores need 32- / 16-byte aligned rows
  if (((reinterpret_cast<uintptr_t>(a.A1) | reinterpret_cast<uintptr_t>(a.m)) & 31) != 0 ||
      (reinterpret_cast<uintptr_t>(a.out) & 15) != 0 || (a.lda1 & 7) != 0 || (a.ldm & 7) != 0)
    return cudaErrorMisalignedAddress;
  const TcMaps *maps = static_cast<const TcMaps *>(w.tmaps);
  static bool configured_dev[64] = {};
  int dev = 0;
  cudaGetDevice(&dev);
  bool &configured = configured_dev[dev & 63];
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(upd_fused_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         (int)kFusedSmem);
    if (e != cudaSuccess) return e;
    configured = true;
  }
  FusedParams q;
  q.a = a;
  q.single_pass = single_pass ? 1 : 0;
  q.pf_dist = tc_pf_dist();
  {
    static int dbg = -1;
    if (dbg < 0) {
      const char *e = getenv("RUVGNN_FUSED_DEBUG");
      dbg = e ? atoi(e) : 0;
    }
    q.debug = dbg;
    static int gate = -1;
    if (gate < 0) {
      const char *e = getenv("RUVGNN_FUSED_GATE");
      gate = e ? atoi(e) : 0;
    }
    q.gate = gate;
  }
  const int64_t tiles = (a.rows + BLOCK_M - 1) / BLOCK_M;
  q.num_tiles = (int)tiles;
  const int grid = (int)(tiles < sm_count ? tiles : sm_count);
  if (launches) ++*launches;
  q.trace = nullptr;
  if (q.debug & 16) {
    static unsigned long long *buf = nullptr;
    if (!buf && cudaMalloc(&buf, 32 * sizeof(unsigned long long)) != cudaSuccess) return cudaErrorMemoryAllocation;
    cudaMemsetAsync(buf, 0, 32 * sizeof(unsigned long long), s);
    q.trace = buf;
  }
  upd_fused_kernel<<<grid, NUM_THREADS, kFusedSmem, s>>>(maps->hi3[TC_W_F], maps->lo3[TC_W_F],
                                                         maps->hi3[TC_W_ZR], maps->lo3[TC_W_ZR],
                                                         maps->hi3[TC_W_H], maps->lo3[TC_W_H], q);
  if (q.debug & 16) {
    unsigned long long h[32];
    cudaStreamSynchronize(s);
    cudaMemcpy(h, q.trace, sizeof h, cudaMemcpyDeviceToHost);
    fprintf(stderr, "upd_fused trace (cycles from S1 start; MMA thread 0-9, epilogue 16-24):");
    for (int i = 0; i < 32; ++i)
      if (h[i]) fprintf(stderr, " [%d]=%lld", i, (long long)(h[i] - h[0]));
    fprintf(stderr, "\n");
  }
  return cudaGetLastError();
}

}  // namespace tc
}  // namespace ruvgnn
