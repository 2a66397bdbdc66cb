// Shared device helpers of the tcgen05 GEMM kernels (operands in shared memory / A operand in TMEM).
#pragma once
#include <cuda.h>

#include "gemm_tc.cuh"

namespace ruvgnn {
namespace tc {

constexpr int BLOCK_M = 128;
constexpr int NUM_THREADS = 320;
constexpr int MAX_STAGES = 8;

struct TcParams {
  GemmArgs g;
  int block_n;
  int stages;
  int single_pass;
  int num_m_tiles, num_n_tiles;
  int pf_dist;  // A-operand L2 prefetch distance in k-blocks (0: off).  Short-K GEMMs (H = 128) keep too few
                // bytes in flight per SM with one k-block of register prefetch (16 KB against the ~43 KB that
                // 6.4 TB/s x ~1 us needs): pulling the rows a few k-blocks ahead into L2 turns the HBM latency
                // the producers see into L2 latency.
  int debug;  // timing experiments only (RUVGNN_TC_DEBUG): 1 = skip A loads/stores, 2 = skip weight TMA,
              // 4 = skip the MMAs, 8 = skip the epilogue's global traffic.  Results are garbage.
};

// ---- PTX wrappers ------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void *p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}
__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_arrive(uint64_t *bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t *bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)),
               "r"(bytes)
               : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t *bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t"
      "}"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}
// A protocol bug must fail fast, never hang the device: trap after 4 s of waiting.
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
  if (mbar_try_wait(bar, parity)) return;
  uint64_t t0;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
  while (!mbar_try_wait(bar, parity)) {
    uint64_t t1;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t1));
    if (t1 - t0 > 4000000000ull) __trap();
  }
}
__device__ __forceinline__ void fence_barrier_init() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void fence_proxy_async() {
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void tc_fence_before() {
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
}
__device__ __forceinline__ void tc_fence_after() {
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
}
__device__ __forceinline__ void tma_load_2d(void *dst, const CUtensorMap *map, uint64_t *bar,
                                            int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, "
      "{%3, %4}], [%2];" ::"r"(smem_u32(dst)),
      "l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void tmem_alloc(uint32_t *dst_smem, uint32_t cols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(
                   smem_u32(dst_smem)),
               "r"(cols)
               : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t addr, uint32_t cols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(addr), "r"(cols)
               : "memory");
}
// D[tmem] (+)= A[smem] * B[smem]^T, kind::tf32, FP32 accumulate
__device__ __forceinline__ void umma_tf32(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc,
                                          uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t"
      "}" ::"r"(tmem_d),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t *bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(
                   smem_u32(bar))
               : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, float (&v)[32]) {
  uint32_t r[32];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]),
        "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]),
        "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]),
        "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]),
        "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr)
      : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
  for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(r[i]);
}

// ---- A operand in TMEM (tcgen05.mma TS form) ----
__device__ __forceinline__ void tmem_st32(uint32_t taddr, const uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, "
      "%17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};" ::"r"(taddr),
      "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]),
      "r"(r[8]), "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]),
      "r"(r[16]), "r"(r[17]), "r"(r[18]), "r"(r[19]), "r"(r[20]), "r"(r[21]), "r"(r[22]),
      "r"(r[23]), "r"(r[24]), "r"(r[25]), "r"(r[26]), "r"(r[27]), "r"(r[28]), "r"(r[29]),
      "r"(r[30]), "r"(r[31])
      : "memory");
}
__device__ __forceinline__ void tmem_st_wait() {
  asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
}
// D[tmem] (+)= A[tmem] * B[smem]^T
__device__ __forceinline__ void umma_tf32_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t bdesc,
                                             uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t"
      "}" ::"r"(tmem_d),
      "r"(tmem_a), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}

// K-major swizzled operand tile: rows of BK*4 bytes (128 or 64), 8-row groups contiguous.
// (cute::UMMA::SmemDescriptor: start>>4 [0,14), LBO>>4 [16,30), SBO>>4 [32,46),
//  version=1 [46,48), layout [61,64): SWIZZLE_128B=2, SWIZZLE_64B=4.)
template <int BK>
__device__ __forceinline__ uint64_t make_desc(uint32_t smem_addr) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr >> 4) & 0x3FFF);
  d |= (uint64_t)1 << 16;                   // LBO (ignored for swizzled K-major), canonical 1
  d |= (uint64_t)((8 * BK * 4) >> 4) << 32;  // SBO = bytes between 8-row groups
  d |= (uint64_t)1 << 46;                   // descriptor version (Blackwell)
  d |= (uint64_t)(BK == 32 ? 2 : 4) << 61;  // SWIZZLE_128B / SWIZZLE_64B
  return d;
}
// Byte offset of 16-byte chunk `chunk` of row `r` inside a swizzled tile (what TMA produces
// for CU_TENSOR_MAP_SWIZZLE_128B / _64B: address bits [4,7) ^= [7,10), resp. [4,6) ^= [7,9)).
template <int BK>
__device__ __forceinline__ uint32_t swz_off(int r, int chunk) {
  if (BK == 32) return (uint32_t)r * 128u + (uint32_t)((chunk ^ (r & 7)) << 4);
  return (uint32_t)r * 64u + (uint32_t)((chunk ^ ((r >> 1) & 3)) << 4);
}
// cute::UMMA::InstrDescriptor: D=F32 (1<<4), A=B=TF32 (2<<7, 2<<10), K-major both,
// N>>3 at [17,23), M>>4 at [24,29).
__device__ __forceinline__ uint32_t make_idesc_tf32(int n) {
  return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(n >> 3) << 17) |
         ((uint32_t)(BLOCK_M >> 4) << 24);
}

// Gate activations of the tensor-core path.  The epilogue warps run them on 128 x BN outputs per
// tile (one warp per scheduler), and with libm expf / tanhf and IEEE divides (~45 instructions per
// element, dependent chains) that is what bounds the short-K shapes (H = 128: the gates epilogue of
// a tile outlasts its 8-k-block main loop).  MUFU forms: exp2 (ex2.approx, 2^-22 relative) and
// reciprocal (rcp.approx, 1 ulp) -- ~7 instructions per sigmoid, ~9 per tanh, absolute error
// < 5e-7, two orders below the 1e-4 parity bound.  -DRUVGNN_PRECISE_ACT restores libm; the FP32
// SIMT mode (gemm_simt.cu) always uses libm.
__device__ __forceinline__ float ex2_fast(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ float rcp_fast(float x) {
  float y;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

__device__ __forceinline__ float sigmoid_ref(float x) {
  // GRUCell::sigmoid (layer.rs:278-285): x > 0 ? 1/(1+exp(-x)) : exp(x)/(1+exp(x)).  Both
  // branches use exp(-|x|); written branch-free with one exp and one reciprocal.
#ifdef RUVGNN_PRECISE_ACT
  const float t = expf(-fabsf(x));
  const float r = 1.0f / (1.0f + t);
#else
  const float t = ex2_fast(-1.4426950408889634f * fabsf(x));
  const float r = rcp_fast(1.0f + t);
#endif
  return x > 0.f ? r : t * r;
}

__device__ __forceinline__ float tanh_ref(float x) {
  // GRUCell candidate state (layer.rs:268-270): tanh(x) = sign(x) (1 - e) / (1 + e), e = exp(-2|x|)
#ifdef RUVGNN_PRECISE_ACT
  return tanhf(x);
#else
  const float e = ex2_fast(-2.8853900817779268f * fabsf(x));
  return copysignf((1.0f - e) * rcp_fast(1.0f + e), x);
#endif
}

// Fused epilogue of one 32-row x 32-column accumulator chunk owned by one warp.
// tcgen05.ld hands every lane one ROW (32 consecutive columns in registers); touching global
// memory in that shape costs 32 wavefronts per access.  The chunk is therefore transposed
// through a padded per-warp shared tile (row stride 36 floats: conflict-free 128-bit row
// writes and 32-bit column reads) so that lane l owns COLUMN col+l and every global load /
// store of the warp is one coalesced 128-byte row segment.
constexpr int EPI_TILE_FLOATS = 32 * 36;  // per epilogue warp

template <int EPI>
__device__ __forceinline__ void epilogue_chunk(const GemmArgs &g, int64_t row0, int col,
                                               const float (&v)[32], float *stile, int lane) {
  __syncwarp();  // the previous chunk's reads are done
  float4 *dst = reinterpret_cast<float4 *>(stile + lane * 36);
#pragma unroll
  for (int j = 0; j < 8; ++j) dst[j] = make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
  __syncwarp();
  // 8 lanes x float4 cover the 32 columns of one row (128 contiguous bytes); one warp
  // instruction covers 4 rows; 8 passes cover the 32 rows.  Same instruction count as the
  // row-wise form, 8x fewer memory wavefronts.
  const int rs = lane >> 3, c4 = (lane & 7) * 4;
  const int c = col + c4;
  const float4 bias = __ldg(reinterpret_cast<const float4 *>(g.bias + c));
  const int64_t rows64 = g.M - row0;
  const int rows = rows64 > 32 ? 32 : (int)rows64;
  float4 acc[8], x1[8], x2[8];
#pragma unroll
  for (int p = 0; p < 8; ++p) {  // every load of the chunk is issued before the first store
    const int r = p * 4 + rs;
    const float4 a = *reinterpret_cast<const float4 *>(stile + r * 36 + c4);
    acc[p] = make_float4(a.x + bias.x, a.y + bias.y, a.z + bias.z, a.w + bias.w);
    x1[p] = make_float4(0.f, 0.f, 0.f, 0.f);
    x2[p] = x1[p];
    if (r < rows) {
      if (EPI == EPI_BIAS_ADD)
        x1[p] = __ldg(reinterpret_cast<const float4 *>(g.add + (row0 + r) * g.ldadd + c));
      if (EPI == EPI_GATES && col >= g.H)
        x1[p] = __ldg(reinterpret_cast<const float4 *>(g.m + (row0 + r) * g.ldm + (c - g.H)));
      if (EPI == EPI_GRU) {
        x1[p] = __ldg(reinterpret_cast<const float4 *>(g.z + (row0 + r) * g.ldz + c));
        x2[p] = __ldg(reinterpret_cast<const float4 *>(g.m + (row0 + r) * g.ldm + c));
      }
    }
  }
#pragma unroll
  for (int p = 0; p < 8; ++p) {
    const int r = p * 4 + rs;
    if (r >= rows) continue;
    const float4 a = acc[p];
    if (EPI == EPI_BIAS) {
      *reinterpret_cast<float4 *>(g.out + (row0 + r) * g.ldo + c) = a;
    } else if (EPI == EPI_BIAS_ADD) {
      *reinterpret_cast<float4 *>(g.out + (row0 + r) * g.ldo + c) =
          make_float4(a.x + x1[p].x, a.y + x1[p].y, a.z + x1[p].z, a.w + x1[p].w);
    } else if (EPI == EPI_GATES) {
      const float4 sg = make_float4(sigmoid_ref(a.x), sigmoid_ref(a.y), sigmoid_ref(a.z), sigmoid_ref(a.w));
      if (col < g.H)
        *reinterpret_cast<float4 *>(g.z + (row0 + r) * g.ldz + c) = sg;
      else
        *reinterpret_cast<float4 *>(g.rh + (row0 + r) * g.ldrh + (c - g.H)) =
            make_float4(sg.x * x1[p].x, sg.y * x1[p].y, sg.z * x1[p].z, sg.w * x1[p].w);
    } else {  // EPI_GRU (layer.rs:266-274, 448-451)
      const float4 z = x1[p], m = x2[p];
      float4 o;
      o.x = ((1.0f - z.x) * m.x + z.x * tanh_ref(a.x)) * g.keep;
      o.y = ((1.0f - z.y) * m.y + z.y * tanh_ref(a.y)) * g.keep;
      o.z = ((1.0f - z.z) * m.z + z.z * tanh_ref(a.z)) * g.keep;
      o.w = ((1.0f - z.w) * m.w + z.w * tanh_ref(a.w)) * g.keep;
      *reinterpret_cast<float4 *>(g.out + (row0 + r) * g.ldo + c) = o;
    }
  }
}

// Row-wise form of the fused epilogue: every lane keeps the ROW tcgen05.ld gave it and
// accesses global memory with 128-bit loads/stores of its own row (uncoalesced across the
// warp, but few instructions).  Measured faster for the gate epilogue, slower for plain bias.
template <int EPI>
__device__ __forceinline__ void epilogue_chunk_rowwise(const GemmArgs &g, int64_t row, int col,
                                               float (&v)[32]) {
    const float4 *b4 = reinterpret_cast<const float4 *>(g.bias + col);
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const float4 b = __ldg(b4 + j);
      v[4 * j + 0] += b.x;
      v[4 * j + 1] += b.y;
      v[4 * j + 2] += b.z;
      v[4 * j + 3] += b.w;
    }
    if (EPI == EPI_BIAS) {
      float4 *o = reinterpret_cast<float4 *>(g.out + row * g.ldo + col);
#pragma unroll
      for (int j = 0; j < 8; ++j)
        o[j] = make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
    } else if (EPI == EPI_BIAS_ADD) {
      const float4 *a = reinterpret_cast<const float4 *>(g.add + row * g.ldadd + col);
      float4 *o = reinterpret_cast<float4 *>(g.out + row * g.ldo + col);
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const float4 x = a[j];
        o[j] = make_float4(v[4 * j] + x.x, v[4 * j + 1] + x.y, v[4 * j + 2] + x.z,
                           v[4 * j + 3] + x.w);
      }
    } else if (EPI == EPI_GATES) {
      if (col < g.H) {
        float4 *o = reinterpret_cast<float4 *>(g.z + row * g.ldz + col);
#pragma unroll
        for (int j = 0; j < 8; ++j)
          o[j] = make_float4(sigmoid_ref(v[4 * j]), sigmoid_ref(v[4 * j + 1]),
                             sigmoid_ref(v[4 * j + 2]), sigmoid_ref(v[4 * j + 3]));
      } else {
        const int cc = col - g.H;
        const float4 *m = reinterpret_cast<const float4 *>(g.m + row * g.ldm + cc);
        float4 *o = reinterpret_cast<float4 *>(g.rh + row * g.ldrh + cc);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const float4 x = m[j];
          o[j] = make_float4(sigmoid_ref(v[4 * j]) * x.x, sigmoid_ref(v[4 * j + 1]) * x.y,
                             sigmoid_ref(v[4 * j + 2]) * x.z, sigmoid_ref(v[4 * j + 3]) * x.w);
        }
      }
    } else {  // EPI_GRU
      const float4 *z4 = reinterpret_cast<const float4 *>(g.z + row * g.ldz + col);
      const float4 *m4 = reinterpret_cast<const float4 *>(g.m + row * g.ldm + col);
      float4 *o = reinterpret_cast<float4 *>(g.out + row * g.ldo + col);
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const float4 z = z4[j], m = m4[j];
        float4 r;
        r.x = ((1.0f - z.x) * m.x + z.x * tanh_ref(v[4 * j + 0])) * g.keep;
        r.y = ((1.0f - z.y) * m.y + z.y * tanh_ref(v[4 * j + 1])) * g.keep;
        r.z = ((1.0f - z.z) * m.z + z.z * tanh_ref(v[4 * j + 2])) * g.keep;
        r.w = ((1.0f - z.w) * m.w + z.w * tanh_ref(v[4 * j + 3])) * g.keep;
        o[j] = r;
      }
    }
  }


__device__ __forceinline__ void prefetch_l2(const void *p) {
  asm volatile("prefetch.global.L2 [%0];" ::"l"(p));
}
// RUVGNN_TC_PF: prefetch distance in k-blocks (default 4, 0 = off)
int tc_pf_dist();

// Host: tensor maps of a layer's weight matrices.
struct TcMaps {
  CUtensorMap hi[TC_W_COUNT], lo[TC_W_COUNT];    // box {block_k, block_n}
  CUtensorMap hi3[TC_W_COUNT], lo3[TC_W_COUNT];  // box {32, block_n3}: A-in-TMEM kernel
  int block_n[TC_W_COUNT];
  int block_n3[TC_W_COUNT];
  int block_k;
  bool ts_ok[TC_W_COUNT];
};

// Which epilogue kinds use the coalesced (smem-transposed) form: bit e = EpiKind e.
// Default from measurements on B200; RUVGNN_TC_EPI_MASK overrides (experiments).
int tc_epi_mask();

// A-in-TMEM (tcgen05.mma TS form) launcher, gemm_tc3.cu
int pick_block_n_ts(uint32_t n);
cudaError_t launch_gemm_tc_ts(const GemmArgs &g, EpiKind epi, const TcMaps &maps, int which,
                              bool single_pass, int sm_count, cudaStream_t s);

}  // namespace tc
}  // namespace ruvgnn
