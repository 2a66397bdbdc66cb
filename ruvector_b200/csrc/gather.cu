// K-agg: neighbour gather + per-head scaled-dot-product softmax aggregation
// + edge-weight aggregation, one warp per node.  HBM-bound.
//
// Replaces, for every node at once (reference: crates/ruvector-gnn/src/layer.rs):
//   split_heads :163-171 + scaled_dot_product_attention :174-210   -> o   (pre W_o)
//   aggregate_messages :421-445                                     -> agg
// Inputs are the projected rows Q | K | V | M written by the projection GEMMs.
// Per edge the kernel reads ONE contiguous 3H-float block (K_j | V_j | M_j):
// formulation R = 3 of SURVEY.md 8(d); algorithmic bytes per node
//   4H (Q) + 16 (row_ptr pair) + deg*(4 + 4 + 12H) + 8H (o, agg).
//
// Softmax is the online (running max / running sum) form so any degree works
// in one pass; it equals the reference's max-subtracted softmax up to f32
// rounding.  sum.max(1e-10) (:198) is kept.
#include "common.cuh"

namespace ruvgnn {
namespace {

__device__ __forceinline__ float4 ldg4(const float4 *p) { return __ldg(p); }

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// Normalised-weight rule of aggregate_messages (:426-432), per edge.
struct WeightRule {
  float wsum;
  float uniform;  // 1/len(w)
  int mode;       // 0: w / wsum, 1: uniform, 2: already normalised
  __device__ __forceinline__ float apply(float w) const {
    return mode == 0 ? w / wsum : (mode == 1 ? uniform : w);
  }
};

__device__ __forceinline__ WeightRule make_rule(const float *edge_w, uint64_t beg, uint64_t end,
                                                int lane, int prenormalized) {
  WeightRule r;
  r.wsum = 1.f;
  r.uniform = 1.0f / (float)(end - beg);
  if (prenormalized) {
    r.mode = 2;
    return r;
  }
  float s = 0.f;
  for (uint64_t e = beg + lane; e < end; e += 32) s += __ldg(edge_w + e);
  s = warp_sum(s);
  r.wsum = s;
  r.mode = s > 0.f ? 0 : 1;
  return r;
}

// Peer-direct addressing: the row of global node g inside its owner's table (row stride ld floats).
__device__ __forceinline__ const float *peer_row(const PeerTable &t, uint64_t g, int64_t ld) {
  int o = 0;
  while (o + 1 < t.world && g >= t.bounds[o + 1]) ++o;
  return t.base[o] + (g - t.bounds[o]) * (uint64_t)ld;
}

template <int P>
struct RowRegs {
  float4 k[P], v[P], m[P];
};

// P = H/128 float4 per lane per row; LH = 32/heads lanes per head; U = edges in flight.
template <int P, int LH, int U, bool PEER>
__global__ void __launch_bounds__(128) gather_fast_kernel(GatherArgs a, float inv_scale) {
  constexpr int H = 128 * P;
  constexpr int V = 32 * P;  // float4 per H-row
  __shared__ PeerTable s_peer;
  if (PEER) {
    if (threadIdx.x == 0) s_peer = *a.peer;
    __syncthreads();
  }
  const uint64_t own_lo = PEER ? s_peer.bounds[s_peer.rank] : 0;
  const int lane = threadIdx.x & 31;
  const uint64_t warp = ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const uint64_t nwarps = ((uint64_t)gridDim.x * blockDim.x) >> 5;

  for (uint64_t node = a.row_begin + warp; node < a.row_end; node += nwarps) {
    const uint64_t beg = __ldg(a.row_ptr + node), end = __ldg(a.row_ptr + node + 1);
    float4 *orow = reinterpret_cast<float4 *>(a.oa + node * (uint64_t)(2 * H));
    if (beg == end) {  // no neighbours: the update kernel takes the LN(m) path (:385-389)
      const float4 zero = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
      for (int i = 0; i < 2 * P; ++i) orow[i * 32 + lane] = zero;
      continue;
    }
    const float4 *qrow = reinterpret_cast<const float4 *>(a.P + node * (uint64_t)a.ldp);
    float4 q[P], acc[P], ag[P];
#pragma unroll
    for (int i = 0; i < P; ++i) {
      q[i] = ldg4(qrow + i * 32 + lane);
      acc[i] = make_float4(0.f, 0.f, 0.f, 0.f);
      ag[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    }
    const WeightRule rule = make_rule(a.edge_w, beg, end, lane, a.prenormalized);
    float mx = -INFINITY, l = 0.f;

    auto row_of = [&](uint32_t col) -> const float4 * {
      if (PEER) {
        const uint64_t l = (uint64_t)col - own_lo;  // own row when inside [lo, hi)
        return l < a.n_local ? reinterpret_cast<const float4 *>(a.P + l * a.ldp + H)
                             : reinterpret_cast<const float4 *>(peer_row(s_peer, col, a.ldp) + H);
      }
      return (uint64_t)col < a.n_local
                 ? reinterpret_cast<const float4 *>(a.P + (uint64_t)col * a.ldp + H)
                 : reinterpret_cast<const float4 *>(a.halo + ((uint64_t)col - a.n_local) * (3 * H));
    };
    auto load = [&](RowRegs<P> &r, const float4 *rp) {
#pragma unroll
      for (int i = 0; i < P; ++i) r.k[i] = ldg4(rp + i * 32 + lane);
#pragma unroll
      for (int i = 0; i < P; ++i) r.v[i] = ldg4(rp + V + i * 32 + lane);
#pragma unroll
      for (int i = 0; i < P; ++i) r.m[i] = ldg4(rp + 2 * V + i * 32 + lane);
    };
    auto consume = [&](const RowRegs<P> &r, float what) {
      float part = 0.f;
#pragma unroll
      for (int i = 0; i < P; ++i) {
        part = fmaf(q[i].x, r.k[i].x, part);
        part = fmaf(q[i].y, r.k[i].y, part);
        part = fmaf(q[i].z, r.k[i].z, part);
        part = fmaf(q[i].w, r.k[i].w, part);
      }
#pragma unroll
      for (int o = LH / 2; o > 0; o >>= 1) part += __shfl_xor_sync(0xffffffffu, part, o);
      const float s = part * inv_scale;  // dot / sqrt(head_dim)  (:184-191)
      const float mnew = fmaxf(mx, s);
      const float corr = expf(mx - mnew);
      const float p = expf(s - mnew);
      l = fmaf(l, corr, p);
      mx = mnew;
#pragma unroll
      for (int i = 0; i < P; ++i) {
        acc[i].x = fmaf(acc[i].x, corr, p * r.v[i].x);
        acc[i].y = fmaf(acc[i].y, corr, p * r.v[i].y);
        acc[i].z = fmaf(acc[i].z, corr, p * r.v[i].z);
        acc[i].w = fmaf(acc[i].w, corr, p * r.v[i].w);
        ag[i].x = fmaf(what, r.m[i].x, ag[i].x);
        ag[i].y = fmaf(what, r.m[i].y, ag[i].y);
        ag[i].z = fmaf(what, r.m[i].z, ag[i].z);
        ag[i].w = fmaf(what, r.m[i].w, ag[i].w);
      }
    };

    for (uint64_t c = beg; c < end; c += 32) {
      const uint64_t e = c + lane;
      const uint32_t my_col = e < end ? __ldg(a.col_idx + e) : 0u;
      const float my_w = e < end ? rule.apply(__ldg(a.edge_w + e)) : 0.f;
      const int cnt = (end - c) < 32 ? (int)(end - c) : 32;
      int jj = 0;
      for (; jj + U <= cnt; jj += U) {
        RowRegs<P> r[U];
        float w[U];
#pragma unroll
        for (int u = 0; u < U; ++u) {
          const uint32_t col = __shfl_sync(0xffffffffu, my_col, jj + u);
          w[u] = __shfl_sync(0xffffffffu, my_w, jj + u);
          load(r[u], row_of(col));
        }
#pragma unroll
        for (int u = 0; u < U; ++u) consume(r[u], w[u]);
      }
      for (; jj < cnt; ++jj) {
        RowRegs<P> r;
        const uint32_t col = __shfl_sync(0xffffffffu, my_col, jj);
        const float w = __shfl_sync(0xffffffffu, my_w, jj);
        load(r, row_of(col));
        consume(r, w);
      }
    }
    const float inv_l = 1.0f / fmaxf(l, 1e-10f);  // sum_exp.max(1e-10) (:198)
#pragma unroll
    for (int i = 0; i < P; ++i) {
      orow[i * 32 + lane] =
          make_float4(acc[i].x * inv_l, acc[i].y * inv_l, acc[i].z * inv_l, acc[i].w * inv_l);
      orow[V + i * 32 + lane] = ag[i];
    }
  }
}


// ---- multi-pass variants (partitioned forward) --------------------------------------------
// The partitioned forward splits a node's neighbour list by where the rows live (own rows,
// then one column range per halo chunk as it arrives) and runs one gather pass per range while
// the next chunk is still on the wire.  The split is made ONCE per graph: build_pass_csr_kernel
// writes a second copy of the edge arrays in which every row's edges are stably grouped by pass
// (stored order kept inside a group), with the edge weights already normalised by the rule of
// aggregate_messages (layer.rs:426-432, the sum runs over ALL of the row's edges), and per row the
// group offsets seg[row][0..npass].  A pass then walks one contiguous group per row with the inner
// loops of the single-pass kernels -- no filtering, no compaction -- and carries the online-softmax
// state (raw accumulators, running max and running sum per head) to the next pass through oa / ml.
// Rows without edges in a middle pass are skipped outright.  The merge is the rescaling the
// single-pass kernels apply edge by edge, so the result equals the single pass up to f32 summation
// order (own tolerance, tests/test_gpu_multipass.py).
__global__ void __launch_bounds__(256) build_pass_csr_kernel(const uint64_t *row_ptr, const uint32_t *col,
                                                             const float *w, uint64_t n_rows,
                                                             const uint64_t *bounds, int npass,
                                                             int prenormalized, uint32_t *col_p,
                                                             float *w_p, uint32_t *seg) {
  const int lane = threadIdx.x & 31;
  const uint64_t warp = ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const uint64_t nwarps = ((uint64_t)gridDim.x * blockDim.x) >> 5;
  for (uint64_t row = warp; row < n_rows; row += nwarps) {
    const uint64_t beg = row_ptr[row], end = row_ptr[row + 1];
    uint32_t *sg = seg + row * (uint64_t)(npass + 1);
    if (beg == end) {
      for (int p = lane; p <= npass; p += 32) sg[p] = 0;
      continue;
    }
    const WeightRule rule = make_rule(w, beg, end, lane, prenormalized);
    uint32_t base = 0;
    for (int p = 0; p < npass; ++p) {
      if (lane == 0) sg[p] = base;
      const uint64_t lo = bounds[p], hi = bounds[p + 1];
      for (uint64_t c = beg; c < end; c += 32) {
        const uint64_t e = c + lane;
        const uint32_t cc = e < end ? col[e] : 0u;
        const bool in = e < end && (uint64_t)cc >= lo && (uint64_t)cc < hi;
        const uint32_t mask = __ballot_sync(0xffffffffu, in);
        if (in) {
          const uint64_t pos = beg + base + __popc(mask & ((1u << lane) - 1u));
          col_p[pos] = cc;
          w_p[pos] = rule.apply(w[e]);
        }
        base += __popc(mask);
      }
    }
    if (lane == 0) sg[npass] = base;  // == degree when the bounds cover every column
  }
}

template <int P, int LH, int U>
__global__ void __launch_bounds__(128) gather_fast_mp_kernel(GatherArgs a, float inv_scale) {
  constexpr int H = 128 * P;
  constexpr int V = 32 * P;  // float4 per H-row
  const int lane = threadIdx.x & 31;
  const uint64_t warp = ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const uint64_t nwarps = ((uint64_t)gridDim.x * blockDim.x) >> 5;
  const int heads = 32 / LH, head = lane / LH;

  for (uint64_t node = a.row_begin + warp; node < a.row_end; node += nwarps) {
    const uint64_t beg0 = __ldg(a.row_ptr + node), end0 = __ldg(a.row_ptr + node + 1);
    float4 *orow = reinterpret_cast<float4 *>(a.oa + node * (uint64_t)(2 * H));
    float *ml = a.ml + node * (uint64_t)(2 * heads);
    if (beg0 == end0) {  // no neighbours: zeros (the update takes the LN(m) path)
      if (a.first || a.last) {
        const float4 zero = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
        for (int i = 0; i < 2 * P; ++i) orow[i * 32 + lane] = zero;
      }
      continue;
    }
    const uint32_t *sg = a.seg + node * (uint64_t)a.npass1;
    const uint64_t beg = beg0 + __ldg(sg + a.pass), end = beg0 + __ldg(sg + a.pass + 1);
    if (beg == end && !a.first && !a.last) continue;  // nothing to merge for this row in this pass
    const float4 *qrow = reinterpret_cast<const float4 *>(a.P + node * (uint64_t)a.ldp);
    float4 q[P], acc[P], ag[P];
    float mx = -INFINITY, l = 0.f;
#pragma unroll
    for (int i = 0; i < P; ++i) q[i] = ldg4(qrow + i * 32 + lane);
    if (a.first) {
#pragma unroll
      for (int i = 0; i < P; ++i) {
        acc[i] = make_float4(0.f, 0.f, 0.f, 0.f);
        ag[i] = make_float4(0.f, 0.f, 0.f, 0.f);
      }
    } else {
#pragma unroll
      for (int i = 0; i < P; ++i) {
        acc[i] = orow[i * 32 + lane];
        ag[i] = orow[V + i * 32 + lane];
      }
      mx = ml[head];
      l = ml[heads + head];
    }

    auto row_of = [&](uint32_t col) -> const float4 * {
      return (uint64_t)col < a.n_local
                 ? reinterpret_cast<const float4 *>(a.P + (uint64_t)col * a.ldp + H)
                 : reinterpret_cast<const float4 *>(a.halo + ((uint64_t)col - a.n_local) * (3 * H));
    };
    auto load = [&](RowRegs<P> &r, const float4 *rp) {
#pragma unroll
      for (int i = 0; i < P; ++i) r.k[i] = ldg4(rp + i * 32 + lane);
#pragma unroll
      for (int i = 0; i < P; ++i) r.v[i] = ldg4(rp + V + i * 32 + lane);
#pragma unroll
      for (int i = 0; i < P; ++i) r.m[i] = ldg4(rp + 2 * V + i * 32 + lane);
    };
    auto consume = [&](const RowRegs<P> &r, float what) {
      float part = 0.f;
#pragma unroll
      for (int i = 0; i < P; ++i) {
        part = fmaf(q[i].x, r.k[i].x, part);
        part = fmaf(q[i].y, r.k[i].y, part);
        part = fmaf(q[i].z, r.k[i].z, part);
        part = fmaf(q[i].w, r.k[i].w, part);
      }
#pragma unroll
      for (int o = LH / 2; o > 0; o >>= 1) part += __shfl_xor_sync(0xffffffffu, part, o);
      const float s = part * inv_scale;
      const float mnew = fmaxf(mx, s);
      const float corr = expf(mx - mnew);
      const float p = expf(s - mnew);
      l = fmaf(l, corr, p);
      mx = mnew;
#pragma unroll
      for (int i = 0; i < P; ++i) {
        acc[i].x = fmaf(acc[i].x, corr, p * r.v[i].x);
        acc[i].y = fmaf(acc[i].y, corr, p * r.v[i].y);
        acc[i].z = fmaf(acc[i].z, corr, p * r.v[i].z);
        acc[i].w = fmaf(acc[i].w, corr, p * r.v[i].w);
        ag[i].x = fmaf(what, r.m[i].x, ag[i].x);
        ag[i].y = fmaf(what, r.m[i].y, ag[i].y);
        ag[i].z = fmaf(what, r.m[i].z, ag[i].z);
        ag[i].w = fmaf(what, r.m[i].w, ag[i].w);
      }
    };

    for (uint64_t c = beg; c < end; c += 32) {
      const uint64_t e = c + lane;
      const uint32_t my_col = e < end ? __ldg(a.col_idx + e) : 0u;
      const float my_w = e < end ? __ldg(a.edge_w + e) : 0.f;  // normalised when the pass CSR was built
      const int cnt = (end - c) < 32 ? (int)(end - c) : 32;
      int jj = 0;
      for (; jj + U <= cnt; jj += U) {
        RowRegs<P> r[U];
        float w[U];
#pragma unroll
        for (int u = 0; u < U; ++u) {
          const uint32_t col = __shfl_sync(0xffffffffu, my_col, jj + u);
          w[u] = __shfl_sync(0xffffffffu, my_w, jj + u);
          load(r[u], row_of(col));
        }
#pragma unroll
        for (int u = 0; u < U; ++u) consume(r[u], w[u]);
      }
      for (; jj < cnt; ++jj) {
        RowRegs<P> r;
        const uint32_t col = __shfl_sync(0xffffffffu, my_col, jj);
        const float w = __shfl_sync(0xffffffffu, my_w, jj);
        load(r, row_of(col));
        consume(r, w);
      }
    }
    if (a.last) {
      const float inv_l = 1.0f / fmaxf(l, 1e-10f);  // sum_exp.max(1e-10) (:198)
#pragma unroll
      for (int i = 0; i < P; ++i) {
        orow[i * 32 + lane] =
            make_float4(acc[i].x * inv_l, acc[i].y * inv_l, acc[i].z * inv_l, acc[i].w * inv_l);
        orow[V + i * 32 + lane] = ag[i];
      }
    } else {
#pragma unroll
      for (int i = 0; i < P; ++i) {
        orow[i * 32 + lane] = acc[i];
        orow[V + i * 32 + lane] = ag[i];
      }
      if (lane % LH == 0) {
        ml[head] = mx;
        ml[heads + head] = l;
      }
    }
  }
}

// Any H / head count (identity column order).  One warp per node, running
// state in shared memory.  Correctness path for shapes the fast kernel does
// not cover (e.g. the reference's 4->8, 2-head unit-test layer).
__global__ void gather_generic_kernel(GatherArgs a, float inv_scale) {
  extern __shared__ float sm[];
  const int H = (int)a.H, h = (int)a.heads, dh = H / h;
  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
  float *q = sm + (size_t)wib * (3 * H + 2 * h);
  float *acc = q + H, *ag = acc + H, *mxs = ag + H, *ls = mxs + h;
  const uint64_t warp = ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const uint64_t nwarps = ((uint64_t)gridDim.x * blockDim.x) >> 5;
  for (uint64_t node = a.row_begin + warp; node < a.row_end; node += nwarps) {
    const uint64_t beg = a.row_ptr[node], end = a.row_ptr[node + 1];
    float *orow = a.oa + node * (uint64_t)(2 * H);
    if (beg == end) {
      for (int t = lane; t < 2 * H; t += 32) orow[t] = 0.f;
      continue;
    }
    const float *qrow = a.P + node * (uint64_t)a.ldp;
    for (int t = lane; t < H; t += 32) {
      q[t] = qrow[t];
      acc[t] = 0.f;
      ag[t] = 0.f;
    }
    for (int g = lane; g < h; g += 32) {
      mxs[g] = -INFINITY;
      ls[g] = 0.f;
    }
    const WeightRule rule = make_rule(a.edge_w, beg, end, lane, a.prenormalized);
    __syncwarp();
    for (uint64_t e = beg; e < end; ++e) {
      const uint64_t col = a.col_idx[e];
      const float what = rule.apply(a.edge_w[e]);
      const float *rp = col < a.n_local ? a.P + col * (uint64_t)a.ldp + H
                                        : a.halo + (col - a.n_local) * (uint64_t)(3 * H);
      const float *kr = rp, *vr = rp + H, *mr = rp + 2 * H;
      for (int g = 0; g < h; ++g) {
        float part = 0.f;
        for (int t = lane; t < dh; t += 32) part = fmaf(q[g * dh + t], kr[g * dh + t], part);
        part = warp_sum(part);
        const float s = part * inv_scale;
        const float mold = mxs[g], lold = ls[g];
        const float mnew = fmaxf(mold, s);
        const float corr = expf(mold - mnew), p = expf(s - mnew);
        for (int t = lane; t < dh; t += 32)
          acc[g * dh + t] = fmaf(acc[g * dh + t], corr, p * vr[g * dh + t]);
        __syncwarp();
        if (lane == 0) {
          mxs[g] = mnew;
          ls[g] = fmaf(lold, corr, p);
        }
      }
      for (int t = lane; t < H; t += 32) ag[t] = fmaf(what, mr[t], ag[t]);
      __syncwarp();
    }
    for (int t = lane; t < H; t += 32) {
      orow[t] = acc[t] * (1.0f / fmaxf(ls[t / dh], 1e-10f));
      orow[H + t] = ag[t];
    }
    __syncwarp();
  }
}


// ---- R = 1 kernel (H = 128) --------------------------------------------------------------
__device__ __forceinline__ float ex2_approx(float x) {  // 2^x, x <= 0 here; flushes to 0 below 2^-126
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

// Sum each of the N per-lane partials over the 32 lanes and leave ALL N totals in every lane.
// Halving butterfly: each xor step exchanges half of the remaining values, so N totals cost
// (N-1) + (5 - log2 N) + N shuffles instead of 5N.
template <int N>
__device__ __forceinline__ void reduce_all(float (&v)[N], int lane) {
  int off = 16;
#pragma unroll
  for (int n = N; n > 1; n >>= 1) {
    const bool up = (lane & off) != 0;
#pragma unroll
    for (int i = 0; i < n / 2; ++i) {
      const float keep = up ? v[i + n / 2] : v[i];
      const float send = up ? v[i] : v[i + n / 2];
      v[i] = keep + __shfl_xor_sync(0xffffffffu, send, off);
    }
    off >>= 1;
  }
#pragma unroll
  for (; off > 0; off >>= 1) v[0] += __shfl_xor_sync(0xffffffffu, v[0], off);
  // lane L now holds the total of value index id(L) = bits of L at the halving offsets
  const float mine = v[0];
#pragma unroll
  for (int g = 0; g < N; ++g) {
    int src = 0, o = 16;
#pragma unroll
    for (int n = N; n > 1; n >>= 1) {
      if (g & (n / 2)) src |= o;
      o >>= 1;
    }
    v[g] = __shfl_sync(0xffffffffu, mine, src);
  }
}

template <int HEADS, int U>
__global__ void __launch_bounds__(128) gather_r1_kernel(GatherR1Args a, float scale_log2) {
  constexpr int H = 128;
  const int lane = threadIdx.x & 31;
  const uint64_t warp = ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const uint64_t nwarps = ((uint64_t)gridDim.x * blockDim.x) >> 5;
  for (uint64_t node = a.row_begin + warp; node < a.row_end; node += nwarps) {
    const uint64_t beg = __ldg(a.row_ptr + node), end = __ldg(a.row_ptr + node + 1);
    float4 *srow = reinterpret_cast<float4 *>(a.sa + node * (uint64_t)((HEADS + 1) * H));
    if (beg == end) {
      const float4 zero = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
      for (int g = 0; g <= HEADS; ++g) srow[g * 32 + lane] = zero;
      continue;
    }
    const float4 *qrow = reinterpret_cast<const float4 *>(a.Qt + node * (uint64_t)(HEADS * H));
    float4 qt[HEADS], S[HEADS];
    float mx[HEADS], l[HEADS];
#pragma unroll
    for (int g = 0; g < HEADS; ++g) {
      const float4 q = ldg4(qrow + g * 32 + lane);
      // scores are taken in log2 units: exp(s - m) == exp2(s' - m') with s' = s * log2(e)
      qt[g] = make_float4(q.x * scale_log2, q.y * scale_log2, q.z * scale_log2, q.w * scale_log2);
      S[g] = make_float4(0.f, 0.f, 0.f, 0.f);
      mx[g] = -INFINITY;
      l[g] = 0.f;
    }
    float4 ag = make_float4(0.f, 0.f, 0.f, 0.f);
    const WeightRule rule = make_rule(a.edge_w, beg, end, lane, a.prenormalized);
    auto row_of = [&](uint32_t col) -> const float4 * {
      return (uint64_t)col < a.n_local
                 ? reinterpret_cast<const float4 *>(a.M + (uint64_t)col * a.ldm)
                 : reinterpret_cast<const float4 *>(a.halo + ((uint64_t)col - a.n_local) * H);
    };
    auto consume = [&](const float4 m, float what) {
      float s[HEADS];
#pragma unroll
      for (int g = 0; g < HEADS; ++g)
        s[g] = fmaf(qt[g].x, m.x, fmaf(qt[g].y, m.y, fmaf(qt[g].z, m.z, qt[g].w * m.w)));
      reduce_all<HEADS>(s, lane);
#pragma unroll
      for (int g = 0; g < HEADS; ++g) {
        // Branch-free online softmax with ONE exp2 per head: with d = s - mx, either the new
        // score is the maximum (p = 1, corr = 2^-d) or the old one stays (corr = 1, p = 2^d).
        const float d = s[g] - mx[g];
        const float e = ex2_approx(-fabsf(d));  // mx = -inf on the first edge: d = +inf, e = 0
        const bool grow = d > 0.f;
        const float corr = grow ? e : 1.0f;
        const float p = grow ? 1.0f : e;
        mx[g] = fmaxf(mx[g], s[g]);
        l[g] = fmaf(l[g], corr, p);
        S[g].x = fmaf(S[g].x, corr, p * m.x);
        S[g].y = fmaf(S[g].y, corr, p * m.y);
        S[g].z = fmaf(S[g].z, corr, p * m.z);
        S[g].w = fmaf(S[g].w, corr, p * m.w);
      }
      ag.x = fmaf(what, m.x, ag.x);
      ag.y = fmaf(what, m.y, ag.y);
      ag.z = fmaf(what, m.z, ag.z);
      ag.w = fmaf(what, m.w, ag.w);
    };
    for (uint64_t c = beg; c < end; c += 32) {
      const uint64_t e = c + lane;
      const uint32_t my_col = e < end ? __ldg(a.col_idx + e) : 0u;
      const float my_w = e < end ? rule.apply(__ldg(a.edge_w + e)) : 0.f;
      const int cnt = (end - c) < 32 ? (int)(end - c) : 32;
      int jj = 0;
      for (; jj + U <= cnt; jj += U) {
        float4 r[U];
        float w[U];
#pragma unroll
        for (int u = 0; u < U; ++u) {
          const uint32_t col = __shfl_sync(0xffffffffu, my_col, jj + u);
          w[u] = __shfl_sync(0xffffffffu, my_w, jj + u);
          r[u] = ldg4(row_of(col) + lane);
        }
#pragma unroll
        for (int u = 0; u < U; ++u) consume(r[u], w[u]);
      }
      for (; jj < cnt; ++jj) {
        const uint32_t col = __shfl_sync(0xffffffffu, my_col, jj);
        const float w = __shfl_sync(0xffffffffu, my_w, jj);
        consume(ldg4(row_of(col) + lane), w);
      }
    }
#pragma unroll
    for (int g = 0; g < HEADS; ++g) {
      const float inv_l = 1.0f / fmaxf(l[g], 1e-10f);
      srow[g * 32 + lane] = make_float4(S[g].x * inv_l, S[g].y * inv_l, S[g].z * inv_l, S[g].w * inv_l);
    }
    srow[HEADS * 32 + lane] = ag;
  }
}

// Sum each of the N per-lane partials over the 32 lanes; lane L returns the total of value
// index L / (32 / N) (halving butterfly as above, without the final broadcast).
template <int N>
__device__ __forceinline__ float reduce_scatter(float (&v)[N], int lane) {
  int off = 16;
#pragma unroll
  for (int n = N; n > 1; n >>= 1) {
    const bool up = (lane & off) != 0;
#pragma unroll
    for (int i = 0; i < n / 2; ++i) {
      const float keep = up ? v[i + n / 2] : v[i];
      const float send = up ? v[i] : v[i + n / 2];
      v[i] = keep + __shfl_xor_sync(0xffffffffu, send, off);
    }
    off >>= 1;
  }
#pragma unroll
  for (; off > 0; off >>= 1) v[0] += __shfl_xor_sync(0xffffffffu, v[0], off);
  return v[0];
}

// Halving butterfly without selects: the caller stores the partial of value index j at
// position j ^ id, id = lane / (32 / N).  Every lane then keeps v[i] and sends v[i + n/2]; lane
// L returns the total of value index id(L).
template <int N>
__device__ __forceinline__ float reduce_scatter_xor(float (&v)[N]) {
  int off = 16;
#pragma unroll
  for (int n = N; n > 1; n >>= 1) {
#pragma unroll
    for (int i = 0; i < n / 2; ++i) v[i] += __shfl_xor_sync(0xffffffffu, v[i + n / 2], off);
    off >>= 1;
  }
#pragma unroll
  for (; off > 0; off >>= 1) v[0] += __shfl_xor_sync(0xffffffffu, v[0], off);
  return v[0];
}

// R = 1, batched: 8 edges per step.  The 8 x HEADS scores of a step are reduced together (one
// halving butterfly leaves ONE (edge, head) total per lane: edge slot = lane / 4, head =
// (lane % 4) / (4 / HEADS)), so the softmax bookkeeping -- running maximum, exp2, normaliser --
// is one instruction per step instead of one per edge and head; the weights are then broadcast
// back for the accumulation.
// Lane-permuted slots: lane L keeps in slot (u, g) the row of edge u ^ (L / 4) and the query /
// accumulator of head g ^ head(L).  That is the layout the select-free butterfly wants, and every
// broadcast becomes an xor-shuffle with a compile-time lane mask; the rows are still read as
// 128-bit loads of the lane's own 4 columns.
// Register budget: 16 resident warps per SM (128 registers).  Measured at 4M nodes: cutting it for 20 / 24
// warps (96 / 80 registers, small spills) is slower (24.8 / 28.3 vs 23.8 ms), and so is requesting the
// next step's rows one step ahead (167 registers, 12 warps: 26.6 ms).
template <int HEADS, int HALO, bool PREFETCH>  // HALO: 0 own rows only, 1 halo buffer, 2 peer-direct
__global__ void __launch_bounds__(32, 16) gather_r1_batched_kernel(GatherR1Args a, float scale_log2) {
  constexpr int H = 128, EB = 8, N = EB * HEADS, REP = 32 / N;
  __shared__ PeerTable s_peer;
  if (HALO == 2) {
    if (threadIdx.x == 0) s_peer = *a.peer;
    __syncwarp();
  }
  const uint64_t own_lo = HALO == 2 ? s_peer.bounds[s_peer.rank] : 0;
  // the message row of column `col`
  auto mrow = [&](uint32_t col) -> const float * {
    if (HALO == 2) {
      const uint64_t l = (uint64_t)col - own_lo;
      return l < a.n_local ? a.M + l * a.ldm : peer_row(s_peer, col, a.ldm);
    }
    if (HALO == 1 && (uint64_t)col >= a.n_local) return a.halo + ((uint64_t)col - a.n_local) * H;
    return a.M + (uint64_t)col * a.ldm;
  };
  // ONE warp per block: the node, its row bounds and every loop trip count then derive from
  // blockIdx, which the compiler knows to be warp-uniform -- the shuffles compile to bare SHFL
  // (with 4 warps per block each one sat in a WARPSYNC / ENDCOLLECTIVE region).
  const int lane = threadIdx.x;
  const uint64_t warp = blockIdx.x;
  const uint64_t nwarps = gridDim.x;
  const int my_slot = lane >> 2;               // edge slot of the (edge, head) pair this lane owns
  const int my_head = (lane / REP) % HEADS;    // and its head
  for (uint64_t node = a.row_begin + warp; node < a.row_end; node += nwarps) {
    const uint64_t beg = __ldg(a.row_ptr + node), end = __ldg(a.row_ptr + node + 1);
    float4 *srow = reinterpret_cast<float4 *>(a.sa + node * (uint64_t)((HEADS + 1) * H));
    if (beg == end) {
      const float4 zero = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
      for (int g = 0; g <= HEADS; ++g) srow[g * 32 + lane] = zero;
      continue;
    }
    const float4 *qrow = reinterpret_cast<const float4 *>(a.Qt + node * (uint64_t)(HEADS * H));
    float4 qt[HEADS], S[HEADS];  // slot g: head g ^ my_head
#pragma unroll
    for (int g = 0; g < HEADS; ++g) {
      const float4 q = ldg4(qrow + (g ^ my_head) * 32 + lane);
      qt[g] = make_float4(q.x * scale_log2, q.y * scale_log2, q.z * scale_log2, q.w * scale_log2);
      S[g] = make_float4(0.f, 0.f, 0.f, 0.f);
    }
    float4 ag = make_float4(0.f, 0.f, 0.f, 0.f);
    float mx = -INFINITY, lsum = 0.f;  // of my head / my (edge slot, head) pair
    const WeightRule rule = make_rule(a.edge_w, beg, end, lane, a.prenormalized);
    for (uint64_t c = beg; c < end; c += 32) {
      const uint64_t e = c + lane;
      const uint32_t my_col = e < end ? __ldg(a.col_idx + e) : 0u;
      const float my_w = e < end ? rule.apply(__ldg(a.edge_w + e)) : 0.f;
      const int cnt = (end - c) < 32 ? (int)(end - c) : 32;
      for (int jj = 0; jj < cnt; jj += EB) {
        float4 r[EB];  // slot u: edge jj + (u ^ my_slot)
        if (PREFETCH && jj + EB < cnt) {
          // one instruction pulls the NEXT step's 8 rows (32 lines of 128 B) towards L1: lane L takes
          // line L % 4 of the row of edge jj + 8 + L / 4, so those loads find their data on chip
          const int nsrc = jj + EB + my_slot;
          const uint32_t ncol = __shfl_sync(0xffffffffu, my_col, nsrc < cnt ? nsrc : cnt - 1);
          // own rows only: an L1 prefetch of a peer address is pathologically slow (measured: 30M remote
          // message rows took 1.9 s with it)
          if (HALO != 2 || (uint64_t)ncol - own_lo < a.n_local) {
            const float *nrow = mrow(ncol);
            asm volatile("prefetch.global.L1 [%0];" ::"l"(nrow + (lane & 3) * 32));
          }
        }
#pragma unroll
        for (int u = 0; u < EB; ++u) {
          // slots beyond the row re-read its last edge (score masked below, weight 0)
          const int src = jj + (u ^ my_slot);
          const uint32_t col = __shfl_sync(0xffffffffu, my_col, src < cnt ? src : cnt - 1);
          r[u] = ldg4(reinterpret_cast<const float4 *>(mrow(col)) + lane);
        }
        float s[N];  // position u * HEADS + g holds the partial of edge slot u ^ my_slot, head g ^ my_head
#pragma unroll
        for (int u = 0; u < EB; ++u)
#pragma unroll
          for (int g = 0; g < HEADS; ++g)
            s[u * HEADS + g] = fmaf(qt[g].x, r[u].x, fmaf(qt[g].y, r[u].y, fmaf(qt[g].z, r[u].z, qt[g].w * r[u].w)));
        float t = reduce_scatter_xor<N>(s);  // log2-unit score of (edge jj + my_slot, my_head)
        if (jj + my_slot >= cnt) t = -INFINITY;
        float bm = t;  // maximum of the step per head: lanes of one head differ in bits 2..4
        bm = fmaxf(bm, __shfl_xor_sync(0xffffffffu, bm, 4));
        bm = fmaxf(bm, __shfl_xor_sync(0xffffffffu, bm, 8));
        bm = fmaxf(bm, __shfl_xor_sync(0xffffffffu, bm, 16));
        const float nm = fmaxf(mx, bm);  // finite: every step holds at least one edge
        const float p = ex2_approx(t - nm);
        const float corr = ex2_approx(mx - nm);  // first step: mx = -inf, corr = 0 (S, lsum are 0)
        mx = nm;
        lsum = fmaf(lsum, corr, p);
#pragma unroll
        for (int g = 0; g < HEADS; ++g) {
          const float cg = g ? __shfl_xor_sync(0xffffffffu, corr, g * REP) : corr;
          S[g].x *= cg; S[g].y *= cg; S[g].z *= cg; S[g].w *= cg;
        }
#pragma unroll
        for (int u = 0; u < EB; ++u) {
#pragma unroll
          for (int g = 0; g < HEADS; ++g) {
            const float pug = (u | g) ? __shfl_xor_sync(0xffffffffu, p, (u * HEADS + g) * REP) : p;
            S[g].x = fmaf(pug, r[u].x, S[g].x);
            S[g].y = fmaf(pug, r[u].y, S[g].y);
            S[g].z = fmaf(pug, r[u].z, S[g].z);
            S[g].w = fmaf(pug, r[u].w, S[g].w);
          }
          const float w = __shfl_sync(0xffffffffu, my_w, jj + (u ^ my_slot));  // 0 beyond the row
          ag.x = fmaf(w, r[u].x, ag.x);
          ag.y = fmaf(w, r[u].y, ag.y);
          ag.z = fmaf(w, r[u].z, ag.z);
          ag.w = fmaf(w, r[u].w, ag.w);
        }
      }
    }
    lsum += __shfl_xor_sync(0xffffffffu, lsum, 4);  // normaliser of my head: sum over the edge slots
    lsum += __shfl_xor_sync(0xffffffffu, lsum, 8);
    lsum += __shfl_xor_sync(0xffffffffu, lsum, 16);
#pragma unroll
    for (int g = 0; g < HEADS; ++g) {
      const float lg = g ? __shfl_xor_sync(0xffffffffu, lsum, g * REP) : lsum;
      const float inv_l = 1.0f / fmaxf(lg, 1e-10f);
      srow[(g ^ my_head) * 32 + lane] = make_float4(S[g].x * inv_l, S[g].y * inv_l, S[g].z * inv_l, S[g].w * inv_l);
    }
    srow[HEADS * 32 + lane] = ag;
  }
}


// Multi-pass form of gather_r1_batched_kernel (see gather_fast_mp_kernel): the state is the raw
// S_g | agg rows in sa and (running max, running sum) per head in ml.
template <int HEADS>
__global__ void __launch_bounds__(32, 16) gather_r1_batched_mp_kernel(GatherR1Args a, float scale_log2) {
  constexpr int H = 128, EB = 8, N = EB * HEADS, REP = 32 / N;
  const int lane = threadIdx.x;
  const uint64_t warp = blockIdx.x;
  const uint64_t nwarps = gridDim.x;
  const int my_slot = lane >> 2;
  const int my_head = (lane / REP) % HEADS;
  for (uint64_t node = a.row_begin + warp; node < a.row_end; node += nwarps) {
    const uint64_t beg0 = __ldg(a.row_ptr + node), end0 = __ldg(a.row_ptr + node + 1);
    float4 *srow = reinterpret_cast<float4 *>(a.sa + node * (uint64_t)((HEADS + 1) * H));
    float *ml = a.ml + node * (uint64_t)(2 * HEADS);
    if (beg0 == end0) {
      if (a.first || a.last) {
        const float4 zero = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
        for (int g = 0; g <= HEADS; ++g) srow[g * 32 + lane] = zero;
      }
      continue;
    }
    const uint32_t *sg = a.seg + node * (uint64_t)a.npass1;
    const uint64_t beg = beg0 + __ldg(sg + a.pass), end = beg0 + __ldg(sg + a.pass + 1);
    if (beg == end && !a.first && !a.last) continue;  // nothing to merge for this row in this pass
    const float4 *qrow = reinterpret_cast<const float4 *>(a.Qt + node * (uint64_t)(HEADS * H));
    float4 qt[HEADS], S[HEADS];  // slot g: head g ^ my_head
    float4 ag = make_float4(0.f, 0.f, 0.f, 0.f);
    float mx = -INFINITY, lsum = 0.f;
#pragma unroll
    for (int g = 0; g < HEADS; ++g) {
      const float4 q = ldg4(qrow + (g ^ my_head) * 32 + lane);
      qt[g] = make_float4(q.x * scale_log2, q.y * scale_log2, q.z * scale_log2, q.w * scale_log2);
      S[g] = a.first ? make_float4(0.f, 0.f, 0.f, 0.f) : srow[(g ^ my_head) * 32 + lane];
    }
    if (!a.first) {
      ag = srow[HEADS * 32 + lane];
      mx = ml[my_head];
      lsum = (my_slot == 0) ? ml[HEADS + my_head] : 0.f;  // the sum is re-split over the edge slots
    }
    for (uint64_t c = beg; c < end; c += 32) {
      const uint64_t e = c + lane;
      const uint32_t my_col = e < end ? __ldg(a.col_idx + e) : 0u;
      const float my_w = e < end ? __ldg(a.edge_w + e) : 0.f;  // normalised when the pass CSR was built
      const int cnt = (end - c) < 32 ? (int)(end - c) : 32;
      for (int jj = 0; jj < cnt; jj += EB) {
        float4 r[EB];  // slot u: edge jj + (u ^ my_slot)
        if (jj + EB < cnt) {
          const int nsrc = jj + EB + my_slot;
          const uint32_t ncol = __shfl_sync(0xffffffffu, my_col, nsrc < cnt ? nsrc : cnt - 1);
          const float *nrow = ((uint64_t)ncol >= a.n_local)
                                  ? a.halo + ((uint64_t)ncol - a.n_local) * H
                                  : a.M + (uint64_t)ncol * a.ldm;
          asm volatile("prefetch.global.L1 [%0];" ::"l"(nrow + (lane & 3) * 32));
        }
#pragma unroll
        for (int u = 0; u < EB; ++u) {
          const int src = jj + (u ^ my_slot);
          const uint32_t col = __shfl_sync(0xffffffffu, my_col, src < cnt ? src : cnt - 1);
          const float4 *row =
              (uint64_t)col < a.n_local
                  ? reinterpret_cast<const float4 *>(a.M + (uint64_t)col * a.ldm)
                  : reinterpret_cast<const float4 *>(a.halo + ((uint64_t)col - a.n_local) * H);
          r[u] = ldg4(row + lane);
        }
        float s[N];
#pragma unroll
        for (int u = 0; u < EB; ++u)
#pragma unroll
          for (int g = 0; g < HEADS; ++g)
            s[u * HEADS + g] = fmaf(qt[g].x, r[u].x, fmaf(qt[g].y, r[u].y, fmaf(qt[g].z, r[u].z, qt[g].w * r[u].w)));
        float t = reduce_scatter_xor<N>(s);
        if (jj + my_slot >= cnt) t = -INFINITY;
        float bm = t;
        bm = fmaxf(bm, __shfl_xor_sync(0xffffffffu, bm, 4));
        bm = fmaxf(bm, __shfl_xor_sync(0xffffffffu, bm, 8));
        bm = fmaxf(bm, __shfl_xor_sync(0xffffffffu, bm, 16));
        const float nm = fmaxf(mx, bm);  // finite: the step holds at least one edge
        const float p = ex2_approx(t - nm);
        const float corr = ex2_approx(mx - nm);
        mx = nm;
        lsum = fmaf(lsum, corr, p);
#pragma unroll
        for (int g = 0; g < HEADS; ++g) {
          const float cg = g ? __shfl_xor_sync(0xffffffffu, corr, g * REP) : corr;
          S[g].x *= cg; S[g].y *= cg; S[g].z *= cg; S[g].w *= cg;
        }
#pragma unroll
        for (int u = 0; u < EB; ++u) {
#pragma unroll
          for (int g = 0; g < HEADS; ++g) {
            const float pug = (u | g) ? __shfl_xor_sync(0xffffffffu, p, (u * HEADS + g) * REP) : p;
            S[g].x = fmaf(pug, r[u].x, S[g].x);
            S[g].y = fmaf(pug, r[u].y, S[g].y);
            S[g].z = fmaf(pug, r[u].z, S[g].z);
            S[g].w = fmaf(pug, r[u].w, S[g].w);
          }
          const int wsrc = jj + (u ^ my_slot);
          const float w = __shfl_sync(0xffffffffu, my_w, wsrc & 31);  // 0 beyond the compacted list
          ag.x = fmaf(w, r[u].x, ag.x);
          ag.y = fmaf(w, r[u].y, ag.y);
          ag.z = fmaf(w, r[u].z, ag.z);
          ag.w = fmaf(w, r[u].w, ag.w);
        }
      }
    }
    lsum += __shfl_xor_sync(0xffffffffu, lsum, 4);  // normaliser of my head: sum over the edge slots
    lsum += __shfl_xor_sync(0xffffffffu, lsum, 8);
    lsum += __shfl_xor_sync(0xffffffffu, lsum, 16);
    if (a.last) {
#pragma unroll
      for (int g = 0; g < HEADS; ++g) {
        const float lg = g ? __shfl_xor_sync(0xffffffffu, lsum, g * REP) : lsum;
        const float inv_l = 1.0f / fmaxf(lg, 1e-10f);
        srow[(g ^ my_head) * 32 + lane] = make_float4(S[g].x * inv_l, S[g].y * inv_l, S[g].z * inv_l, S[g].w * inv_l);
      }
    } else {
#pragma unroll
      for (int g = 0; g < HEADS; ++g) srow[(g ^ my_head) * 32 + lane] = S[g];
      if (my_slot == 0 && lane % REP == 0) {
        ml[my_head] = mx;
        ml[HEADS + my_head] = lsum;
      }
    }
    srow[HEADS * 32 + lane] = ag;
  }
}

template <int HEADS>
cudaError_t launch_r1_h(const GatherR1Args &a, float scale_log2, int sm_count, cudaStream_t s) {
  const uint64_t rows = a.row_end - a.row_begin;
  const int wpb = 4;
  uint64_t blocks = (rows + wpb - 1) / wpb;
  const uint64_t cap = (uint64_t)sm_count * 16;
  if (blocks > cap) blocks = cap;
  // RUVGNN_R1_BATCH=0 keeps the edge-at-a-time kernel (A/B runs)
  static int batched = -1;
  if (batched < 0) {
    const char *e = getenv("RUVGNN_R1_BATCH");
    batched = (e && atoi(e) == 0) ? 0 : 1;
  }
  if (a.multipass) {
    if constexpr (HEADS <= 4) {
      uint64_t warps = rows < (uint64_t)sm_count * 64 ? rows : (uint64_t)sm_count * 64;
      gather_r1_batched_mp_kernel<HEADS><<<(unsigned)warps, 32, 0, s>>>(a, scale_log2);
      return cudaGetLastError();
    } else {
      return cudaErrorNotSupported;
    }
  }
  if constexpr (HEADS <= 4) {
    if (batched) {
      uint64_t warps = rows < (uint64_t)sm_count * 64 ? rows : (uint64_t)sm_count * 64;
      static int pf = -1;  // RUVGNN_R1_PF=0: no L1 prefetch of the next step's rows (A/B runs)
      if (pf < 0) {
        const char *e = getenv("RUVGNN_R1_PF");
        pf = (e && atoi(e) == 0) ? 0 : 1;
      }
      if (a.peer)
        gather_r1_batched_kernel<HEADS, 2, true><<<(unsigned)warps, 32, 0, s>>>(a, scale_log2);
      else if (a.halo)
        gather_r1_batched_kernel<HEADS, 1, true><<<(unsigned)warps, 32, 0, s>>>(a, scale_log2);
      else if (pf)
        gather_r1_batched_kernel<HEADS, 0, true><<<(unsigned)warps, 32, 0, s>>>(a, scale_log2);
      else
        gather_r1_batched_kernel<HEADS, 0, false><<<(unsigned)warps, 32, 0, s>>>(a, scale_log2);
      return cudaGetLastError();
    }
  }
  if (a.peer) return cudaErrorNotSupported;
  gather_r1_kernel<HEADS, (HEADS <= 4 ? 8 : 4)><<<(unsigned)blocks, 32 * wpb, 0, s>>>(a, scale_log2);
  return cudaGetLastError();
}

template <int P, int U>
cudaError_t launch_fast_p(const GatherArgs &a, float inv_scale, int sm_count, cudaStream_t s) {
  const uint64_t rows = a.row_end - a.row_begin;
  const int wpb = 4;
  uint64_t blocks = (rows + wpb - 1) / wpb;
  // grid-stride: a few resident waves, a multiple of the SM count
  const uint64_t cap = (uint64_t)sm_count * 16;
  if (blocks > cap) blocks = cap;
  if (blocks == 0) return cudaSuccess;
  dim3 grid((unsigned)blocks), block(32 * wpb);
  if (a.multipass) {
    switch (32 / a.heads) {
      case 32: gather_fast_mp_kernel<P, 32, U><<<grid, block, 0, s>>>(a, inv_scale); break;
      case 16: gather_fast_mp_kernel<P, 16, U><<<grid, block, 0, s>>>(a, inv_scale); break;
      case 8: gather_fast_mp_kernel<P, 8, U><<<grid, block, 0, s>>>(a, inv_scale); break;
      case 4: gather_fast_mp_kernel<P, 4, U><<<grid, block, 0, s>>>(a, inv_scale); break;
      case 2: gather_fast_mp_kernel<P, 2, U><<<grid, block, 0, s>>>(a, inv_scale); break;
      case 1: gather_fast_mp_kernel<P, 1, U><<<grid, block, 0, s>>>(a, inv_scale); break;
      default: return cudaErrorInvalidValue;
    }
    return cudaGetLastError();
  }
  if (a.peer) {
    switch (32 / a.heads) {
      case 32: gather_fast_kernel<P, 32, U, true><<<grid, block, 0, s>>>(a, inv_scale); break;
      case 16: gather_fast_kernel<P, 16, U, true><<<grid, block, 0, s>>>(a, inv_scale); break;
      case 8: gather_fast_kernel<P, 8, U, true><<<grid, block, 0, s>>>(a, inv_scale); break;
      case 4: gather_fast_kernel<P, 4, U, true><<<grid, block, 0, s>>>(a, inv_scale); break;
      case 2: gather_fast_kernel<P, 2, U, true><<<grid, block, 0, s>>>(a, inv_scale); break;
      case 1: gather_fast_kernel<P, 1, U, true><<<grid, block, 0, s>>>(a, inv_scale); break;
      default: return cudaErrorInvalidValue;
    }
    return cudaGetLastError();
  }
  switch (32 / a.heads) {
    case 32: gather_fast_kernel<P, 32, U, false><<<grid, block, 0, s>>>(a, inv_scale); break;
    case 16: gather_fast_kernel<P, 16, U, false><<<grid, block, 0, s>>>(a, inv_scale); break;
    case 8: gather_fast_kernel<P, 8, U, false><<<grid, block, 0, s>>>(a, inv_scale); break;
    case 4: gather_fast_kernel<P, 4, U, false><<<grid, block, 0, s>>>(a, inv_scale); break;
    case 2: gather_fast_kernel<P, 2, U, false><<<grid, block, 0, s>>>(a, inv_scale); break;
    case 1: gather_fast_kernel<P, 1, U, false><<<grid, block, 0, s>>>(a, inv_scale); break;
    default: return cudaErrorInvalidValue;
  }
  return cudaGetLastError();
}

}  // namespace

cudaError_t launch_build_pass_csr(const uint64_t *row_ptr, const uint32_t *col, const float *w,
                                  uint64_t n_rows, const uint64_t *bounds_dev, int npass,
                                  int prenormalized, uint32_t *col_p, float *w_p, uint32_t *seg,
                                  cudaStream_t s, uint64_t *launches) {
  if (n_rows == 0) return cudaSuccess;
  if (launches) ++*launches;
  uint64_t blocks = (n_rows + 7) / 8;
  if (blocks > 148ull * 32) blocks = 148ull * 32;
  build_pass_csr_kernel<<<(unsigned)blocks, 256, 0, s>>>(row_ptr, col, w, n_rows, bounds_dev, npass,
                                                        prenormalized, col_p, w_p, seg);
  return cudaGetLastError();
}

cudaError_t launch_gather_r1(const GatherR1Args &a, int sm_count, cudaStream_t s,
                             uint64_t *launches) {
  if (a.row_end <= a.row_begin) return cudaSuccess;
  const float scale_log2 = 1.4426950408889634f / sqrtf((float)(128 / a.heads));
  if (launches) ++*launches;
  switch (a.heads) {
    case 1: return launch_r1_h<1>(a, scale_log2, sm_count, s);
    case 2: return launch_r1_h<2>(a, scale_log2, sm_count, s);
    case 4: return launch_r1_h<4>(a, scale_log2, sm_count, s);
    case 8: return launch_r1_h<8>(a, scale_log2, sm_count, s);
    default: return cudaErrorInvalidValue;
  }
}

cudaError_t launch_gather(const GatherArgs &a, bool fast, int sm_count, cudaStream_t s,
                          uint64_t *launches) {
  if (a.row_end <= a.row_begin) return cudaSuccess;
  const float inv_scale = 1.0f / sqrtf((float)(a.H / a.heads));
  if (launches) ++*launches;
  if (fast) {
    switch (a.H / 128) {
      case 1: return launch_fast_p<1, 4>(a, inv_scale, sm_count, s);
      case 2: return launch_fast_p<2, 2>(a, inv_scale, sm_count, s);
      case 3: return launch_fast_p<3, 2>(a, inv_scale, sm_count, s);
      case 4: return launch_fast_p<4, 1>(a, inv_scale, sm_count, s);
      case 6: return launch_fast_p<6, 1>(a, inv_scale, sm_count, s);
      case 8: return launch_fast_p<8, 1>(a, inv_scale, sm_count, s);
      default: return cudaErrorInvalidValue;
    }
  }
  if (a.multipass || a.peer) return cudaErrorNotSupported;  // generic layouts: single-pass, halo buffer only
  const size_t per_warp = (size_t)(3 * a.H + 2 * a.heads) * sizeof(float);
  int wpb = (int)((size_t)(96 * 1024) / per_warp);
  if (wpb > 4) wpb = 4;
  if (wpb < 1) wpb = 1;
  const size_t smem = per_warp * wpb;
  if (smem > 48 * 1024) {
    cudaError_t e = cudaFuncSetAttribute(gather_generic_kernel,
                                         cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
  }
  const uint64_t rows = a.row_end - a.row_begin;
  uint64_t blocks = (rows + wpb - 1) / wpb;
  const uint64_t cap = (uint64_t)sm_count * 8;
  if (blocks > cap) blocks = cap;
  gather_generic_kernel<<<(unsigned)blocks, 32 * wpb, smem, s>>>(a, inv_scale);
  return cudaGetLastError();
}

}  // namespace ruvgnn
