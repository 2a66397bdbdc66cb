// Range-partitioned forward inside the library (include/ruvgnn.h, "range-partitioned forward
// inside the library"): NCCL communicator per GPU, halo plan built on the device,
// halo exchange on a communication stream overlapped with the gather passes that do not need it.
//
// The reference has no multi-device path.  What is partitioned is the caller-side
// loop over nodes (crates/ruvector-graph-transformer/src/verified_training.rs:422-436): rank p
// owns a contiguous node range, projects its own rows (layer.rs:392-396,135-137) and needs the
// projected rows of the remote neighbours its rows reference.
//
// libnccl.so.2 is resolved with dlopen at first use, so libruvgnn.so itself has no link-time
// dependency on NCCL (it loads on machines without it; the dist entries then return
// RUVGNN_ERR_NCCL).
#include <dlfcn.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <memory>
#include <mutex>
#include <string>
#include <vector>

#include "abi_guard.h"
#include "engine.h"

using namespace ruvgnn;

// ---- NCCL, by name (ABI of nccl.h 2.x; these entry points and enum values have been stable
// since 2.0) ----------------------------------------------------------------------------------
namespace {

typedef struct ncclComm *ncclComm_t;
typedef struct {
  char internal[RUVGNN_NCCL_ID_BYTES];
} ncclUniqueId;
typedef int ncclResult_t;  // 0 = ncclSuccess
enum { NCCL_UINT8 = 1, NCCL_UINT32 = 3, NCCL_UINT64 = 5, NCCL_FLOAT32 = 7 };

struct NcclApi {
  void *handle = nullptr;
  ncclResult_t (*GetUniqueId)(ncclUniqueId *) = nullptr;
  ncclResult_t (*CommInitRank)(ncclComm_t *, int, ncclUniqueId, int) = nullptr;
  ncclResult_t (*CommInitAll)(ncclComm_t *, int, const int *) = nullptr;
  ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
  ncclResult_t (*CommGetAsyncError)(ncclComm_t, ncclResult_t *) = nullptr;
  ncclResult_t (*AllGather)(const void *, void *, size_t, int, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*Send)(const void *, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*Recv)(void *, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*GroupStart)() = nullptr;
  ncclResult_t (*GroupEnd)() = nullptr;
  const char *(*GetErrorString)(ncclResult_t) = nullptr;
  std::string error;
};

NcclApi *nccl() {
  static NcclApi api;
  static std::once_flag once;
  std::call_once(once, [] {
    const char *env = getenv("RUVGNN_NCCL_LIB");
    // an NCCL already in the process (e.g. the one torch ships) is reused, never duplicated
    void *h = env ? dlopen(env, RTLD_NOW | RTLD_GLOBAL) : dlopen("libnccl.so.2", RTLD_NOW | RTLD_NOLOAD);
    if (!h && !env) h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
    if (!h) {
      const char *e = dlerror();
      api.error = std::string("libnccl.so.2 could not be loaded: ") + (e ? e : "unknown error");
      return;
    }
    api.handle = h;
    bool ok = true;
    auto sym = [&](const char *name) {
      void *p = dlsym(h, name);
      if (!p) {
        ok = false;
        api.error = std::string("libnccl.so.2 lacks ") + name;
      }
      return p;
    };
#define NCCL_SYM(field, name) api.field = reinterpret_cast<decltype(api.field)>(sym(name))
    NCCL_SYM(GetUniqueId, "ncclGetUniqueId");
    NCCL_SYM(CommInitRank, "ncclCommInitRank");
    NCCL_SYM(CommInitAll, "ncclCommInitAll");
    NCCL_SYM(CommDestroy, "ncclCommDestroy");
    NCCL_SYM(CommGetAsyncError, "ncclCommGetAsyncError");
    NCCL_SYM(AllGather, "ncclAllGather");
    NCCL_SYM(Send, "ncclSend");
    NCCL_SYM(Recv, "ncclRecv");
    NCCL_SYM(GroupStart, "ncclGroupStart");
    NCCL_SYM(GroupEnd, "ncclGroupEnd");
    NCCL_SYM(GetErrorString, "ncclGetErrorString");
#undef NCCL_SYM
    if (!ok) api.handle = nullptr;
  });
  return &api;
}

int32_t nccl_fail(const char *what, ncclResult_t r) {
  NcclApi *a = nccl();
  return fail(RUVGNN_ERR_NCCL, std::string(what) + ": " +
                                   (a->GetErrorString ? a->GetErrorString(r) : "NCCL error"));
}
#define NCCL_TRY(expr)                                  \
  do {                                                  \
    ncclResult_t _r = (expr);                           \
    if (_r != 0) return nccl_fail(#expr, _r);           \
  } while (0)

constexpr int kMaxChunks = 16;
constexpr size_t kXchgBytes = 128;  // cudaIpcMemHandle_t (64) + ok flag, padded

}  // namespace

struct ruvgnn_dist {
  ruvgnn_ctx *ctx = nullptr;
  int rank = 0, world = 1;
  ncclComm_t comm = nullptr;
  cudaStream_t comm_stream = nullptr;
  cudaEvent_t ev_ready = nullptr;              // compute stream: halo source rows are ready
  cudaEvent_t ev_chunk[kMaxChunks] = {};       // comm stream: chunk c has landed
  cudaEvent_t ev_x0 = nullptr, ev_x1 = nullptr;  // comm stream: exchange start / end (timed)
  cudaEvent_t ev_w0[kMaxChunks] = {}, ev_w1[kMaxChunks] = {};  // compute stream around each wait
  uint32_t waits_used = 0;
  float *send = nullptr, *halo = nullptr, *mstage = nullptr;
  size_t send_floats = 0, halo_floats = 0, mstage_floats = 0;
  uint32_t chunks = 0;  // 0: chosen from the world size when a partition is created
  // peer-direct plane: every rank maps its peers' workspaces with CUDA IPC (one process per GPU only)
  bool ipc_capable = false;
  struct PeerMap {
    cudaIpcMemHandle_t handle;
    void *ptr = nullptr;
  } peers[kMaxPeers];
  PeerTable *peer_table_dev = nullptr;
  unsigned char *xchg_dev = nullptr;  // [world + 1][kXchgBytes]: all-gathered records | own record
  cudaEvent_t ev_gdone = nullptr, ev_free = nullptr, ev_xchg = nullptr;
  bool free_pending = false;  // ev_free: every rank has finished reading this rank's table
  bool offline = false;  // world > 1 without a communicator: plans can be built, nothing exchanged
  ruvgnn_dist_stats stats = {};
  bool stats_valid = false;
};

struct ruvgnn_partition {
  ruvgnn_dist *d = nullptr;
  std::vector<uint64_t> bounds;
  uint64_t n_local = 0, n_total = 0, n_edges = 0, n_halo = 0, n_send = 0, remote_edges = 0;
  ruvgnn_graph graph;  // row_ptr / renamed col_idx / edge_w (owned here)
  uint32_t *need_ids = nullptr;  // [n_halo] global ids of the halo slots (sparse plane)
  uint32_t *send_idx = nullptr;  // [n_send] own-local rows to send, grouped by peer
  uint32_t *col_global = nullptr;  // [n_edges] the caller's global column ids (peer-direct plane)
  std::vector<uint64_t> recv_counts, send_counts;
  bool dense = false;
  std::vector<uint64_t> starts;  // dense plane: row chunk boundaries
  std::vector<std::pair<uint64_t, uint64_t>> interior, boundary;
  uint64_t interior_rows = 0;
  PassCsr pass;  // pass-ordered edge arrays of the overlapped forward: own rows, then one group per chunk
};

namespace {

// ---- plan kernels (integer work, HBM / L2-atomic bound) ------------------------------------
__global__ void mark_remote_kernel(const uint32_t *col, uint64_t n_edges, uint64_t lo, uint64_t hi,
                                   unsigned long long *bits, unsigned long long *remote_edges) {
  const uint64_t tid = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
  unsigned long long mine = 0;
  for (uint64_t e = tid; e < n_edges; e += stride) {
    const uint64_t c = col[e];
    if (c < lo || c >= hi) {
      const unsigned long long bit = 1ull << (c & 63);
      if (!(bits[c >> 6] & bit)) atomicOr(&bits[c >> 6], bit);
      ++mine;
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) mine += __shfl_xor_sync(0xffffffffu, mine, o);
  if ((threadIdx.x & 31) == 0 && mine) atomicAdd(remote_edges, mine);
}

// block sums of popcounts, 1024 words per block
__global__ void __launch_bounds__(256) popc_block_kernel(const unsigned long long *bits,
                                                         uint64_t n_words, uint32_t *block_sum) {
  __shared__ uint32_t red[256];
  const uint64_t base = (uint64_t)blockIdx.x * 1024;
  uint32_t s = 0;
  for (int k = 0; k < 4; ++k) {
    const uint64_t w = base + k * 256 + threadIdx.x;
    if (w < n_words) s += __popcll(bits[w]);
  }
  red[threadIdx.x] = s;
  __syncthreads();
  for (int o = 128; o > 0; o >>= 1) {
    if (threadIdx.x < o) red[threadIdx.x] += red[threadIdx.x + o];
    __syncthreads();
  }
  if (threadIdx.x == 0) block_sum[blockIdx.x] = red[0];
}

// pre[w] = number of set bits in words [0, w); block_off = exclusive scan of the block sums
__global__ void __launch_bounds__(256) popc_scan_kernel(const unsigned long long *bits,
                                                        uint64_t n_words, const uint32_t *block_off,
                                                        uint32_t *pre) {
  __shared__ uint32_t part[256];
  const uint64_t base = (uint64_t)blockIdx.x * 1024;
  uint32_t c[4], s = 0;
  for (int k = 0; k < 4; ++k) {  // thread t owns words base + 4t .. base + 4t + 3
    const uint64_t w = base + (uint64_t)threadIdx.x * 4 + k;
    c[k] = w < n_words ? __popcll(bits[w]) : 0;
    s += c[k];
  }
  part[threadIdx.x] = s;
  __syncthreads();
  for (int o = 1; o < 256; o <<= 1) {  // inclusive scan of the per-thread sums
    const uint32_t v = threadIdx.x >= o ? part[threadIdx.x - o] : 0;
    __syncthreads();
    part[threadIdx.x] += v;
    __syncthreads();
  }
  uint32_t run = block_off[blockIdx.x] + part[threadIdx.x] - s;
  for (int k = 0; k < 4; ++k) {
    const uint64_t w = base + (uint64_t)threadIdx.x * 4 + k;
    if (w < n_words) pre[w] = run;
    run += c[k];
  }
}

__device__ __forceinline__ uint32_t rank_of(const unsigned long long *bits, const uint32_t *pre,
                                            uint64_t id) {
  const unsigned long long below = bits[id >> 6] & ((1ull << (id & 63)) - 1ull);
  return pre[id >> 6] + (uint32_t)__popcll(below);
}

// sparse plane: own rows -> [0, n_local), remote id -> n_local + slot(id)
__global__ void remap_sparse_kernel(uint32_t *col, uint64_t n_edges, uint64_t lo, uint64_t hi,
                                    const unsigned long long *bits, const uint32_t *pre) {
  const uint64_t tid = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
  const uint32_t n_local = (uint32_t)(hi - lo);
  for (uint64_t e = tid; e < n_edges; e += stride) {
    const uint64_t c = col[e];
    col[e] = (c < lo || c >= hi) ? n_local + rank_of(bits, pre, c) : (uint32_t)(c - lo);
  }
}

// dense plane: the remote row (owner r, local row i) lives at n_local + world*starts[k] + r*size_k
// + (i - starts[k]), k = the row chunk of i (chunk k of all ranks lands in one contiguous region)
__global__ void remap_dense_kernel(uint32_t *col, uint64_t n_edges, uint64_t lo, uint64_t hi,
                                   uint64_t per, uint32_t world, const uint64_t *starts,
                                   uint32_t n_chunks) {
  const uint64_t tid = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
  for (uint64_t e = tid; e < n_edges; e += stride) {
    const uint64_t c = col[e];
    if (c >= lo && c < hi) {
      col[e] = (uint32_t)(c - lo);
      continue;
    }
    const uint64_t r = c / per, i = c % per;
    uint32_t k = 0;
    while (k + 1 < n_chunks && i >= starts[k + 1]) ++k;
    col[e] = (uint32_t)(per + world * starts[k] + r * (starts[k + 1] - starts[k]) + (i - starts[k]));
  }
}

__global__ void need_ids_kernel(const unsigned long long *bits, uint64_t n_words,
                                const uint32_t *pre, uint32_t *need) {
  const uint64_t tid = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
  for (uint64_t w = tid; w < n_words; w += stride) {
    unsigned long long b = bits[w];
    uint32_t at = pre[w];
    while (b) {
      const int k = __ffsll((long long)b) - 1;
      need[at++] = (uint32_t)(w * 64 + k);
      b &= b - 1;
    }
  }
}

__global__ void owner_edges_kernel(const unsigned long long *bits, const uint32_t *pre,
                                   const uint64_t *bounds, uint32_t world, uint64_t n_total,
                                   uint64_t n_halo, uint64_t *owner_edges) {
  const uint32_t p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p > world) return;
  owner_edges[p] = bounds[p] >= n_total ? n_halo : rank_of(bits, pre, bounds[p]);
}

__global__ void sub_offset_kernel(uint32_t *ids, uint64_t n, const uint64_t *owner_edges,
                                  const uint64_t *bounds, uint32_t world) {
  const uint64_t tid = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
  for (uint64_t s = tid; s < n; s += stride) {
    uint32_t p = 0;
    while (p + 1 < world && s >= owner_edges[p + 1]) ++p;
    ids[s] -= (uint32_t)bounds[p];
  }
}

// rows whose neighbours are all own rows
__global__ void interior_flag_kernel(const uint64_t *row_ptr, const uint32_t *col, uint64_t n_local,
                                     uint8_t *is_interior) {
  const int lane = threadIdx.x & 31;
  const uint64_t warp = ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const uint64_t nwarps = ((uint64_t)gridDim.x * blockDim.x) >> 5;
  for (uint64_t r = warp; r < n_local; r += nwarps) {
    bool remote = false;
    for (uint64_t e = row_ptr[r] + lane; e < row_ptr[r + 1]; e += 32) remote |= col[e] >= n_local;
    remote = __any_sync(0xffffffffu, remote);
    if (lane == 0) is_interior[r] = remote ? 0 : 1;
  }
}

__global__ void check_send_idx_kernel(const uint32_t *idx, uint64_t n, uint64_t n_local, int *flag) {
  const uint64_t tid = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
  for (uint64_t s = tid; s < n; s += stride)
    if (idx[s] >= n_local) atomicOr(flag, 1);
}

struct DevMem {  // RAII device scratch
  void *p = nullptr;
  ~DevMem() {
    if (p) cudaFree(p);
  }
  cudaError_t alloc(size_t bytes) { return cudaMalloc(&p, bytes ? bytes : 16); }
  template <class T>
  T *as() const {
    return static_cast<T *>(p);
  }
  void *release() {
    void *q = p;
    p = nullptr;
    return q;
  }
};

unsigned grid_1d(uint64_t n, int block = 256) {
  uint64_t b = (n + block - 1) / block;
  if (b > 148ull * 32) b = 148ull * 32;
  if (b == 0) b = 1;
  return (unsigned)b;
}

// The (at most max_runs) longest contiguous runs of >= min_rows interior rows, sorted; the
// complement as boundary ranges.  Range partitions of a locality-ordered graph keep their boundary
// rows at the two ends of the range, and contiguous runs keep every kernel on plain row blocks.
void interior_runs(const std::vector<uint8_t> &is_int, uint64_t min_rows, size_t max_runs,
                   std::vector<std::pair<uint64_t, uint64_t>> &interior,
                   std::vector<std::pair<uint64_t, uint64_t>> &boundary) {
  const uint64_t n = is_int.size();
  std::vector<std::pair<uint64_t, uint64_t>> runs;
  for (uint64_t i = 0; i < n;) {
    if (!is_int[i]) {
      ++i;
      continue;
    }
    uint64_t j = i;
    while (j < n && is_int[j]) ++j;
    if (j - i >= min_rows) runs.push_back({i, j});
    i = j;
  }
  std::stable_sort(runs.begin(), runs.end(),
                   [](const auto &a, const auto &b) { return a.second - a.first > b.second - b.first; });
  if (runs.size() > max_runs) runs.resize(max_runs);
  std::sort(runs.begin(), runs.end());
  interior = runs;
  boundary.clear();
  uint64_t at = 0;
  for (const auto &r : runs) {
    if (r.first > at) boundary.push_back({at, r.first});
    at = r.second;
  }
  if (at < n) boundary.push_back({at, n});
}

int32_t ensure_buf(ruvgnn_ctx *c, float **buf, size_t *have, size_t need) {
  if (need <= *have) return RUVGNN_OK;
  if (*buf) {
    RUVGNN_CUDA_TRY(cudaDeviceSynchronize());
    cudaFree(*buf);
    *buf = nullptr;
    *have = 0;
  }
  (void)c;
  RUVGNN_CUDA_TRY(cudaMalloc(buf, need * sizeof(float)));
  *have = need;
  return RUVGNN_OK;
}

// All-gather of one kXchgBytes record per rank on the communication stream, read back by the host.
int32_t exchange_records(ruvgnn_dist *d, const unsigned char *mine, std::vector<unsigned char> *all) {
  NcclApi *a = nccl();
  cudaStream_t cs = d->comm_stream;
  unsigned char *own = d->xchg_dev + (size_t)d->world * kXchgBytes;
  RUVGNN_CUDA_TRY(cudaMemcpyAsync(own, mine, kXchgBytes, cudaMemcpyHostToDevice, cs));
  NCCL_TRY(a->AllGather(own, d->xchg_dev, kXchgBytes, NCCL_UINT8, d->comm, cs));
  all->resize((size_t)d->world * kXchgBytes);
  RUVGNN_CUDA_TRY(cudaMemcpyAsync(all->data(), d->xchg_dev, all->size(), cudaMemcpyDeviceToHost, cs));
  RUVGNN_CUDA_TRY(cudaStreamSynchronize(cs));
  return RUVGNN_OK;
}

// Maps peer p's allocation (handle from its record) unless that very allocation is already mapped.
cudaError_t map_peer(ruvgnn_dist *d, int p, const cudaIpcMemHandle_t &h) {
  ruvgnn_dist::PeerMap &m = d->peers[p];
  if (m.ptr && memcmp(&m.handle, &h, sizeof h) == 0) return cudaSuccess;
  if (m.ptr) {
    cudaIpcCloseMemHandle(m.ptr);
    m.ptr = nullptr;
  }
  cudaError_t e = cudaIpcOpenMemHandle(&m.ptr, h, cudaIpcMemLazyEnablePeerAccess);
  if (e != cudaSuccess) {
    m.ptr = nullptr;
    cudaGetLastError();
    return e;
  }
  m.handle = h;
  return cudaSuccess;
}

// Collective probe at creation: can every rank map every other rank's device memory?
int32_t probe_ipc(ruvgnn_dist *d) {
  d->ipc_capable = false;
  if (d->world < 2 || d->world > kMaxPeers || !d->comm) return RUVGNN_OK;
  const char *env = getenv("RUVGNN_PEER_DIRECT");
  const bool want = !(env && atoi(env) == 0);
  RUVGNN_CUDA_TRY(cudaMalloc(&d->xchg_dev, (size_t)(d->world + 1) * kXchgBytes));
  RUVGNN_CUDA_TRY(cudaMalloc(&d->peer_table_dev, sizeof(PeerTable)));
  void *probe = nullptr;
  RUVGNN_CUDA_TRY(cudaMalloc(&probe, 4096));
  unsigned char rec[kXchgBytes] = {};
  cudaIpcMemHandle_t h;
  bool ok = want && cudaIpcGetMemHandle(&h, probe) == cudaSuccess;
  if (!ok) cudaGetLastError();
  if (ok) memcpy(rec, &h, sizeof h);
  rec[64] = ok ? 1 : 0;
  std::vector<unsigned char> all;
  int32_t rc = exchange_records(d, rec, &all);
  if (rc) {
    cudaFree(probe);
    return rc;
  }
  for (int p = 0; p < d->world && ok; ++p) {
    if (!all[(size_t)p * kXchgBytes + 64]) ok = false;
    if (p == d->rank || !ok) continue;
    cudaIpcMemHandle_t ph;
    memcpy(&ph, &all[(size_t)p * kXchgBytes], sizeof ph);
    void *q = nullptr;
    if (cudaIpcOpenMemHandle(&q, ph, cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) {
      cudaGetLastError();
      ok = false;
    } else {
      cudaIpcCloseMemHandle(q);
    }
  }
  memset(rec, 0, sizeof rec);
  rec[64] = ok ? 1 : 0;
  rc = exchange_records(d, rec, &all);  // the verdict must be unanimous (also keeps `probe` alive until all are done)
  cudaFree(probe);
  if (rc) return rc;
  for (int p = 0; p < d->world; ++p) ok = ok && all[(size_t)p * kXchgBytes + 64];
  d->ipc_capable = ok;
  return RUVGNN_OK;
}

// Pass-ordered copy of G's edge arrays for the column bounds given (host, npass + 1 ascending values).
int32_t build_pass_csr(ruvgnn_ctx *c, const ruvgnn_graph *G, const std::vector<uint64_t> &bounds, PassCsr *out) {
  const int npass = (int)bounds.size() - 1;
  DevMem d_bounds;
  RUVGNN_CUDA_TRY(d_bounds.alloc(bounds.size() * 8));
  RUVGNN_CUDA_TRY(cudaMemcpyAsync(d_bounds.p, bounds.data(), bounds.size() * 8, cudaMemcpyHostToDevice, c->stream));
  RUVGNN_CUDA_TRY(cudaMalloc(&out->col, (G->n_edges + 4) * sizeof(uint32_t)));
  RUVGNN_CUDA_TRY(cudaMalloc(&out->w, (G->n_edges + 4) * sizeof(float)));
  RUVGNN_CUDA_TRY(cudaMalloc(&out->seg, (G->n_nodes * (uint64_t)(npass + 1) + 4) * sizeof(uint32_t)));
  out->npass = npass;
  RUVGNN_CUDA_TRY(launch_build_pass_csr(G->row_ptr, G->col_idx, G->edge_w, G->n_nodes, d_bounds.as<uint64_t>(),
                                        npass, (G->flags & RUVGNN_GRAPH_PRENORMALIZED) ? 1 : 0, out->col, out->w,
                                        out->seg, c->stream, &c->launches));
  RUVGNN_CUDA_TRY(cudaStreamSynchronize(c->stream));
  return RUVGNN_OK;
}

int32_t make_dist(ruvgnn_ctx *ctx, ncclComm_t comm, int rank, int world, ruvgnn_dist **out) {
  std::unique_ptr<ruvgnn_dist> d(new ruvgnn_dist());
  d->ctx = ctx;
  d->rank = rank;
  d->world = world;
  d->comm = comm;
  RUVGNN_CUDA_TRY(cudaStreamCreateWithFlags(&d->comm_stream, cudaStreamNonBlocking));
  RUVGNN_CUDA_TRY(cudaEventCreateWithFlags(&d->ev_ready, cudaEventDisableTiming));
  RUVGNN_CUDA_TRY(cudaEventCreate(&d->ev_x0));
  RUVGNN_CUDA_TRY(cudaEventCreate(&d->ev_x1));
  RUVGNN_CUDA_TRY(cudaEventCreateWithFlags(&d->ev_gdone, cudaEventDisableTiming));
  RUVGNN_CUDA_TRY(cudaEventCreateWithFlags(&d->ev_free, cudaEventDisableTiming));
  RUVGNN_CUDA_TRY(cudaEventCreateWithFlags(&d->ev_xchg, cudaEventDisableTiming));
  for (int i = 0; i < kMaxChunks; ++i) {
    RUVGNN_CUDA_TRY(cudaEventCreateWithFlags(&d->ev_chunk[i], cudaEventDisableTiming));
    RUVGNN_CUDA_TRY(cudaEventCreate(&d->ev_w0[i]));
    RUVGNN_CUDA_TRY(cudaEventCreate(&d->ev_w1[i]));
  }
  *out = d.release();
  return RUVGNN_OK;
}

}  // namespace

extern "C" {

int32_t ruvgnn_dist_unique_id(uint8_t *id) {
  RUVGNN_ABI_TRY
  if (!id) return fail(RUVGNN_ERR_INVALID_INPUT, "id is null");
  NcclApi *a = nccl();
  if (!a->handle) return fail(RUVGNN_ERR_NCCL, a->error);
  ncclUniqueId u;
  NCCL_TRY(a->GetUniqueId(&u));
  memcpy(id, u.internal, RUVGNN_NCCL_ID_BYTES);
  return RUVGNN_OK;
  RUVGNN_ABI_CATCH
}

int32_t ruvgnn_dist_create(ruvgnn_ctx *ctx, const uint8_t *id, int32_t rank, int32_t world,
                           ruvgnn_dist **out) {
  RUVGNN_ABI_TRY
  if (!out) return fail(RUVGNN_ERR_INVALID_INPUT, "out is null");
  *out = nullptr;
  if (!ctx) return fail(RUVGNN_ERR_INVALID_INPUT, "ctx is null");
  if (world < 1 || rank < 0 || rank >= world)
    return fail(RUVGNN_ERR_INVALID_INPUT, "rank / world out of range");
  DeviceGuard dg(ctx->device);
  if (!dg.ok) return fail(RUVGNN_ERR_CUDA, "cudaSetDevice failed");
  ncclComm_t comm = nullptr;
  if (world > 1 && id) {
    NcclApi *a = nccl();
    if (!a->handle) return fail(RUVGNN_ERR_NCCL, a->error);
    ncclUniqueId u;
    memcpy(u.internal, id, RUVGNN_NCCL_ID_BYTES);
    NCCL_TRY(a->CommInitRank(&comm, world, u, rank));
  }
  int32_t rc = make_dist(ctx, comm, rank, world, out);
  if (rc && comm) nccl()->CommDestroy(comm);
  if (!rc) (*out)->offline = world > 1 && !id;
  if (!rc && comm) {
    rc = probe_ipc(*out);
    if (rc) {
      ruvgnn_dist_destroy(*out);
      *out = nullptr;
    }
  }
  return rc;
  RUVGNN_ABI_CATCH
}

int32_t ruvgnn_dist_create_all(ruvgnn_ctx *const *ctxs, int32_t world, ruvgnn_dist **out) {
  RUVGNN_ABI_TRY
  if (!ctxs || !out || world < 1) return fail(RUVGNN_ERR_INVALID_INPUT, "null argument");
  for (int i = 0; i < world; ++i) out[i] = nullptr;
  std::vector<int> devs(world);
  for (int i = 0; i < world; ++i) {
    if (!ctxs[i]) return fail(RUVGNN_ERR_INVALID_INPUT, "ctx is null");
    devs[i] = ctxs[i]->device;
    for (int j = 0; j < i; ++j)
      if (devs[j] == devs[i]) return fail(RUVGNN_ERR_INVALID_INPUT, "two ranks on one device");
  }
  std::vector<ncclComm_t> comms(world, nullptr);
  if (world > 1) {
    NcclApi *a = nccl();
    if (!a->handle) return fail(RUVGNN_ERR_NCCL, a->error);
    NCCL_TRY(a->CommInitAll(comms.data(), world, devs.data()));
  }
  for (int i = 0; i < world; ++i) {
    DeviceGuard dg(ctxs[i]->device);
    int32_t rc = make_dist(ctxs[i], comms[i], i, world, &out[i]);
    if (rc) {
      for (int j = 0; j < world; ++j) {
        if (out[j]) {
          out[j]->comm = nullptr;
          ruvgnn_dist_destroy(out[j]);
          out[j] = nullptr;
        }
        if (comms[j]) nccl()->CommDestroy(comms[j]);
      }
      return rc;
    }
  }
  return RUVGNN_OK;
  RUVGNN_ABI_CATCH
}

void ruvgnn_dist_destroy(ruvgnn_dist *d) {
  if (!d) return;
  DeviceGuard dg(d->ctx->device);
  cudaStreamSynchronize(d->comm_stream);
  cudaStreamSynchronize(d->ctx->stream);
  if (d->comm) nccl()->CommDestroy(d->comm);
  if (d->send) cudaFree(d->send);
  if (d->halo) cudaFree(d->halo);
  if (d->mstage) cudaFree(d->mstage);
  for (int p = 0; p < kMaxPeers; ++p)
    if (d->peers[p].ptr) cudaIpcCloseMemHandle(d->peers[p].ptr);
  if (d->peer_table_dev) cudaFree(d->peer_table_dev);
  if (d->xchg_dev) cudaFree(d->xchg_dev);
  if (d->ev_gdone) cudaEventDestroy(d->ev_gdone);
  if (d->ev_free) cudaEventDestroy(d->ev_free);
  if (d->ev_xchg) cudaEventDestroy(d->ev_xchg);
  if (d->ev_ready) cudaEventDestroy(d->ev_ready);
  if (d->ev_x0) cudaEventDestroy(d->ev_x0);
  if (d->ev_x1) cudaEventDestroy(d->ev_x1);
  for (int i = 0; i < kMaxChunks; ++i) {
    if (d->ev_chunk[i]) cudaEventDestroy(d->ev_chunk[i]);
    if (d->ev_w0[i]) cudaEventDestroy(d->ev_w0[i]);
    if (d->ev_w1[i]) cudaEventDestroy(d->ev_w1[i]);
  }
  if (d->comm_stream) cudaStreamDestroy(d->comm_stream);
  delete d;
}

int32_t ruvgnn_dist_info(const ruvgnn_dist *d, int32_t *rank, int32_t *world) {
  RUVGNN_ABI_TRY
  if (!d) return fail(RUVGNN_ERR_INVALID_INPUT, "dist is null");
  if (rank) *rank = d->rank;
  if (world) *world = d->world;
  return RUVGNN_OK;
  RUVGNN_ABI_CATCH
}

int32_t ruvgnn_dist_set_chunks(ruvgnn_dist *d, uint32_t chunks) {
  RUVGNN_ABI_TRY
  if (!d) return fail(RUVGNN_ERR_INVALID_INPUT, "dist is null");
  if (chunks > (uint32_t)kMaxChunks)
    return fail(RUVGNN_ERR_INVALID_INPUT, "chunks must be in [0, 16] (0: chosen from the world size)");
  d->chunks = chunks;
  return RUVGNN_OK;
  RUVGNN_ABI_CATCH
}

void ruvgnn_partition_destroy(ruvgnn_partition *p) {
  if (!p) return;
  DeviceGuard dg(p->graph.device);
  if (p->graph.row_ptr) cudaFree(p->graph.row_ptr);
  if (p->graph.col_idx) cudaFree(p->graph.col_idx);
  if (p->graph.edge_w) cudaFree(p->graph.edge_w);
  if (p->need_ids) cudaFree(p->need_ids);
  if (p->send_idx) cudaFree(p->send_idx);
  if (p->col_global) cudaFree(p->col_global);
  if (p->pass.col) cudaFree(p->pass.col);
  if (p->pass.w) cudaFree(p->pass.w);
  if (p->pass.seg) cudaFree(p->pass.seg);
  delete p;
}

int32_t ruvgnn_partition_create(ruvgnn_dist *d, const uint64_t *bounds, const uint64_t *row_ptr,
                                const uint32_t *col_idx, const float *edge_w, uint32_t flags,
                                ruvgnn_partition **out) {
  RUVGNN_ABI_TRY
  if (!out) return fail(RUVGNN_ERR_INVALID_INPUT, "out is null");
  *out = nullptr;
  if (!d || !bounds || !row_ptr) return fail(RUVGNN_ERR_INVALID_INPUT, "null argument");
  ruvgnn_ctx *c = d->ctx;
  const int world = d->world, rank = d->rank;
  if (bounds[0] != 0) return fail(RUVGNN_ERR_INVALID_INPUT, "bounds[0] must be 0");
  for (int p = 0; p < world; ++p)
    if (bounds[p + 1] < bounds[p]) return fail(RUVGNN_ERR_INVALID_INPUT, "bounds must be monotone");
  const uint64_t n_total = bounds[world], lo = bounds[rank], hi = bounds[rank + 1], n = hi - lo;
  if (n_total > 0xffffffffull) return fail(RUVGNN_ERR_INVALID_INPUT, "node ids exceed u32 col_idx");
  DeviceGuard dg(c->device);
  if (!dg.ok) return fail(RUVGNN_ERR_CUDA, "cudaSetDevice failed");
  const bool dev_ptrs = flags & RUVGNN_PART_DEVICE_PTRS;
  cudaStream_t s = c->stream;

  std::unique_ptr<ruvgnn_partition, void (*)(ruvgnn_partition *)> P(new ruvgnn_partition(),
                                                                    ruvgnn_partition_destroy);
  P->d = d;
  P->bounds.assign(bounds, bounds + world + 1);
  P->n_local = n;
  P->n_total = n_total;
  ruvgnn_graph &G = P->graph;
  G.device = c->device;
  G.owns = false;
  G.n_nodes = n;
  RUVGNN_CUDA_TRY(cudaMalloc(&G.row_ptr, (n + 1) * sizeof(uint64_t)));
  RUVGNN_CUDA_TRY(cudaMemcpyAsync(G.row_ptr, row_ptr, (n + 1) * sizeof(uint64_t),
                                  dev_ptrs ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice, s));
  uint64_t ends[2] = {0, 0};
  RUVGNN_CUDA_TRY(cudaMemcpyAsync(&ends[0], G.row_ptr, sizeof(uint64_t), cudaMemcpyDeviceToHost, s));
  RUVGNN_CUDA_TRY(cudaMemcpyAsync(&ends[1], G.row_ptr + n, sizeof(uint64_t), cudaMemcpyDeviceToHost, s));
  RUVGNN_CUDA_TRY(cudaStreamSynchronize(s));
  if (ends[0] != 0) return fail(RUVGNN_ERR_INVALID_INPUT, "row_ptr[0] must be 0");
  const uint64_t E = ends[1];
  if (E && (!col_idx || !edge_w)) return fail(RUVGNN_ERR_INVALID_INPUT, "col_idx / edge_w is null");
  P->n_edges = G.n_edges = E;
  RUVGNN_CUDA_TRY(cudaMalloc(&G.col_idx, (E + 4) * sizeof(uint32_t)));
  RUVGNN_CUDA_TRY(cudaMalloc(&G.edge_w, (E + 4) * sizeof(float)));
  if (E) {
    const cudaMemcpyKind k = dev_ptrs ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice;
    RUVGNN_CUDA_TRY(cudaMemcpyAsync(G.col_idx, col_idx, E * sizeof(uint32_t), k, s));
    RUVGNN_CUDA_TRY(cudaMemcpyAsync(G.edge_w, edge_w, E * sizeof(float), k, s));
  }
  // validate: monotone row_ptr, global columns inside the node range
  RUVGNN_CUDA_TRY(cudaMemsetAsync(c->d_flag, 0, sizeof(int), s));
  RUVGNN_CUDA_TRY(launch_validate_csr(G.row_ptr, G.col_idx, n, n_total, E, c->d_flag, s, &c->launches));
  int flag = 0;
  RUVGNN_CUDA_TRY(cudaMemcpyAsync(&flag, c->d_flag, sizeof(int), cudaMemcpyDeviceToHost, s));
  RUVGNN_CUDA_TRY(cudaStreamSynchronize(s));
  if (flag)
    return fail(RUVGNN_ERR_INVALID_INPUT, (flag & 1) ? "row_ptr is not monotone non-decreasing"
                                                     : "column id outside the global node range");

  if (world > 1 && !d->offline && E) {  // kept for the peer-direct plane
    RUVGNN_CUDA_TRY(cudaMalloc(&P->col_global, (E + 4) * sizeof(uint32_t)));
    RUVGNN_CUDA_TRY(cudaMemcpyAsync(P->col_global, G.col_idx, E * sizeof(uint32_t), cudaMemcpyDeviceToDevice, s));
  }
  // ---- de-duplicate the remote columns: bitmap + prefix popcount --------------------------
  const uint64_t n_words = (n_total + 63) / 64 + 1;
  const uint64_t n_blocks = (n_words + 1023) / 1024;
  DevMem bits, pre, bsum, d_remote, d_bounds, d_oe;
  RUVGNN_CUDA_TRY(bits.alloc(n_words * 8));
  RUVGNN_CUDA_TRY(pre.alloc(n_words * 4));
  RUVGNN_CUDA_TRY(bsum.alloc(n_blocks * 4));
  RUVGNN_CUDA_TRY(d_remote.alloc(8));
  RUVGNN_CUDA_TRY(d_bounds.alloc((world + 1) * 8));
  RUVGNN_CUDA_TRY(d_oe.alloc((world + 1) * 8));
  RUVGNN_CUDA_TRY(cudaMemsetAsync(bits.p, 0, n_words * 8, s));
  RUVGNN_CUDA_TRY(cudaMemsetAsync(d_remote.p, 0, 8, s));
  RUVGNN_CUDA_TRY(cudaMemcpyAsync(d_bounds.p, bounds, (world + 1) * 8, cudaMemcpyHostToDevice, s));
  if (E) {
    mark_remote_kernel<<<grid_1d(E), 256, 0, s>>>(G.col_idx, E, lo, hi, bits.as<unsigned long long>(),
                                                  d_remote.as<unsigned long long>());
    ++c->launches;
  }
  popc_block_kernel<<<(unsigned)n_blocks, 256, 0, s>>>(bits.as<unsigned long long>(), n_words,
                                                       bsum.as<uint32_t>());
  std::vector<uint32_t> hsum(n_blocks);
  RUVGNN_CUDA_TRY(cudaMemcpyAsync(hsum.data(), bsum.p, n_blocks * 4, cudaMemcpyDeviceToHost, s));
  RUVGNN_CUDA_TRY(cudaMemcpyAsync(&P->remote_edges, d_remote.p, 8, cudaMemcpyDeviceToHost, s));
  RUVGNN_CUDA_TRY(cudaStreamSynchronize(s));
  uint64_t n_halo = 0;
  for (uint64_t b = 0; b < n_blocks; ++b) {
    const uint32_t v = hsum[b];
    hsum[b] = (uint32_t)n_halo;
    n_halo += v;
  }
  if (n + n_halo > 0xffffffffull)
    return fail(RUVGNN_ERR_INVALID_INPUT, "local + halo rows exceed u32 column indices");
  P->n_halo = n_halo;
  RUVGNN_CUDA_TRY(cudaMemcpyAsync(bsum.p, hsum.data(), n_blocks * 4, cudaMemcpyHostToDevice, s));
  popc_scan_kernel<<<(unsigned)n_blocks, 256, 0, s>>>(bits.as<unsigned long long>(), n_words,
                                                      bsum.as<uint32_t>(), pre.as<uint32_t>());
  owner_edges_kernel<<<1, 64 < world + 1 ? ((world + 1 + 31) / 32) * 32 : 64, 0, s>>>(
      bits.as<unsigned long long>(), pre.as<uint32_t>(), d_bounds.as<uint64_t>(), (uint32_t)world,
      n_total, n_halo, d_oe.as<uint64_t>());
  c->launches += 3;
  std::vector<uint64_t> oe(world + 1);
  RUVGNN_CUDA_TRY(cudaMemcpyAsync(oe.data(), d_oe.p, (world + 1) * 8, cudaMemcpyDeviceToHost, s));
  RUVGNN_CUDA_TRY(cudaStreamSynchronize(s));
  P->recv_counts.resize(world);
  for (int p = 0; p < world; ++p) P->recv_counts[p] = oe[p + 1] - oe[p];
  if (P->recv_counts[rank] != 0) return fail(RUVGNN_ERR_INVALID_INPUT, "internal: own rows marked remote");

  // ---- exchange the per-peer request counts (one all-gather of world counters) ---------------
  std::vector<uint64_t> matrix((size_t)world * world, 0);  // [asker][owner]
  if (d->offline) {  // plan only: what the peers ask of this rank stays unknown
    for (int q = 0; q < world; ++q) matrix[(size_t)rank * world + q] = P->recv_counts[q];
  } else if (world > 1) {
    NcclApi *a = nccl();
    DevMem d_row, d_mat;
    RUVGNN_CUDA_TRY(d_row.alloc(world * 8));
    RUVGNN_CUDA_TRY(d_mat.alloc((size_t)world * world * 8));
    RUVGNN_CUDA_TRY(cudaMemcpyAsync(d_row.p, P->recv_counts.data(), world * 8, cudaMemcpyHostToDevice, s));
    NCCL_TRY(a->AllGather(d_row.p, d_mat.p, world, NCCL_UINT64, d->comm, s));
    RUVGNN_CUDA_TRY(cudaMemcpyAsync(matrix.data(), d_mat.p, (size_t)world * world * 8,
                                    cudaMemcpyDeviceToHost, s));
    RUVGNN_CUDA_TRY(cudaStreamSynchronize(s));
  }
  P->send_counts.resize(world);
  for (int p = 0; p < world; ++p) P->send_counts[p] = matrix[(size_t)p * world + rank];
  uint64_t n_send = 0;
  for (int p = 0; p < world; ++p) n_send += P->send_counts[p];

  // ---- dense or sparse: a collective decision from the same matrix on every rank --------------
  bool equal = true;
  for (int p = 0; p < world; ++p) equal = equal && (bounds[p + 1] - bounds[p]) == n;
  bool dense = false;
  if (world > 1 && equal && n > 0 && !d->offline) {
    dense = true;
    for (int p = 0; p < world; ++p) {
      uint64_t need = 0;
      for (int q = 0; q < world; ++q) need += matrix[(size_t)p * world + q];
      if ((double)need < 0.75 * (double)(n_total - n)) dense = false;
    }
  }
  if (flags & RUVGNN_PART_FORCE_SPARSE) dense = false;
  if (flags & RUVGNN_PART_FORCE_DENSE) {
    if (!equal) return fail(RUVGNN_ERR_INVALID_INPUT, "the dense plane needs equal ranges");
    dense = world > 1;
  }
  P->dense = dense;

  if (dense) {
    // Chunks let K | V of chunk k be rebuilt while chunk k+1 is on the wire (message-row halo), and in
    // the opt-in overlapped mode each chunk is one more gather pass.  4 measured best or equal at 2 and
    // 8 ranks (profiles/r02i_n2_*.json, r02g_n8_*.json).
    uint32_t chunks = d->chunks ? d->chunks : 4u;
    if ((uint64_t)chunks > n) chunks = (uint32_t)(n ? n : 1);
    P->starts.resize(chunks + 1);
    for (uint32_t k = 0; k <= chunks; ++k) P->starts[k] = (n * k) / chunks;
    DevMem d_starts;
    RUVGNN_CUDA_TRY(d_starts.alloc((chunks + 1) * 8));
    RUVGNN_CUDA_TRY(cudaMemcpyAsync(d_starts.p, P->starts.data(), (chunks + 1) * 8, cudaMemcpyHostToDevice, s));
    if (n + n_total > 0xffffffffull)
      return fail(RUVGNN_ERR_INVALID_INPUT, "local + gathered rows exceed u32 column indices");
    if (E) {
      remap_dense_kernel<<<grid_1d(E), 256, 0, s>>>(G.col_idx, E, lo, hi, n, (uint32_t)world,
                                                    d_starts.as<uint64_t>(), chunks);
      ++c->launches;
    }
    G.n_cols = n + n_total;
    RUVGNN_CUDA_TRY(cudaStreamSynchronize(s));
  } else {
    // halo slots in id order (grouped by owner), then the peer-local rows to request
    RUVGNN_CUDA_TRY(cudaMalloc(&P->need_ids, (n_halo + 4) * sizeof(uint32_t)));
    need_ids_kernel<<<grid_1d(n_words), 256, 0, s>>>(bits.as<unsigned long long>(), n_words,
                                                     pre.as<uint32_t>(), P->need_ids);
    if (E) remap_sparse_kernel<<<grid_1d(E), 256, 0, s>>>(G.col_idx, E, lo, hi,
                                                          bits.as<unsigned long long>(), pre.as<uint32_t>());
    c->launches += 2;
    G.n_cols = n + n_halo;
    RUVGNN_CUDA_TRY(cudaMalloc(&P->send_idx, (n_send + 4) * sizeof(uint32_t)));
    if (world > 1 && !d->offline) {
      NcclApi *a = nccl();
      DevMem req;  // need_ids relative to their owner's first row
      RUVGNN_CUDA_TRY(req.alloc((n_halo + 4) * sizeof(uint32_t)));
      RUVGNN_CUDA_TRY(cudaMemcpyAsync(req.p, P->need_ids, n_halo * sizeof(uint32_t), cudaMemcpyDeviceToDevice, s));
      if (n_halo) {
        sub_offset_kernel<<<grid_1d(n_halo), 256, 0, s>>>(req.as<uint32_t>(), n_halo, d_oe.as<uint64_t>(),
                                                          d_bounds.as<uint64_t>(), (uint32_t)world);
        ++c->launches;
      }
      NCCL_TRY(a->GroupStart());
      uint64_t soff = 0;
      for (int p = 0; p < world; ++p) {
        if (p == rank) continue;
        if (P->recv_counts[p]) NCCL_TRY(a->Send(req.as<uint32_t>() + oe[p], P->recv_counts[p], NCCL_UINT32, p, d->comm, s));
      }
      for (int p = 0; p < world; ++p) {
        if (p != rank && P->send_counts[p])
          NCCL_TRY(a->Recv(P->send_idx + soff, P->send_counts[p], NCCL_UINT32, p, d->comm, s));
        soff += P->send_counts[p];
      }
      NCCL_TRY(a->GroupEnd());
      RUVGNN_CUDA_TRY(cudaMemsetAsync(c->d_flag, 0, sizeof(int), s));
      if (n_send) check_send_idx_kernel<<<grid_1d(n_send), 256, 0, s>>>(P->send_idx, n_send, n, c->d_flag);
      RUVGNN_CUDA_TRY(cudaMemcpyAsync(&flag, c->d_flag, sizeof(int), cudaMemcpyDeviceToHost, s));
      RUVGNN_CUDA_TRY(cudaStreamSynchronize(s));
      if (flag) return fail(RUVGNN_ERR_INVALID_INPUT, "a peer requested a row this rank does not own");
    }
    RUVGNN_CUDA_TRY(cudaStreamSynchronize(s));
  }
  P->n_send = dense ? 0 : n_send;

  // ---- interior runs (exact mode runs them under the exchange) ---------------------------------
  if (n) {
    DevMem flags_d;
    RUVGNN_CUDA_TRY(flags_d.alloc(n));
    interior_flag_kernel<<<grid_1d(n * 32), 256, 0, s>>>(G.row_ptr, G.col_idx, n, flags_d.as<uint8_t>());
    ++c->launches;
    std::vector<uint8_t> is_int(n);
    RUVGNN_CUDA_TRY(cudaMemcpyAsync(is_int.data(), flags_d.p, n, cudaMemcpyDeviceToHost, s));
    RUVGNN_CUDA_TRY(cudaStreamSynchronize(s));
    interior_runs(is_int, 1024, 4, P->interior, P->boundary);
    for (const auto &r : P->interior) P->interior_rows += r.second - r.first;
  }
  // ---- pass-ordered edge arrays for the overlapped forward ---------------------------------------
  if (world > 1 && n && !d->offline) {
    std::vector<uint64_t> pb;  // column bounds of the passes: own rows, then one range per chunk
    pb.push_back(0);
    pb.push_back(n);
    if (dense)
      for (size_t k = 1; k < P->starts.size(); ++k) pb.push_back(n + (uint64_t)world * P->starts[k]);
    else
      pb.push_back(n + n_halo);
    int32_t rc = build_pass_csr(c, &G, pb, &P->pass);
    if (rc) return rc;
  }
  RUVGNN_CUDA_TRY(cudaGetLastError());
  *out = P.release();
  return RUVGNN_OK;
  RUVGNN_ABI_CATCH
}

int32_t ruvgnn_partition_stats(const ruvgnn_partition *p, ruvgnn_partition_info *out) {
  RUVGNN_ABI_TRY
  if (!p || !out) return fail(RUVGNN_ERR_INVALID_INPUT, "null argument");
  out->n_local = p->n_local;
  out->n_total = p->n_total;
  out->n_edges = p->n_edges;
  out->n_halo = p->n_halo;
  out->n_send = p->n_send;
  out->remote_edges = p->remote_edges;
  out->interior_rows = p->interior_rows;
  out->dense = p->dense ? 1 : 0;
  out->chunks = p->dense ? (uint32_t)(p->starts.size() - 1) : 1;
  return RUVGNN_OK;
  RUVGNN_ABI_CATCH
}

int32_t ruvgnn_partition_download(const ruvgnn_partition *p, uint32_t *col_idx_local,
                                  uint32_t *need_ids, uint32_t *send_idx, uint64_t *recv_counts,
                                  uint64_t *send_counts) {
  RUVGNN_ABI_TRY
  if (!p) return fail(RUVGNN_ERR_INVALID_INPUT, "partition is null");
  DeviceGuard dg(p->graph.device);
  if (col_idx_local && p->n_edges)
    RUVGNN_CUDA_TRY(cudaMemcpy(col_idx_local, p->graph.col_idx, p->n_edges * sizeof(uint32_t),
                               cudaMemcpyDeviceToHost));
  if (need_ids && p->need_ids && p->n_halo)
    RUVGNN_CUDA_TRY(cudaMemcpy(need_ids, p->need_ids, p->n_halo * sizeof(uint32_t), cudaMemcpyDeviceToHost));
  if (send_idx && p->send_idx && p->n_send)
    RUVGNN_CUDA_TRY(cudaMemcpy(send_idx, p->send_idx, p->n_send * sizeof(uint32_t), cudaMemcpyDeviceToHost));
  const size_t world = p->recv_counts.size();
  if (recv_counts) memcpy(recv_counts, p->recv_counts.data(), world * 8);
  if (send_counts) memcpy(send_counts, p->send_counts.data(), world * 8);
  return RUVGNN_OK;
  RUVGNN_ABI_CATCH
}

int32_t ruvgnn_dist_forward(ruvgnn_dist *d, const ruvgnn_partition *P, const ruvgnn_layer *L,
                            const float *x, float *out, uint32_t flags) {
  RUVGNN_ABI_TRY
  if (!d || !P || !L || !x || !out) return fail(RUVGNN_ERR_INVALID_INPUT, "null argument");
  if (P->d != d) return fail(RUVGNN_ERR_INVALID_INPUT, "partition belongs to another dist");
  ruvgnn_ctx *c = d->ctx;
  if (d->offline) return fail(RUVGNN_ERR_NCCL, "this dist was created without an NCCL id (plan-only)");
  if (L->device != c->device) return fail(RUVGNN_ERR_INVALID_INPUT, "layer lives on another device");
  const uint64_t n = P->n_local;
  if (n == 0 && d->world == 1) return RUVGNN_OK;
  DeviceGuard dg(c->device);
  if (!dg.ok) return fail(RUVGNN_ERR_CUDA, "cudaSetDevice failed");
  NcclApi *a = nccl();
  const int world = d->world, rank = d->rank;
  const uint32_t H = L->H, D = L->D, W = L->r1 ? H : 3 * H;
  const bool dev_ptrs = flags & RUVGNN_FWD_DEVICE_PTRS;
  const bool mp_ok = L->fast && (!L->r1 || L->heads <= 4);
  const bool exact = !(flags & RUVGNN_DIST_OVERLAP) || (flags & RUVGNN_DIST_EXACT) || !mp_ok || world == 1 || !P->pass.col;
  const ruvgnn_graph *G = &P->graph;
  cudaStream_t s = c->stream, cs = d->comm_stream;
  int32_t rc;
  if (dev_ptrs && x == out) return fail(RUVGNN_ERR_INVALID_INPUT, "out must not alias x (Jacobi update)");

  // Peer-direct plane (opt-in, RUVGNN_DIST_PEER_DIRECT): on a sparse plan the gather reads the remote
  // rows straight from their owners' tables over NVLink -- single pass, stored order, so the bits of
  // the single-GPU run -- instead of packing, exchanging and re-reading a halo buffer.  Measured
  // (2 B200s, config-2 shape, 5% remote edges): the NVLink reads run at ~700 GB/s but add to the gather
  // time instead of hiding under it, and they are not de-duplicated, so the pack + all-to-all-v path
  // wins whenever a remote row is referenced more than once or twice per rank.
  const bool direct = world > 1 && d->ipc_capable && !P->dense && P->col_global && mp_ok &&
                      (flags & RUVGNN_DIST_PEER_DIRECT);
  if (direct && d->free_pending) {  // peers may still be reading the table the projection overwrites
    RUVGNN_CUDA_TRY(cudaStreamWaitEvent(s, d->ev_free, 0));
    d->free_pending = false;
  }
  // ---- buffers --------------------------------------------------------------------------------
  c->state_per_node = !exact && !direct;
  c->peer_dev = nullptr;
  if ((rc = ensure_ws(c, L, n))) return rc;
  const float *x_dev = x;
  float *out_dev = out;
  if (!dev_ptrs) {
    if ((rc = ensure_stage(c, &c->x_stage, &c->x_stage_floats, (size_t)n * D))) return rc;
    if ((rc = ensure_stage(c, &c->out_stage, &c->out_stage_floats, (size_t)n * H))) return rc;
    x_dev = c->x_stage;
    out_dev = c->out_stage;
  }
  const uint64_t halo_rows = direct ? 0 : (P->dense ? P->n_total : P->n_halo);
  // Message-row halo for K|V|M layers (dense plane, default): peers ship M = W_msg x + b only (a third
  // of the bytes) and K, V are rebuilt where they arrive with the projection's own GEMM
  // (layer.rs:136-137; the bits of the owner's projection), chunk by chunk while the next chunk is on
  // the wire.  Measured at 8 GPUs, config-2 shape: 100.1 ms per step against 107.4 with K|V|M rows.
  const bool msg_halo = !(flags & RUVGNN_DIST_ROW_HALO) && P->dense && !L->r1 && world > 1;
  if (world > 1 && !direct) {
    if ((rc = ensure_buf(c, &d->halo, &d->halo_floats, (size_t)halo_rows * W + 4))) return rc;
    const uint64_t send_rows = P->dense ? (L->r1 ? 0 : n) : P->n_send;  // message rows go out in place
    if ((rc = ensure_buf(c, &d->send, &d->send_floats, (size_t)send_rows * (msg_halo ? H : W) + 4))) return rc;
    if (msg_halo && (rc = ensure_buf(c, &d->mstage, &d->mstage_floats, (size_t)P->n_total * H + 4))) return rc;
  }
  c->timing_valid = c->timing_on;
  if (c->timing_on) {
    begin_timing_slot(c);
    cudaEventRecord(c->ev[0], s);
  }

  // ---- project (host buffers: row blocks are projected as their H2D copies land) --------------
  // Message-row halos leave as soon as the message GEMM is done; the query GEMM runs beside the
  // exchange.  K|V|M halos need the whole projection.
  const bool early = (L->r1 || msg_halo) && world > 1 && dev_ptrs && !direct;  // the halo rows are message rows
  if (dev_ptrs) {
    if (c->timing_on) cudaEventRecord(c->ev[1], s);
    if ((rc = do_project_rows(c, L, n, x_dev, 0, n, early ? 1 : 3))) return rc;
  } else {
    const uint64_t target = 131072;
    uint64_t nb = (n + target - 1) / target;
    if (nb > (uint64_t)ruvgnn_ctx::kMaxBlocks) nb = ruvgnn_ctx::kMaxBlocks;
    if (nb < 1) nb = 1;
    const uint64_t per = ((n + nb - 1) / nb + 127) / 128 * 128;
    RUVGNN_CUDA_TRY(cudaEventRecord(c->ev_start, s));
    RUVGNN_CUDA_TRY(cudaStreamWaitEvent(c->s_in, c->ev_start, 0));
    RUVGNN_CUDA_TRY(cudaStreamWaitEvent(c->s_out, c->ev_start, 0));
    uint64_t used = 0;
    for (uint64_t b = 0; b < nb; ++b) {
      const uint64_t rb = b * per, re = rb + per < n ? rb + per : n;
      if (rb >= n) break;
      ++used;
      RUVGNN_CUDA_TRY(cudaMemcpyAsync(c->x_stage + rb * D, x + rb * D, (re - rb) * D * sizeof(float),
                                      cudaMemcpyHostToDevice, c->s_in));
      RUVGNN_CUDA_TRY(cudaEventRecord(c->ev_in[b], c->s_in));
    }
    for (uint64_t b = 0; b < used; ++b) {
      const uint64_t rb = b * per, re = rb + per < n ? rb + per : n;
      RUVGNN_CUDA_TRY(cudaStreamWaitEvent(s, c->ev_in[b], 0));
      if (c->timing_on && b == used - 1) cudaEventRecord(c->ev[1], s);
      if ((rc = do_project_rows(c, L, n, x_dev, rb, re, 3))) return rc;
    }
    if (n == 0) invalidate_projection(c);
  }
  Ws w = carve(c, L, n);

  // ---- exchange, on the communication stream --------------------------------------------------
  const uint32_t n_chunks = (world == 1 || direct) ? 0 : (P->dense ? (uint32_t)(P->starts.size() - 1) : 1);
  double bytes_sent = 0, bytes_recv = 0;
  if (direct) {
    // every rank publishes the IPC handle of its workspace (the projected table sits at its start);
    // the all-gather is also the barrier "every rank has finished projecting"
    RUVGNN_CUDA_TRY(cudaEventRecord(d->ev_ready, s));
    RUVGNN_CUDA_TRY(cudaStreamWaitEvent(cs, d->ev_ready, 0));
    RUVGNN_CUDA_TRY(cudaEventRecord(d->ev_x0, cs));
    unsigned char rec[kXchgBytes] = {};
    cudaIpcMemHandle_t h;
    RUVGNN_CUDA_TRY(cudaIpcGetMemHandle(&h, c->ws));
    memcpy(rec, &h, sizeof h);
    rec[64] = 1;
    std::vector<unsigned char> all;
    if ((rc = exchange_records(d, rec, &all))) return rc;
    PeerTable t;
    t.world = world;
    t.rank = rank;
    for (int p = 0; p <= world; ++p) t.bounds[p] = P->bounds[p];
    for (int p = 0; p < world; ++p) {
      if (p == rank) {
        t.base[p] = c->ws;
        continue;
      }
      cudaIpcMemHandle_t ph;
      memcpy(&ph, &all[(size_t)p * kXchgBytes], sizeof ph);
      cudaError_t e = map_peer(d, p, ph);
      if (e != cudaSuccess)
        return fail(RUVGNN_ERR_CUDA, std::string("cudaIpcOpenMemHandle (peer workspace): ") + cudaGetErrorString(e));
      t.base[p] = static_cast<const float *>(d->peers[p].ptr);
    }
    RUVGNN_CUDA_TRY(cudaMemcpyAsync(d->peer_table_dev, &t, sizeof t, cudaMemcpyHostToDevice, cs));
    RUVGNN_CUDA_TRY(cudaEventRecord(d->ev_x1, cs));
    RUVGNN_CUDA_TRY(cudaEventRecord(d->ev_xchg, cs));
    RUVGNN_CUDA_TRY(cudaStreamWaitEvent(s, d->ev_xchg, 0));
    c->peer_dev = d->peer_table_dev;
    c->peer_cols = P->col_global;
    bytes_recv = (double)P->remote_edges * W * 4;  // pulled by the gather's own loads, no de-duplication
  } else if (world > 1) {
    const float *src = w.halo_src;
    const uint32_t Wx = msg_halo ? H : W;  // floats per row on the wire
    float *land = msg_halo ? d->mstage : d->halo;
    if (msg_halo && n) {  // message rows of every own row, made contiguous
      RUVGNN_CUDA_TRY(cudaMemcpy2DAsync(d->send, (size_t)H * 4, w.m, (size_t)w.ldm * 4, (size_t)H * 4, n,
                                        cudaMemcpyDeviceToDevice, s));
      src = d->send;
    } else if (P->dense && !L->r1 && n) {  // K|V|M of every own row, made contiguous
      RUVGNN_CUDA_TRY(cudaMemcpy2DAsync(d->send, (size_t)W * 4, w.halo_src, (size_t)w.ldm * 4,
                                        (size_t)W * 4, n, cudaMemcpyDeviceToDevice, s));
      src = d->send;
    } else if (!P->dense && P->n_send) {
      RUVGNN_CUDA_TRY(launch_pack_rows(w.halo_src, w.ldm, W, P->send_idx, P->n_send, d->send, s,
                                       &c->launches));
    }
    RUVGNN_CUDA_TRY(cudaEventRecord(d->ev_ready, s));
    RUVGNN_CUDA_TRY(cudaStreamWaitEvent(cs, d->ev_ready, 0));
    RUVGNN_CUDA_TRY(cudaEventRecord(d->ev_x0, cs));
    if (P->dense) {
      for (uint32_t k = 0; k < n_chunks; ++k) {
        const uint64_t a0 = P->starts[k], a1 = P->starts[k + 1];
        if (a1 > a0)
          NCCL_TRY(a->AllGather(src + a0 * Wx, land + (uint64_t)world * a0 * Wx, (a1 - a0) * Wx,
                                NCCL_FLOAT32, d->comm, cs));
        RUVGNN_CUDA_TRY(cudaEventRecord(d->ev_chunk[k], cs));
      }
      bytes_sent = (double)n * Wx * 4 * (world - 1);
      bytes_recv = (double)(P->n_total - n) * Wx * 4;
    } else {
      NCCL_TRY(a->GroupStart());
      uint64_t soff = 0, roff = 0;
      for (int p = 0; p < world; ++p) {
        if (p != rank && P->send_counts[p])
          NCCL_TRY(a->Send(d->send + soff * W, P->send_counts[p] * W, NCCL_FLOAT32, p, d->comm, cs));
        if (p != rank && P->recv_counts[p])
          NCCL_TRY(a->Recv(d->halo + roff * W, P->recv_counts[p] * W, NCCL_FLOAT32, p, d->comm, cs));
        soff += P->send_counts[p];
        roff += P->recv_counts[p];
      }
      NCCL_TRY(a->GroupEnd());
      RUVGNN_CUDA_TRY(cudaEventRecord(d->ev_chunk[0], cs));
      bytes_sent = (double)P->n_send * W * 4;
      bytes_recv = (double)P->n_halo * W * 4;
    }
    RUVGNN_CUDA_TRY(cudaEventRecord(d->ev_x1, cs));
  }
  if (early && (rc = do_project_rows(c, L, n, x_dev, 0, n, 2))) return rc;
  if (c->timing_on) cudaEventRecord(c->ev[2], s);

  // host buffers: every finished row block of the output goes down on the copy-out stream
  uint32_t out_blocks = 0;
  auto finish_rows = [&](uint64_t rb, uint64_t re) -> int32_t {
    if (dev_ptrs || re <= rb) return RUVGNN_OK;
    cudaEvent_t ev = c->ev_out[out_blocks % ruvgnn_ctx::kMaxBlocks];
    ++out_blocks;
    RUVGNN_CUDA_TRY(cudaEventRecord(ev, s));
    RUVGNN_CUDA_TRY(cudaStreamWaitEvent(c->s_out, ev, 0));
    RUVGNN_CUDA_TRY(cudaMemcpyAsync(out + rb * H, c->out_stage + rb * H, (re - rb) * H * sizeof(float),
                                    cudaMemcpyDeviceToHost, c->s_out));
    return RUVGNN_OK;
  };
  // K-agg (final pass) + K-upd block by block, handing finished blocks to the copy-out stream
  auto finish_range = [&](uint64_t rb, uint64_t re, const GatherPass &pass) -> int32_t {
    if (dev_ptrs) return do_aggregate_update(c, L, G, d->halo, rb, re, out_dev, true, pass);  // row-block pipeline inside
    const uint64_t B = std::min<uint64_t>(upd_block(), 131072);
    for (uint64_t b0 = rb; b0 < re; b0 += B) {
      const uint64_t b1 = std::min(b0 + B, re);
      int32_t r = gather_rows(c, L, G, d->halo, b0, b1, pass, true);
      if (!r) r = update_rows(c, L, G, b0, b1, out_dev);
      if (!r) r = finish_rows(b0, b1);
      if (r) return r;
    }
    return RUVGNN_OK;
  };
  d->waits_used = 0;
  // message-row halo: K | V | M of chunk k's rows, rebuilt from the gathered message rows (own block skipped)
  auto rebuild_chunk = [&](uint32_t k) -> int32_t {
    if (!msg_halo) return RUVGNN_OK;
    const uint64_t a0 = P->starts[k], sz = P->starts[k + 1] - a0;
    const uint64_t base = (uint64_t)world * a0;
    const uint64_t lo[2] = {0, (uint64_t)(rank + 1) * sz}, hi[2] = {(uint64_t)rank * sz, (uint64_t)world * sz};
    for (int part = 0; part < 2; ++part)
      if (hi[part] > lo[part]) {
        int32_t r = ruvgnn_layer_expand_halo(c, L, d->mstage + (base + lo[part]) * H, hi[part] - lo[part],
                                             d->halo + (base + lo[part]) * W);
        if (r) return r;
      }
    return RUVGNN_OK;
  };
  auto wait_chunk = [&](uint32_t k) -> int32_t {
    const uint32_t i = d->waits_used++;
    RUVGNN_CUDA_TRY(cudaEventRecord(d->ev_w0[i], s));
    RUVGNN_CUDA_TRY(cudaStreamWaitEvent(s, d->ev_chunk[k], 0));
    RUVGNN_CUDA_TRY(cudaEventRecord(d->ev_w1[i], s));
    return rebuild_chunk(k);
  };

  // RUVGNN_COMM_SMS=k: size the grids for k fewer SMs while the exchange is in flight (experiments)
  static int comm_sms = -1;
  if (comm_sms < 0) {
    const char *e = getenv("RUVGNN_COMM_SMS");
    comm_sms = e ? atoi(e) : 0;
    if (comm_sms < 0 || comm_sms >= c->sm_count_dev) comm_sms = 0;
  }
  const int sm_full = c->sm_count;
  if (world > 1 && comm_sms > 0) c->sm_count = c->sm_count_dev - comm_sms;
  uint32_t passes = 1;
  if (direct) {
    if ((rc = finish_range(0, n, GatherPass()))) return rc;
    c->peer_dev = nullptr;
    // "every rank has finished reading": a tiny all-gather on the communication stream; the next
    // forward's projection (which overwrites the table) waits for it
    RUVGNN_CUDA_TRY(cudaEventRecord(d->ev_gdone, s));
    RUVGNN_CUDA_TRY(cudaStreamWaitEvent(cs, d->ev_gdone, 0));
    NCCL_TRY(a->AllGather(d->xchg_dev + (size_t)world * kXchgBytes, d->xchg_dev, 1, NCCL_UINT8, d->comm, cs));
    RUVGNN_CUDA_TRY(cudaEventRecord(d->ev_free, cs));
    d->free_pending = true;
  } else if (exact) {
    // all-local row runs first (they need no halo), then everything else after the halo landed
    if (world > 1)
      for (const auto &r : P->interior)
        if ((rc = finish_range(r.first, r.second, GatherPass()))) return rc;
    if (world > 1) {
      // chunk by chunk: K, V of chunk k are rebuilt (message-row halo) while chunk k+1 is on the wire
      for (uint32_t k = 0; k < n_chunks; ++k)
        if ((rc = wait_chunk(k))) return rc;
      for (const auto &r : P->boundary)
        if ((rc = finish_range(r.first, r.second, GatherPass()))) return rc;
    } else if ((rc = finish_range(0, n, GatherPass()))) {
      return rc;
    }
  } else {
    // pass 0: the edges to own rows, for all rows, while the halo is on the wire
    GatherPass p0;
    p0.csr = &P->pass;
    p0.index = 0;
    p0.first = true;
    p0.last = false;
    if (n && (rc = gather_rows(c, L, G, d->halo, 0, n, p0, true))) return rc;
    passes = 1 + n_chunks;
    for (uint32_t k = 0; k < n_chunks; ++k) {
      if ((rc = wait_chunk(k))) return rc;
      GatherPass pk;
      pk.csr = &P->pass;
      pk.index = (int)k + 1;
      pk.first = false;
      pk.last = k + 1 == n_chunks;
      if (pk.last) c->sm_count = sm_full;  // the exchange is over
      if (!pk.last) {
        if (n && (rc = gather_rows(c, L, G, d->halo, 0, n, pk, true))) return rc;
      } else if ((rc = finish_range(0, n, pk))) {
        return rc;
      }
    }
  }
  c->sm_count = sm_full;
  if (c->timing_on) cudaEventRecord(c->ev[4], s);
  if (!dev_ptrs) {
    RUVGNN_CUDA_TRY(cudaEventRecord(c->ev_done, c->s_out));
    RUVGNN_CUDA_TRY(cudaStreamWaitEvent(s, c->ev_done, 0));
  }
  if (c->timing_on) cudaEventRecord(c->ev[5], s);
  c->state_per_node = false;
  d->stats.bytes_sent = bytes_sent;
  d->stats.bytes_received = bytes_recv;
  d->stats.passes = passes;
  d->stats.exact = (exact || direct) ? 1 : 0;
  d->stats.peer_direct = direct ? 1 : 0;
  d->stats_valid = true;
  if (!(flags & RUVGNN_FWD_ASYNC) || !dev_ptrs) RUVGNN_CUDA_TRY(cudaStreamSynchronize(s));
  if (d->comm) {
    ncclResult_t async_err = 0;
    if (a->CommGetAsyncError(d->comm, &async_err) == 0 && async_err != 0 && async_err != 7 /* in progress */)
      return nccl_fail("asynchronous NCCL error", async_err);
  }
  return RUVGNN_OK;
  RUVGNN_ABI_CATCH
}

int32_t ruvgnn_dist_last_stats(ruvgnn_dist *d, ruvgnn_dist_stats *out) {
  RUVGNN_ABI_TRY
  if (!d || !out) return fail(RUVGNN_ERR_INVALID_INPUT, "null argument");
  if (!d->stats_valid) return fail(RUVGNN_ERR_INVALID_INPUT, "no forward has run on this dist");
  DeviceGuard dg(d->ctx->device);
  RUVGNN_CUDA_TRY(cudaStreamSynchronize(d->ctx->stream));
  RUVGNN_CUDA_TRY(cudaStreamSynchronize(d->comm_stream));
  float wait = 0.f;
  for (uint32_t i = 0; i < d->waits_used; ++i) {
    float t = 0.f;
    RUVGNN_CUDA_TRY(cudaEventElapsedTime(&t, d->ev_w0[i], d->ev_w1[i]));
    wait += t;
  }
  d->stats.exposed_wait_ms = wait;
  d->stats.exchange_ms = 0.f;
  if (d->world > 1) RUVGNN_CUDA_TRY(cudaEventElapsedTime(&d->stats.exchange_ms, d->ev_x0, d->ev_x1));
  *out = d->stats;
  return RUVGNN_OK;
  RUVGNN_ABI_CATCH
}

// Single-device diagnostic form of the multi-pass gather (include/ruvgnn.h).
int32_t ruvgnn_layer_forward_split(ruvgnn_ctx *c, const ruvgnn_layer *L, const ruvgnn_graph *G,
                                   const float *x_dev, float *out_dev, const uint64_t *splits,
                                   uint32_t n_splits) {
  RUVGNN_ABI_TRY
  if (!c || !L || !G || !x_dev || !out_dev || (n_splits && !splits))
    return fail(RUVGNN_ERR_INVALID_INPUT, "null argument");
  if (L->device != c->device || G->device != c->device)
    return fail(RUVGNN_ERR_INVALID_INPUT, "layer / graph live on another device than ctx");
  if (G->n_cols != G->n_nodes)
    return fail(RUVGNN_ERR_INVALID_INPUT, "graph addresses halo rows; use ruvgnn_dist_forward");
  if (!L->fast || (L->r1 && L->heads > 4))
    return fail(RUVGNN_ERR_INVALID_SHAPE, "the multi-pass gather needs hidden_dim % 128 == 0 and heads | 32");
  for (uint32_t i = 0; i < n_splits; ++i)
    if (splits[i] > G->n_cols || (i && splits[i] < splits[i - 1]))
      return fail(RUVGNN_ERR_INVALID_INPUT, "splits must be ascending column bounds");
  const uint64_t n = G->n_nodes;
  if (n == 0) return RUVGNN_OK;
  DeviceGuard dg(c->device);
  if (!dg.ok) return fail(RUVGNN_ERR_CUDA, "cudaSetDevice failed");
  c->timing_valid = false;
  std::vector<uint64_t> pb;
  pb.push_back(0);
  for (uint32_t i = 0; i < n_splits; ++i) pb.push_back(splits[i]);
  pb.push_back(G->n_cols);
  PassCsr pc;
  int32_t rc = n_splits ? build_pass_csr(c, G, pb, &pc) : RUVGNN_OK;
  c->state_per_node = true;
  if (!rc) rc = do_project(c, L, n, x_dev);
  for (uint32_t i = 0; i <= n_splits && !rc; ++i) {
    GatherPass p;
    p.csr = n_splits ? &pc : nullptr;
    p.index = (int)i;
    p.first = i == 0;
    p.last = i == n_splits;
    if (!p.last)
      rc = gather_rows(c, L, G, nullptr, 0, n, p, false);
    else
      rc = do_aggregate_update(c, L, G, nullptr, 0, n, out_dev, false, p);
  }
  cudaStreamSynchronize(c->stream);
  if (pc.col) cudaFree(pc.col);
  if (pc.w) cudaFree(pc.w);
  if (pc.seg) cudaFree(pc.seg);
  c->state_per_node = false;
  if (!rc) RUVGNN_CUDA_TRY(cudaStreamSynchronize(c->stream));
  invalidate_projection(c);
  return rc;
  RUVGNN_ABI_CATCH
}

}  // extern "C"
