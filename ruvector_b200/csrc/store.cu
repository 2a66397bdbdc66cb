// Embedding store files as the X source and the output sink of the batched forward:
// the rows stream file -> pinned host block -> device block -> projection, and the new embeddings
// stream device block -> pinned host block -> file, so neither X nor the output is ever resident
// as a whole, on the host or in HBM.
//
// Layouts of the reference that are read and written in place, without conversion:
//   MmapManager  (crates/ruvector-gnn/src/mmap.rs:118-186): headerless dense row-major
//                [max_nodes][d_embed] f32, node i at byte i * d_embed * 4 (:199-203)
//   FeatureStorage (crates/ruvector-gnn/src/cold_tier.rs:17-18,45-75): 24-byte little-endian header
//                dim u64 | num_nodes u64 | block_size u64, then one page-aligned block of block_size
//                bytes per node whose first dim * 4 bytes are the features (:120-149)
// Host side: plain pread / pwrite on the caller's files (no mmap: a 100M x 128 store is 51 GB and is
// touched exactly once per layer, front to back).
#include <errno.h>
#include <fcntl.h>
#include <string.h>
#include <sys/stat.h>
#include <unistd.h>

#include <string>

#include "abi_guard.h"
#include "engine.h"

using namespace ruvgnn;

namespace {

struct Fd {
  int fd = -1;
  ~Fd() {
    if (fd >= 0) close(fd);
  }
};

struct Pinned {
  void *p = nullptr;
  ~Pinned() {
    if (p) cudaFreeHost(p);
  }
};

struct DevBlock {
  void *p = nullptr;
  ~DevBlock() {
    if (p) cudaFree(p);
  }
};

struct Ev {
  cudaEvent_t e = nullptr;
  ~Ev() {
    if (e) cudaEventDestroy(e);
  }
};

int32_t io_fail(const char *what, const char *path) {
  return fail(RUVGNN_ERR_INVALID_INPUT, std::string(what) + " " + (path ? path : "(null)") + ": " + strerror(errno));
}

bool read_all(int fd, void *buf, size_t n, uint64_t off) {
  char *p = static_cast<char *>(buf);
  while (n) {
    const ssize_t r = pread(fd, p, n, (off_t)off);
    if (r < 0 && errno == EINTR) continue;
    if (r <= 0) return false;
    p += r;
    off += (uint64_t)r;
    n -= (size_t)r;
  }
  return true;
}

bool write_all(int fd, const void *buf, size_t n, uint64_t off) {
  const char *p = static_cast<const char *>(buf);
  while (n) {
    const ssize_t r = pwrite(fd, p, n, (off_t)off);
    if (r < 0 && errno == EINTR) continue;
    if (r <= 0) return false;
    p += r;
    off += (uint64_t)r;
    n -= (size_t)r;
  }
  return true;
}

int32_t check_store(const ruvgnn_store *s, uint32_t dim, uint64_t rows, const char *what) {
  if (!s->path) return fail(RUVGNN_ERR_INVALID_INPUT, std::string(what) + " store has no path");
  if (s->dim != dim) {
    char buf[128];
    snprintf(buf, sizeof buf, "Tensor dimension mismatch: expected %u, got %u (%s store)", dim, s->dim, what);
    return fail(RUVGNN_ERR_DIMENSION_MISMATCH, buf);
  }
  if (s->row_stride_bytes < (uint64_t)dim * 4 || s->row_stride_bytes % 4)
    return fail(RUVGNN_ERR_INVALID_INPUT, std::string(what) + " store: row stride below dim * 4 or not a multiple of 4");
  if (s->rows < rows) return fail(RUVGNN_ERR_INVALID_INPUT, std::string(what) + " store holds fewer rows than the graph has nodes");
  return RUVGNN_OK;
}

}  // namespace

extern "C" {

int32_t ruvgnn_store_describe_mmap(const char *path, uint32_t d_embed, uint64_t max_nodes,
                                   ruvgnn_store *out) {
  RUVGNN_ABI_TRY
  if (!path || !out) return fail(RUVGNN_ERR_INVALID_INPUT, "null argument");
  if (d_embed == 0) return fail(RUVGNN_ERR_INVALID_INPUT, "d_embed must be > 0");
  struct stat st;
  if (stat(path, &st) != 0) return io_fail("cannot stat", path);
  if ((uint64_t)st.st_size < max_nodes * (uint64_t)d_embed * 4)
    return fail(RUVGNN_ERR_INVALID_INPUT, "mmap file holds fewer than max_nodes x d_embed f32 values");
  out->path = path;
  out->offset_bytes = 0;
  out->row_stride_bytes = (uint64_t)d_embed * 4;
  out->dim = d_embed;
  out->rows = max_nodes;
  return RUVGNN_OK;
  RUVGNN_ABI_CATCH
}

int32_t ruvgnn_store_describe_feature_storage(const char *path, ruvgnn_store *out) {
  RUVGNN_ABI_TRY
  if (!path || !out) return fail(RUVGNN_ERR_INVALID_INPUT, "null argument");
  Fd f;
  f.fd = open(path, O_RDONLY);
  if (f.fd < 0) return io_fail("cannot open", path);
  unsigned char head[24];
  if (!read_all(f.fd, head, sizeof head, 0))
    return fail(RUVGNN_ERR_INVALID_INPUT, "feature storage file is shorter than its header");
  uint64_t v[3];
  for (int i = 0; i < 3; ++i) {  // little-endian on disk (cold_tier.rs:61-63)
    v[i] = 0;
    for (int b = 7; b >= 0; --b) v[i] = (v[i] << 8) | head[8 * i + b];
  }
  const uint64_t dim = v[0], num_nodes = v[1], block = v[2];
  if (dim == 0 || dim > 0xffffffffull || block < dim * 4 || block % 4)
    return fail(RUVGNN_ERR_INVALID_INPUT, "feature storage header is inconsistent");
  struct stat st;
  if (fstat(f.fd, &st) != 0) return io_fail("cannot stat", path);
  if (num_nodes && ((uint64_t)st.st_size < 24 || ((uint64_t)st.st_size - 24) / block < num_nodes))
    return fail(RUVGNN_ERR_INVALID_INPUT, "feature storage file is truncated");
  out->path = path;
  out->offset_bytes = 24;
  out->row_stride_bytes = block;
  out->dim = (uint32_t)dim;
  out->rows = num_nodes;
  return RUVGNN_OK;
  RUVGNN_ABI_CATCH
}

int32_t ruvgnn_layer_forward_store(ruvgnn_ctx *c, const ruvgnn_layer *L, const ruvgnn_graph *G,
                                   const ruvgnn_store *xs, const ruvgnn_store *os,
                                   uint64_t block_rows) {
  RUVGNN_ABI_TRY
  if (!c || !L || !G || !xs || !os) return fail(RUVGNN_ERR_INVALID_INPUT, "null argument");
  if (L->device != c->device || G->device != c->device)
    return fail(RUVGNN_ERR_INVALID_INPUT, "layer / graph live on another device than ctx");
  if (G->n_cols != G->n_nodes)
    return fail(RUVGNN_ERR_INVALID_INPUT, "graph addresses halo rows; use ruvgnn_dist_forward");
  const uint64_t n = G->n_nodes;
  const uint32_t D = L->D, H = L->H;
  int32_t rc;
  if ((rc = check_store(xs, D, n, "input"))) return rc;
  if ((rc = check_store(os, H, n, "output"))) return rc;
  if (n == 0) return RUVGNN_OK;
  if (block_rows == 0) block_rows = 131072;
  if (block_rows > upd_block()) block_rows = upd_block();
  if (block_rows > n) block_rows = n;
  Fd fx, fo;
  fx.fd = open(xs->path, O_RDONLY);
  if (fx.fd < 0) return io_fail("cannot open", xs->path);
  fo.fd = open(os->path, O_WRONLY);
  if (fo.fd < 0) return io_fail("cannot open for writing", os->path);
  DeviceGuard dg(c->device);
  if (!dg.ok) return fail(RUVGNN_ERR_CUDA, "cudaSetDevice failed");
  c->timing_valid = false;
  c->state_per_node = false;
  if ((rc = ensure_ws(c, L, n))) return rc;

  // two pinned blocks each way (file rows keep their on-disk stride in the pinned block; the 2D
  // copies to and from the device pack / unpack them), two device blocks each way
  const size_t xin_bytes = (size_t)block_rows * xs->row_stride_bytes;
  const size_t oout_bytes = (size_t)block_rows * os->row_stride_bytes;
  Pinned hx[2], ho[2];
  DevBlock dx[2], dout[2];
  Ev in_done[2], proj_done[2], out_ready[2], out_copied[2];
  for (int i = 0; i < 2; ++i) {
    RUVGNN_CUDA_TRY(cudaMallocHost(&hx[i].p, xin_bytes));
    RUVGNN_CUDA_TRY(cudaMallocHost(&ho[i].p, oout_bytes));
    memset(ho[i].p, 0, oout_bytes);  // the padding of page-strided output blocks is written as zeros
    RUVGNN_CUDA_TRY(cudaMalloc(&dx[i].p, (size_t)block_rows * D * 4));
    RUVGNN_CUDA_TRY(cudaMalloc(&dout[i].p, (size_t)block_rows * H * 4));
    RUVGNN_CUDA_TRY(cudaEventCreateWithFlags(&in_done[i].e, cudaEventDisableTiming));
    RUVGNN_CUDA_TRY(cudaEventCreateWithFlags(&proj_done[i].e, cudaEventDisableTiming));
    RUVGNN_CUDA_TRY(cudaEventCreateWithFlags(&out_ready[i].e, cudaEventDisableTiming));
    RUVGNN_CUDA_TRY(cudaEventCreateWithFlags(&out_copied[i].e, cudaEventDisableTiming));
  }
  cudaStream_t s = c->stream;
  RUVGNN_CUDA_TRY(cudaEventRecord(c->ev_start, s));
  RUVGNN_CUDA_TRY(cudaStreamWaitEvent(c->s_in, c->ev_start, 0));
  RUVGNN_CUDA_TRY(cudaStreamWaitEvent(c->s_out, c->ev_start, 0));

  // ---- phase A: file -> pinned -> device block -> projection, block b+1 read while b is in flight
  const uint64_t nb = (n + block_rows - 1) / block_rows;
  for (uint64_t b = 0; b < nb; ++b) {
    const int k = (int)(b & 1);
    const uint64_t rb = b * block_rows, re = rb + block_rows < n ? rb + block_rows : n;
    if (b >= 2) {  // the pinned block is free once its H2D copy is done, the device block once projected
      RUVGNN_CUDA_TRY(cudaEventSynchronize(in_done[k].e));
      RUVGNN_CUDA_TRY(cudaStreamWaitEvent(c->s_in, proj_done[k].e, 0));
    }
    const size_t span = (size_t)(re - rb - 1) * xs->row_stride_bytes + (size_t)D * 4;
    if (!read_all(fx.fd, hx[k].p, span, xs->offset_bytes + rb * xs->row_stride_bytes))
      return io_fail("short read from", xs->path);
    RUVGNN_CUDA_TRY(cudaMemcpy2DAsync(dx[k].p, (size_t)D * 4, hx[k].p, xs->row_stride_bytes, (size_t)D * 4,
                                      re - rb, cudaMemcpyHostToDevice, c->s_in));
    RUVGNN_CUDA_TRY(cudaEventRecord(in_done[k].e, c->s_in));
    RUVGNN_CUDA_TRY(cudaStreamWaitEvent(s, in_done[k].e, 0));
    // do_project_rows indexes X by the node's row: shift the block's base accordingly
    const float *x_base = static_cast<const float *>(dx[k].p) - rb * D;
    if ((rc = do_project_rows(c, L, n, x_base, rb, re))) return rc;
    RUVGNN_CUDA_TRY(cudaEventRecord(proj_done[k].e, s));
  }
  // ---- phase B: gather + update per block -> device block -> pinned -> file
  uint64_t pending_rb[2] = {0, 0}, pending_re[2] = {0, 0};
  bool pending[2] = {false, false};
  auto flush = [&](int k) -> int32_t {  // block k's rows are on their way to ho[k]: wait, then write them out
    if (!pending[k]) return RUVGNN_OK;
    RUVGNN_CUDA_TRY(cudaEventSynchronize(out_copied[k].e));
    const uint64_t rb = pending_rb[k], re = pending_re[k];
    const size_t span = os->row_stride_bytes == (uint64_t)H * 4
                            ? (size_t)(re - rb) * H * 4
                            : (size_t)(re - rb) * os->row_stride_bytes;
    if (!write_all(fo.fd, ho[k].p, span, os->offset_bytes + rb * os->row_stride_bytes))
      return io_fail("short write to", os->path);
    pending[k] = false;
    return RUVGNN_OK;
  };
  for (uint64_t b = 0; b < nb; ++b) {
    const int k = (int)(b & 1);
    const uint64_t rb = b * block_rows, re = rb + block_rows < n ? rb + block_rows : n;
    if ((rc = flush(k))) return rc;  // frees ho[k]; dout[k] was released by the copy that filled it
    float *o_base = static_cast<float *>(dout[k].p) - rb * H;
    if (b >= 2) RUVGNN_CUDA_TRY(cudaStreamWaitEvent(s, out_copied[k].e, 0));
    if ((rc = do_aggregate_update(c, L, G, nullptr, rb, re, o_base, false))) return rc;
    RUVGNN_CUDA_TRY(cudaEventRecord(out_ready[k].e, s));
    RUVGNN_CUDA_TRY(cudaStreamWaitEvent(c->s_out, out_ready[k].e, 0));
    RUVGNN_CUDA_TRY(cudaMemcpy2DAsync(ho[k].p, os->row_stride_bytes, dout[k].p, (size_t)H * 4, (size_t)H * 4,
                                      re - rb, cudaMemcpyDeviceToHost, c->s_out));
    RUVGNN_CUDA_TRY(cudaEventRecord(out_copied[k].e, c->s_out));
    pending[k] = true;
    pending_rb[k] = rb;
    pending_re[k] = re;
  }
  for (int k = 0; k < 2; ++k)
    if ((rc = flush((int)((nb + k) & 1)))) return rc;
  RUVGNN_CUDA_TRY(cudaStreamSynchronize(s));
  invalidate_projection(c);
  return RUVGNN_OK;
  RUVGNN_ABI_CATCH
}

}  // extern "C"
