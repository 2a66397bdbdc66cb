"""Range-partitioned multi-GPU forward with halo exchange.

Nodes are split into contiguous ranges; rank p owns rows ``[bounds[p], bounds[p+1])`` of X
and of the CSR, projects its own rows, and needs the projected rows of every *remote*
neighbour its rows reference.  The product path lives in the library (``csrc/dist.cu``, C ABI
``ruvgnn_dist_*`` / ``ruvgnn_partition_*``): NCCL communicator, halo plan built on the device,
exchange on a communication stream overlapped with the gather passes that do not need it.
:class:`Dist` and :class:`Partition` below are thin ``ctypes`` callers of it; there is no
``torch.distributed`` call on the data path (the host program only ships the 128-byte NCCL id
to the other ranks, by any transport it likes).

:class:`HaloPlan` and the NumPy helpers are the host-side *model* of the plan the library
builds on the device (de-duplicated remote columns grouped by owner, neighbour order inside
every row preserved, dense-plane column map, interior runs).  The tests check the device plan
against them, and the gloo tests exercise the same logic across processes on CPU.

The reference has no multi-device path at all; weights (< 30 MB) are
replicated, exactly like its ``Clone``-able layer (gnn_cache.rs:183-188).
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import List, Optional, Sequence

import numpy as np

from .graph import CsrGraph


def even_bounds(n_total: int, world: int) -> np.ndarray:
    """Contiguous ranges [p*N/P, (p+1)*N/P)."""
    return np.array([(n_total * p) // world for p in range(world + 1)], dtype=np.int64)


@dataclass
class HaloPlan:
    """Host-side exchange plan of one rank.

    ``graph``       local CSR with remapped columns: < n_local own row, >= n_local halo slot
    ``need_ids``    global ids of the halo rows, slot order (sorted => grouped by owner)
    ``recv_counts`` rows received from each peer (sum == n_halo)
    ``req_local``   per peer: the peer-local row indices this rank asks that peer for
    ``send_idx``    (after :meth:`set_requests`) own-local rows to send, concatenated by peer
    ``send_counts`` rows sent to each peer
    """

    rank: int
    world: int
    bounds: np.ndarray
    graph: CsrGraph
    need_ids: np.ndarray
    recv_counts: np.ndarray
    req_local: List[np.ndarray]
    send_idx: Optional[np.ndarray] = None
    send_counts: Optional[np.ndarray] = None

    @property
    def n_local(self) -> int:
        return int(self.bounds[self.rank + 1] - self.bounds[self.rank])

    @property
    def n_halo(self) -> int:
        return int(self.need_ids.size)

    @classmethod
    def build(cls, local_rows: CsrGraph, rank: int, bounds: Sequence[int]) -> "HaloPlan":
        """``local_rows``: this rank's CSR rows with GLOBAL column ids."""
        bounds = np.asarray(bounds, dtype=np.int64)
        world = bounds.size - 1
        lo, hi = int(bounds[rank]), int(bounds[rank + 1])
        if local_rows.n_nodes != hi - lo:
            raise ValueError("local_rows must hold exactly this rank's rows")
        col = local_rows.col_idx.astype(np.int64)
        if col.size and (col.min() < 0 or col.max() >= bounds[-1]):
            raise ValueError("column id outside the global node range")
        remote = (col < lo) | (col >= hi)
        need = np.unique(col[remote])  # sorted: grouped by owner, ascending id inside
        new_col = np.where(remote, 0, col - lo)
        if need.size:
            new_col[remote] = (hi - lo) + np.searchsorted(need, col[remote])
        if (hi - lo) + need.size > 0xFFFFFFFF:
            raise ValueError("local + halo rows exceed u32 column indices")
        owner_edges = np.searchsorted(need, bounds)  # need[owner_edges[p]:owner_edges[p+1]] live on p
        recv_counts = np.diff(owner_edges).astype(np.int64)
        req_local = [(need[owner_edges[p]:owner_edges[p + 1]] - bounds[p]).astype(np.int64) for p in range(world)]
        g = CsrGraph(local_rows.row_ptr.copy(), new_col.astype(np.uint32), local_rows.edge_w.copy())
        return cls(rank, world, bounds, g, need, recv_counts, req_local)

    def set_requests(self, requests_from_peers: Sequence[np.ndarray]):
        """``requests_from_peers[p]`` = own-local row indices peer p asked for."""
        n = self.n_local
        for p, r in enumerate(requests_from_peers):
            if r.size and (r.min() < 0 or r.max() >= n):
                raise ValueError(f"peer {p} requested a row this rank does not own")
        self.send_counts = np.array([r.size for r in requests_from_peers], dtype=np.int64)
        self.send_idx = (np.concatenate(requests_from_peers) if requests_from_peers else np.zeros(0)).astype(np.uint32)
        return self


def exchange_requests(plan: HaloPlan, group=None) -> HaloPlan:
    """All-to-all of the request lists (once per graph).  Works on gloo (CPU) and nccl."""
    import torch
    import torch.distributed as dist

    if plan.world == 1:
        return plan.set_requests([np.zeros(0, dtype=np.int64)])
    dev = torch.device("cuda", torch.cuda.current_device()) if dist.get_backend(group) == "nccl" else torch.device("cpu")
    out_counts = torch.tensor(plan.recv_counts, dtype=torch.int64, device=dev)  # what I ask of each peer
    in_counts = torch.empty_like(out_counts)
    dist.all_to_all_single(in_counts, out_counts, group=group)
    send = torch.from_numpy(np.concatenate(plan.req_local).astype(np.int64)).to(dev)
    recv = torch.empty(int(in_counts.sum().item()), dtype=torch.int64, device=dev)
    dist.all_to_all_single(recv, send, in_counts.tolist(), out_counts.tolist(), group=group)
    recv = recv.cpu().numpy()
    offs = np.concatenate([[0], np.cumsum(in_counts.cpu().numpy())])
    return plan.set_requests([recv[offs[p]:offs[p + 1]] for p in range(plan.world)])


def exchange_rows(plan: HaloPlan, send_rows, group=None):
    """Per-layer all-to-all-v: ``send_rows`` [n_send, W] (rows ``plan.send_idx`` of the projected
    block, already packed) -> halo [n_halo, W] in slot order."""
    import torch
    import torch.distributed as dist

    halo = torch.empty((plan.n_halo, send_rows.shape[1]), dtype=send_rows.dtype, device=send_rows.device)
    if plan.world == 1:
        return halo
    dist.all_to_all_single(halo, send_rows, plan.recv_counts.tolist(), plan.send_counts.tolist(), group=group)
    return halo


def choose_dense(plan: HaloPlan, threshold: float = 0.75, group=None) -> bool:
    """Collective decision: when every rank needs at least ``threshold`` of all remote rows
    (random graphs), one all-gather of the projected block beats the de-duplicated
    all-to-all-v (no per-peer send copies).  Needs equal partition sizes."""
    import torch
    import torch.distributed as dist

    if plan.world == 1:
        return False
    sizes = np.diff(plan.bounds)
    equal = bool((sizes == sizes[0]).all())
    remote = int(plan.bounds[-1]) - plan.n_local
    frac = plan.n_halo / remote if remote else 0.0
    dev = torch.device("cuda", torch.cuda.current_device()) if dist.get_backend(group) == "nccl" else torch.device("cpu")
    t = torch.tensor([frac if equal else 0.0], dtype=torch.float64, device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MIN, group=group)
    return bool(t.item() >= threshold)


def interior_runs(plan: HaloPlan, min_rows: int = 1024, max_runs: int = 4):
    """Split the local rows into *interior* ranges (every neighbour of every row is a local
    row, so they can be aggregated while the halo is still in flight) and *boundary* ranges.

    Range partitions of a locality-ordered graph have their boundary rows at the two ends of
    the range, so interior rows are taken as the (at most ``max_runs``) longest contiguous runs
    of >= ``min_rows`` interior rows: contiguous ranges keep every kernel of the path on
    plain row blocks (no row permutation, no indirection).  Returns ``(interior, boundary)``,
    two sorted lists of ``(begin, end)`` that together tile ``[0, n_local)``."""
    n = plan.n_local
    if n == 0:
        return [], []
    remote = (plan.graph.col_idx >= n).astype(np.int64)
    cs = np.concatenate([[0], np.cumsum(remote)])
    rp = plan.graph.row_ptr.astype(np.int64)
    is_int = (cs[rp[1:]] - cs[rp[:-1]]) == 0
    edge = np.flatnonzero(np.diff(np.concatenate([[0], is_int.view(np.int8), [0]])))
    starts, ends = edge[0::2], edge[1::2]
    keep = np.flatnonzero(ends - starts >= max(min_rows, 1))
    keep = keep[np.argsort(-(ends - starts)[keep], kind="stable")[:max_runs]]
    interior = sorted((int(starts[k]), int(ends[k])) for k in keep)
    boundary, at = [], 0
    for a, b in interior:
        if a > at:
            boundary.append((at, a))
        at = b
    if at < n:
        boundary.append((at, n))
    return interior, boundary


def chunk_starts(n_local: int, chunks: int) -> np.ndarray:
    """Row chunks of the dense plane: ``chunks`` near-equal contiguous pieces of the local rows."""
    chunks = max(1, min(int(chunks), max(n_local, 1)))
    return np.array([(n_local * c) // chunks for c in range(chunks + 1)], dtype=np.int64)


def dense_columns(plan: HaloPlan, starts: Optional[np.ndarray] = None) -> np.ndarray:
    """Column indices for the all-gather data plane: own rows stay local, the remote row with
    global id g = (owner r, local row i) is addressed as ``n_local + pos(g)``.  The local rows
    are gathered in chunks (``starts``, see :func:`chunk_starts`); chunk c of all ranks lands in
    one contiguous region ``[world*starts[c], world*starts[c+1])`` ordered by owner, so
    ``pos = world*starts[c] + r*size_c + (i - starts[c])`` (one chunk: pos = g)."""
    n = plan.n_local
    if starts is None:
        starts = np.array([0, n], dtype=np.int64)
    col = plan.graph.col_idx.astype(np.int64)
    remote = col >= n
    g = plan.need_ids[col[remote] - n]
    r, i = g // max(n, 1), g % max(n, 1)
    c = np.searchsorted(starts, i, side="right") - 1
    size = np.diff(starts)[c]
    out = col.copy()
    out[remote] = n + plan.world * starts[c] + r * size + (i - starts[c])
    return out.astype(np.uint32)


class Dist:
    """``ruvgnn_dist``: this rank's NCCL communicator + communication stream (csrc/dist.cu)."""

    def __init__(self, ctx, rank: int = 0, world: int = 1, unique_id: Optional[bytes] = None):
        import ctypes as C

        from . import _lib

        self._lib = _lib.load()
        self.ctx, self.rank, self.world = ctx, int(rank), int(world)
        h = C.c_void_p()
        idbuf = None  # world > 1 without an id: plan-only handle (no communicator)
        if world > 1 and unique_id is not None:
            if len(unique_id) != _lib.NCCL_ID_BYTES:
                raise ValueError("unique_id must be the 128 bytes ruvgnn_dist_unique_id() returned on rank 0")
            idbuf = (C.c_uint8 * _lib.NCCL_ID_BYTES).from_buffer_copy(unique_id)
        _lib.check(self._lib.ruvgnn_dist_create(ctx.handle, idbuf, self.rank, self.world, C.byref(h)))
        self.handle = h

    @staticmethod
    def unique_id() -> bytes:
        import ctypes as C

        from . import _lib

        buf = (C.c_uint8 * _lib.NCCL_ID_BYTES)()
        _lib.check(_lib.load().ruvgnn_dist_unique_id(buf))
        return bytes(buf)

    @classmethod
    def from_env(cls, ctx, group=None) -> "Dist":
        """One process per GPU under torchrun: rank 0 makes the NCCL id and ``torch.distributed``
        (already initialised by the host program, any backend) carries it to the other ranks."""
        import torch.distributed as dist

        if not dist.is_available() or not dist.is_initialized():
            return cls(ctx, 0, 1)
        rank, world = dist.get_rank(group), dist.get_world_size(group)
        box = [cls.unique_id() if rank == 0 else None]
        dist.broadcast_object_list(box, src=0, group=group)
        return cls(ctx, rank, world, box[0])

    def set_chunks(self, chunks: int):
        """Row chunks of the dense plane, for partitions created afterwards."""
        from . import _lib

        _lib.check(self._lib.ruvgnn_dist_set_chunks(self.handle, int(chunks)))

    def last_stats(self) -> dict:
        import ctypes as C

        from . import _lib

        st = _lib.DistStats()
        _lib.check(self._lib.ruvgnn_dist_last_stats(self.handle, C.byref(st)))
        return {k: getattr(st, k) for k, _ in st._fields_}

    def close(self):
        if getattr(self, "handle", None):
            self._lib.ruvgnn_dist_destroy(self.handle)
            self.handle = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class Partition:
    """``ruvgnn_partition``: this rank's rows of a range-partitioned graph plus the halo plan the
    library built for them on the device (collective constructor).

    ``row_ptr`` / ``col_idx`` (GLOBAL node ids) / ``edge_w`` are NumPy arrays, or -- with
    ``device_ptrs=True`` -- objects with ``data_ptr()`` (device tensors: ``row_ptr`` int64/uint64
    ``[n_local+1]``, ``col_idx`` int32/uint32, ``edge_w`` float32).

    ``forward`` / ``forward_host`` run one layer.  Default: single-pass gather once the halo has
    arrived, chunk by chunk (bit-identical to the single-GPU run); ``overlap=True``: a gather pass over
    the own-row edges under the exchange, then one pass per halo chunk (equal up to f32 summation order)."""

    def __init__(self, dist_: Dist, bounds, row_ptr, col_idx, edge_w, device_ptrs: bool = False,
                 plane: Optional[str] = None):
        import ctypes as C

        from . import _lib

        self._lib = _lib.load()
        self.dist = dist_
        self.bounds = np.ascontiguousarray(bounds, dtype=np.uint64)
        if self.bounds.size != dist_.world + 1:
            raise ValueError("bounds must have world + 1 entries")
        flags = {None: 0, "sparse": _lib.PART_FORCE_SPARSE, "dense": _lib.PART_FORCE_DENSE}[plane]
        if device_ptrs:
            flags |= _lib.PART_DEVICE_PTRS
            ptrs = [C.c_void_p(t.data_ptr()) for t in (row_ptr, col_idx, edge_w)]
            self._keep = (row_ptr, col_idx, edge_w)
        else:
            rp = np.ascontiguousarray(row_ptr, dtype=np.uint64)
            ci = np.ascontiguousarray(col_idx, dtype=np.uint32)
            ew = np.ascontiguousarray(edge_w, dtype=np.float32)
            ptrs = [C.c_void_p(a.ctypes.data) for a in (rp, ci, ew)]
            self._keep = (rp, ci, ew)
        h = C.c_void_p()
        _lib.check(self._lib.ruvgnn_partition_create(
            dist_.handle, self.bounds.ctypes.data_as(C.POINTER(C.c_uint64)), *ptrs, flags, C.byref(h)))
        self.handle = h
        self._keep = None
        info = _lib.PartitionInfo()
        _lib.check(self._lib.ruvgnn_partition_stats(h, C.byref(info)))
        self.info = {k: int(getattr(info, k)) for k, _ in info._fields_}
        self.n_local, self.n_halo = self.info["n_local"], self.info["n_halo"]
        self.dense, self.interior_rows = bool(self.info["dense"]), self.info["interior_rows"]
        self._x = self._out = None

    @classmethod
    def from_graph(cls, dist_: Dist, local_rows: CsrGraph, n_per_rank: int, plane: Optional[str] = None) -> "Partition":
        """Equal ranges of ``n_per_rank`` rows; ``local_rows`` holds this rank's rows, global columns."""
        return cls(dist_, even_bounds(n_per_rank * dist_.world, dist_.world), local_rows.row_ptr,
                   local_rows.col_idx, local_rows.edge_w, plane=plane)

    def download(self) -> dict:
        """Host copies of the device plan (checks)."""
        import ctypes as C

        i, world = self.info, self.dist.world
        col = np.empty(i["n_edges"], np.uint32)
        need = np.empty(0 if self.dense else i["n_halo"], np.uint32)
        send = np.empty(i["n_send"], np.uint32)
        rc, sc = np.empty(world, np.uint64), np.empty(world, np.uint64)
        u32, u64 = C.POINTER(C.c_uint32), C.POINTER(C.c_uint64)
        from . import _lib

        _lib.check(self._lib.ruvgnn_partition_download(
            self.handle, col.ctypes.data_as(u32), need.ctypes.data_as(u32), send.ctypes.data_as(u32),
            rc.ctypes.data_as(u64), sc.ctypes.data_as(u64)))
        return {"col_idx": col, "need_ids": need, "send_idx": send, "recv_counts": rc, "send_counts": sc}

    @staticmethod
    def _flags(overlap, msg_halo, peer_direct):
        from . import _lib

        f = _lib.DIST_OVERLAP if overlap else _lib.DIST_EXACT
        if msg_halo is not None:
            f |= _lib.DIST_MSG_HALO if msg_halo else _lib.DIST_ROW_HALO
        return f | (_lib.DIST_PEER_DIRECT if peer_direct else 0)

    def forward(self, layer, x_dev, out_dev, exact: bool = True, asynchronous: bool = True,
                msg_halo: Optional[bool] = None, peer_direct: bool = False, overlap: bool = False):
        """Device buffers (objects with ``data_ptr()`` or raw pointers); enqueue only by default.
        The work runs on the ctx stream (``ruvgnn_ctx_stream``): order the buffers against it.
        ``overlap=True`` (or ``exact=False``): the multi-pass overlapped mode; ``msg_halo``: None = the
        library's choice (message rows for K|V|M layers on the dense plane), True / False force it."""
        import ctypes as C

        from . import _lib

        xp = x_dev.data_ptr() if hasattr(x_dev, "data_ptr") else int(x_dev)
        op = out_dev.data_ptr() if hasattr(out_dev, "data_ptr") else int(out_dev)
        flags = (_lib.FWD_DEVICE_PTRS | (_lib.FWD_ASYNC if asynchronous else 0)
                 | self._flags(overlap or not exact, msg_halo, peer_direct))
        _lib.check(self._lib.ruvgnn_dist_forward(self.dist.handle, self.handle, layer._device_layer(),
                                                 C.c_void_p(xp), C.c_void_p(op), flags))

    def forward_host(self, layer, x_np: np.ndarray, out_np: np.ndarray, exact: bool = True,
                     msg_halo: Optional[bool] = None, peer_direct: bool = False, overlap: bool = False):
        """Host buffers in, host buffers out (H2D / D2H inside the call, pipelined in row blocks when the
        buffers are pinned): the e2e leg of bench.py.  Blocks until the output is there."""
        import ctypes as C

        from . import _lib

        if x_np.dtype != np.float32 or out_np.dtype != np.float32 or not x_np.flags.c_contiguous or not out_np.flags.c_contiguous:
            raise ValueError("x / out must be C-contiguous float32 arrays")
        if x_np.shape != (self.n_local, layer.input_dim) or out_np.shape != (self.n_local, layer.hidden_dim):
            raise ValueError("x must be [n_local, input_dim] and out [n_local, hidden_dim]")
        _lib.check(self._lib.ruvgnn_dist_forward(self.dist.handle, self.handle, layer._device_layer(),
                                                 C.c_void_p(x_np.ctypes.data), C.c_void_p(out_np.ctypes.data),
                                                 self._flags(overlap or not exact, msg_halo, peer_direct)))

    def close(self):
        if getattr(self, "handle", None):
            self._lib.ruvgnn_partition_destroy(self.handle)
            self.handle = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
