"""Range-partitioned multi-GPU forward with halo exchange.

One process per GPU (``torch.distributed``).  Nodes are split into contiguous ranges;
rank p owns rows ``[bounds[p], bounds[p+1])`` of X and of the CSR, projects its own rows,
and needs the projected K|V|M rows of every *remote* neighbour its rows reference.  Those
are de-duplicated per peer once per graph (:class:`HaloPlan`, pure host logic) and moved
once per layer with a single all-to-all-v over NCCL/NVLink into a halo buffer the
gather kernel addresses as rows ``n_local + slot``.  Neighbour order inside every row is
preserved (indices are remapped, never reordered), so a partitioned run performs exactly
the per-row arithmetic of the single-GPU run.

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


class Partition:
    """Device-side state of one rank: plan, device CSR, send indices, buffers.

    Two data planes with identical results: ``sparse`` = pack the rows each peer asked for
    and exchange them with one all-to-all-v; ``dense`` = copy the whole projected K|V|M
    block contiguously and all-gather it (chosen collectively when the halo is almost
    everything, see :func:`choose_dense`).

    The dense plane moves message rows only (H floats instead of the 3H of K|V|M): the rows are
    all-gathered in ``chunks`` pieces and K|V of each piece is rebuilt on arrival
    (``ruvgnn_layer_expand_halo``, the projection's own GEMM, so the bits are the owner's) while
    the next piece is still on the wire; during that phase the kernels' grids leave ``comm_sms``
    SMs to the collective.

    ``overlap``: the collective is issued asynchronously (NCCL runs it on its own stream) and
    the interior row ranges (:func:`interior_runs`) are aggregated and updated while it is in
    flight; the launching stream then waits for the halo and finishes the boundary ranges.
    Each row sees exactly the same arithmetic either way."""

    def __init__(self, plan: HaloPlan, ctx, dense: bool = False, overlap: bool = True, chunks: int = 4,
                 comm_sms: Optional[int] = None):
        import torch

        from .graph import CsrGraph

        self.plan, self.ctx, self.dense = plan, ctx, dense
        import os

        # SMs left to the collective while kernels run beside it.  Measured on 2 B200s: 0, 16 and 32 give the
        # same step time (the GEMMs that run beside the exchange are L2-bandwidth-bound and so is the
        # collective's copy, so they share that, not SMs) and capping NCCL's blocks costs bandwidth: default 0.
        self.comm_sms = int(os.environ.get("RUVGNN_COMM_SMS", "0")) if comm_sms is None else int(comm_sms)
        if dense:
            n_total = int(plan.bounds[-1])
            self.starts = chunk_starts(plan.n_local, chunks)
            g = CsrGraph(plan.graph.row_ptr, dense_columns(plan, self.starts), plan.graph.edge_w)
            self.graph = g.to_device(ctx, n_cols=plan.n_local + n_total)
            self.send_idx = torch.zeros(0, dtype=torch.int32, device="cuda")
        else:
            self.graph = plan.graph.to_device(ctx, n_cols=plan.n_local + plan.n_halo)
            self.send_idx = torch.from_numpy(plan.send_idx.astype(np.int64)).to(torch.int32).cuda().contiguous()
        self._send = self._halo = self._x = self._out = self._s_in = self._s_out = None
        self._ring = []
        self.bytes_exchanged = 0
        self.interior, self.boundary = interior_runs(plan) if overlap else ([], [(0, plan.n_local)] if plan.n_local else [])
        self.interior_rows = sum(b - a for a, b in self.interior)
        self._wait_ev = None
        self.trace = None  # set to [] to collect (label, event) pairs of the next forward (diagnostics)

    def _mark(self, label):
        if self.trace is not None:
            import torch

            ev = torch.cuda.Event(enable_timing=True)
            ev.record()
            self.trace.append((label, ev))

    def trace_ms(self):
        """[(label, ms since the previous mark)] of the traced forward.  Synchronises."""
        out = []
        for (_, e0), (label, e1) in zip(self.trace, self.trace[1:]):
            e1.synchronize()
            out.append((label, round(float(e0.elapsed_time(e1)), 3)))
        return out

    @classmethod
    def build(cls, local_rows: CsrGraph, rank: int, world: int, n_per_rank: int, ctx, overlap: bool = True) -> "Partition":
        plan = exchange_requests(HaloPlan.build(local_rows, rank, even_bounds(n_per_rank * world, world)))
        return cls(plan, ctx, dense=choose_dense(plan), overlap=overlap)

    def exposed_wait_ms(self) -> float:
        """Device time the launching stream spent waiting for the halo in the last forward
        (the part of the exchange that interior compute did not hide).  Synchronises."""
        if self._wait_ev is None:
            return 0.0
        self._wait_ev[1].synchronize()
        return float(self._wait_ev[0].elapsed_time(self._wait_ev[1]))

    def forward(self, layer, x_dev, out_dev, timing: bool = False):
        """project -> pack -> exchange (NCCL, asynchronous) || interior rows -> boundary rows.
        Enqueue only."""
        self._forward(layer, x_dev, out_dev)

    def _forward(self, layer, x_dev, out_dev, x_host=None, out_host=None, blocks: int = 8):
        """Device buffers, or (``x_host`` / ``out_host`` given, pinned) the host-buffer pipeline: X is
        copied in row blocks on a copy stream and each block is projected as it lands; each finished
        row range of the output is copied back on a second copy stream while the next range computes."""
        import ctypes as C

        import torch
        import torch.distributed as dist

        from . import _lib

        lib, plan, H = _lib.load(), self.plan, layer.hidden_dim
        h = layer._device_layer()
        n = plan.n_local
        main = torch.cuda.current_stream()
        self._mark("start")
        if x_host is None:
            _lib.check(lib.ruvgnn_layer_project(self.ctx.handle, h, n, C.c_void_p(x_dev.data_ptr())))
        else:
            if self._s_in is None:
                self._s_in, self._s_out = torch.cuda.Stream(), torch.cuda.Stream()
            self._s_in.wait_stream(main)
            self._s_out.wait_stream(main)
            edges = [(n * b) // blocks for b in range(blocks + 1)]
            landed = []
            with torch.cuda.stream(self._s_in):
                for a, b in zip(edges, edges[1:]):
                    if b > a:
                        x_dev[a:b].copy_(x_host[a:b], non_blocking=True)
                        landed.append((a, b, self._s_in.record_event()))
            for a, b, ev in landed:
                main.wait_event(ev)
                _lib.check(lib.ruvgnn_layer_project_rows(self.ctx.handle, h, n, C.c_void_p(x_dev.data_ptr()), a, b))
        self._mark("project")
        W = int(lib.ruvgnn_layer_halo_width(h))  # 3H (K|V|M rows) or H (message rows, R=1 layers)

        work, works = None, {}
        if self.dense:
            n_total, P = int(plan.bounds[-1]), plan.world
            nchunks = len(self.starts) - 1
            if self._send is None or self._send.shape != (n, H):
                self._send = torch.empty((n, H), dtype=torch.float32, device=x_dev.device)
                self._halo = torch.empty((n_total, W), dtype=torch.float32, device=x_dev.device)
                # K|V|M halo rows: the gathered message rows pass through a 2-deep ring of chunk
                # buffers; message-row halos (W == H) are gathered straight into place
                big = int(np.diff(self.starts).max())
                self._ring = [] if W == H else [torch.empty((P * big, H), dtype=torch.float32, device=x_dev.device)
                                                for _ in range(min(2, nchunks))]
            _lib.check(lib.ruvgnn_layer_messages(self.ctx.handle, h, 0, n, C.c_void_p(self._send.data_ptr())))

            def landing(c):
                a, b = int(self.starts[c]), int(self.starts[c + 1])
                return self._halo[P * a:P * b] if W == H else self._ring[c % 2][:P * (b - a)]

            def issue(c):  # the collective starts after everything enqueued so far on this stream
                a, b = int(self.starts[c]), int(self.starts[c + 1])
                works[c] = dist.all_gather_into_tensor(landing(c), self._send[a:b], async_op=True)

            for c in range(nchunks if W == H else min(2, nchunks)):
                issue(c)
            self.bytes_exchanged = (n_total - n) * H * 4 * 2
        else:
            n_send = int(self.send_idx.numel())
            if self._send is None or self._send.shape != (n_send, W):
                self._send = torch.empty((n_send, W), dtype=torch.float32, device=x_dev.device)
            if n_send:
                _lib.check(lib.ruvgnn_halo_pack(self.ctx.handle, h, C.c_void_p(self.send_idx.data_ptr()), n_send,
                                                C.c_void_p(self._send.data_ptr())))
            if self._halo is None or self._halo.shape != (plan.n_halo, W):
                self._halo = torch.empty((plan.n_halo, W), dtype=torch.float32, device=x_dev.device)
            if plan.world > 1:
                work = dist.all_to_all_single(self._halo, self._send, plan.recv_counts.tolist(),
                                              plan.send_counts.tolist(), async_op=True)
            self.bytes_exchanged = (n_send + plan.n_halo) * W * 4
        halo_ptr = self._halo.data_ptr() if (plan.n_halo or self.dense) else None
        step = max((n + blocks - 1) // blocks, 1) if out_host is not None else max(n, 1)

        def rows(a, b):
            for a0 in range(a, b, step):  # host pipeline: hand finished sub-ranges to the copy-out stream
                b0 = min(a0 + step, b)
                _lib.check(lib.ruvgnn_layer_aggregate_update(self.ctx.handle, h, self.graph.handle,
                                                             C.c_void_p(halo_ptr), a0, b0, C.c_void_p(out_dev.data_ptr())))
                if out_host is not None:
                    self._s_out.wait_event(main.record_event())
                    with torch.cuda.stream(self._s_out):
                        out_host[a0:b0].copy_(out_dev[a0:b0], non_blocking=True)

        self._mark("pack")
        limited = bool(works or work is not None) and self.comm_sms > 0
        if limited:  # leave SMs to the collective: persistent grids must not queue behind its blocks
            _lib.check(lib.ruvgnn_ctx_set_sm_limit(self.ctx.handle, -self.comm_sms))
        for a, b in self.interior:  # no remote neighbour: runs while the collective is in flight
            rows(a, b)
        if self._wait_ev is None:
            self._wait_ev = (torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True))
        self._wait_ev[0].record()
        if work is not None:
            work.wait()  # the launching stream waits for the halo; the host does not block
        if self.dense:  # rebuild K|V of chunk c while chunk c+1 is on the wire
            P, r = plan.world, plan.rank
            for c in range(nchunks):
                works[c].wait()
                self._mark(f"wait{c}")
                if W == H:
                    continue
                a, b = int(self.starts[c]), int(self.starts[c + 1])
                src, dst = landing(c), self._halo[P * a:P * b]
                for lo, hi in ((0, r * (b - a)), ((r + 1) * (b - a), P * (b - a))):  # skip own block
                    if hi > lo:
                        _lib.check(lib.ruvgnn_layer_expand_halo(self.ctx.handle, h, C.c_void_p(src[lo:hi].data_ptr()),
                                                                hi - lo, C.c_void_p(dst[lo:hi].data_ptr())))
                self._mark(f"expand{c}")
                if c + 2 < nchunks:
                    issue(c + 2)  # reuses this chunk's ring buffer: ordered after the rebuild just enqueued
        self._wait_ev[1].record()
        if limited:
            _lib.check(lib.ruvgnn_ctx_set_sm_limit(self.ctx.handle, 0))
        for a, b in self.boundary:
            rows(a, b)
        self._mark("boundary")
        if out_host is not None:
            main.wait_stream(self._s_out)

    def forward_host(self, layer, x_np: np.ndarray, out_np: np.ndarray):
        """Host buffers in, host buffers out (H2D / D2H inside, pipelined against the projection and the
        update when the buffers are pinned): the e2e leg of bench.py.  Blocks until the output is there."""
        import torch

        xt, ot = torch.from_numpy(x_np), torch.from_numpy(out_np)
        if self._x is None or self._x.shape != xt.shape:
            self._x = torch.empty(xt.shape, dtype=torch.float32, device="cuda")
            self._out = torch.empty(ot.shape, dtype=torch.float32, device="cuda")
        self._forward(layer, self._x, self._out, x_host=xt, out_host=ot)
        torch.cuda.current_stream().synchronize()
