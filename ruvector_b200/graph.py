"""HNSW neighbour lists -> CSR (host side of SURVEY.md row a12).

The reference never exposes its HNSW adjacency to ``ruvector-gnn`` (SURVEY.md
fact 2); the ground truth is hnsw_rs's ``Point::get_neighborhood_id()``
(patches/hnsw_rs/src/hnsw.rs:220-240) which returns, per point and per layer,
a ``Vec<Neighbour>`` of ``#[repr(C)] {d_id: usize, distance: f32, p_id}``
records (hnsw.rs:97-106).  :func:`pack_neighbour_lists` feeds such records to
the C packer ``ruvgnn_pack_neighbours``; integer/byte work, bit-exact.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass
from typing import Optional, Sequence

import numpy as np

from . import _lib
from .errors import InvalidInput

# hnsw_rs `Neighbour` as laid out by rustc for #[repr(C)] on x86-64:
# usize d_id @0, f32 distance @8, PointId(u8, i32) @12 (8 bytes) -> 24 bytes.
NEIGHBOUR_DTYPE = np.dtype(
    {"names": ["d_id", "distance", "p_layer", "p_rank"],
     "formats": [np.uint64, np.float32, np.uint8, np.int32],
     "offsets": [0, 8, 12, 16], "itemsize": 24}
)


@dataclass
class CsrGraph:
    """Host CSR of one HNSW level: ``row_ptr`` u64[N+1], ``col_idx`` u32[E], ``edge_w`` f32[E]."""

    row_ptr: np.ndarray
    col_idx: np.ndarray
    edge_w: np.ndarray

    @property
    def n_nodes(self) -> int:
        return int(self.row_ptr.size - 1)

    @property
    def n_edges(self) -> int:
        return int(self.col_idx.size)

    def degrees(self) -> np.ndarray:
        return np.diff(self.row_ptr.astype(np.int64))

    def to_device(self, ctx=None, n_cols: Optional[int] = None, flags: int = 0):
        from .layer import DeviceGraph, default_context

        return DeviceGraph(ctx or default_context(0), self.row_ptr, self.col_idx, self.edge_w, n_cols, flags)


def pack_neighbour_lists(lists: Sequence[np.ndarray]) -> CsrGraph:
    """``lists[i]`` is node i's ``NEIGHBOUR_DTYPE`` record array (one HNSW level), in stored order."""
    lib = _lib.load()
    n = len(lists)
    recs = [np.ascontiguousarray(l, dtype=NEIGHBOUR_DTYPE) for l in lists]
    counts = np.array([r.size for r in recs], dtype=np.uint64)
    total = int(counts.sum())
    ptrs = (C.c_void_p * max(n, 1))(*[r.ctypes.data if r.size else None for r in recs])
    row_ptr = np.empty(n + 1, dtype=np.uint64)
    col_idx = np.empty(total, dtype=np.uint32)
    edge_w = np.empty(total, dtype=np.float32)
    _lib.check(
        lib.ruvgnn_pack_neighbours(
            n, ptrs, counts.ctypes.data_as(C.POINTER(C.c_uint64)), NEIGHBOUR_DTYPE.itemsize,
            NEIGHBOUR_DTYPE.fields["d_id"][1], NEIGHBOUR_DTYPE.fields["distance"][1],
            row_ptr.ctypes.data_as(C.POINTER(C.c_uint64)), col_idx.ctypes.data_as(C.POINTER(C.c_uint32)),
            edge_w.ctypes.data_as(_lib.c_float_p),
        )
    )
    return CsrGraph(row_ptr, col_idx, edge_w)


def pack_id_distance_lists(ids: Sequence[Sequence[int]], dists: Sequence[Sequence[float]]) -> CsrGraph:
    """Convenience: builds the ``Neighbour`` records from parallel id / distance lists."""
    if len(ids) != len(dists):
        raise InvalidInput("ids and dists must have one list per node")
    recs = []
    for a, d in zip(ids, dists):
        if len(a) != len(d):
            raise InvalidInput("each node needs as many distances as neighbour ids")
        r = np.zeros(len(a), dtype=NEIGHBOUR_DTYPE)
        r["d_id"] = np.asarray(a, dtype=np.uint64)
        r["distance"] = np.asarray(d, dtype=np.float32)
        recs.append(r)
    return pack_neighbour_lists(recs)


def batch_to_graph(operations) -> tuple:
    """``GnnCache::batch_forward`` request shape -> (X, CsrGraph, query_rows).

    ``operations`` is a list of ``(node_embedding, neighbor_embeddings, edge_weights)``
    (crates/ruvector-cli/src/mcp/gnn_cache.rs:309-313).  Every op becomes one
    query row followed by its neighbour rows (no out-edges), so one batched
    forward computes all ops; ``query_rows`` indexes the results.  Requires
    ``len(edge_weights) == len(neighbor_embeddings)`` (the wasm binding's check,
    crates/ruvector-gnn-wasm/src/lib.rs:112-119).
    """
    rows, row_ptr, cols, ws, qrows = [], [0], [], [], []
    for node, nbrs, w in operations:
        if len(nbrs) != len(w):
            raise InvalidInput("Number of neighbors must match number of edge weights")
        q = len(rows)
        qrows.append(q)
        rows.append(np.asarray(node, dtype=np.float32))
        cols.extend(range(q + 1, q + 1 + len(nbrs)))
        ws.extend(np.asarray(w, dtype=np.float32).tolist())
        row_ptr.append(len(cols))
        for v in nbrs:
            rows.append(np.asarray(v, dtype=np.float32))
            row_ptr.append(len(cols))
    x = np.stack(rows) if rows else np.zeros((0, 0), dtype=np.float32)
    g = CsrGraph(np.asarray(row_ptr, dtype=np.uint64), np.asarray(cols, dtype=np.uint32),
                 np.asarray(ws, dtype=np.float32))
    return x, g, np.asarray(qrows, dtype=np.int64)
