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


def bfs_order(g: CsrGraph) -> np.ndarray:
    """The reference's own locality reorder, ``DataLoader::reorder_bfs``
    (crates/ruvector-gnn/src/cold_tier.rs:307-342), over the CSR in stored order
    (``ruvgnn_bfs_order``, native, O(N + E)): ``perm[p]`` = the node that moves to position ``p``."""
    lib = _lib.load()
    n = g.n_nodes
    rp = np.ascontiguousarray(g.row_ptr, dtype=np.uint64)
    ci = np.ascontiguousarray(g.col_idx, dtype=np.uint32)
    order = np.empty(n, dtype=np.uint64)
    _lib.check(lib.ruvgnn_bfs_order(n, rp.ctypes.data_as(C.POINTER(C.c_uint64)), ci.ctypes.data_as(C.POINTER(C.c_uint32)),
                                    order.ctypes.data_as(C.POINTER(C.c_uint64))))
    return order.astype(np.int64)


def flatten_levels(levels: Sequence[CsrGraph]) -> CsrGraph:
    """Union-of-levels adjacency sorted by distance, the CSR analogue of hnsw_rs ``flatten_point``
    (patches/hnsw_rs/src/flatten.rs:74-89); ``levels[l]`` is the CSR of HNSW level ``l`` over the same nodes."""
    lib = _lib.load()
    if not levels:
        raise InvalidInput("at least one level is required")
    n = levels[0].n_nodes
    if any(l.n_nodes != n for l in levels):
        raise InvalidInput("every level must cover the same nodes")
    rps = [np.ascontiguousarray(l.row_ptr, dtype=np.uint64) for l in levels]
    cis = [np.ascontiguousarray(l.col_idx, dtype=np.uint32) for l in levels]
    ews = [np.ascontiguousarray(l.edge_w, dtype=np.float32) for l in levels]
    k = len(levels)
    total = sum(int(c.size) for c in cis)
    vp = C.c_void_p * k
    row_ptr = np.empty(n + 1, dtype=np.uint64)
    col_idx = np.empty(total, dtype=np.uint32)
    edge_w = np.empty(total, dtype=np.float32)
    _lib.check(lib.ruvgnn_flatten_levels(
        n, k, vp(*[a.ctypes.data for a in rps]), vp(*[a.ctypes.data for a in cis]), vp(*[a.ctypes.data for a in ews]),
        row_ptr.ctypes.data_as(C.POINTER(C.c_uint64)), col_idx.ctypes.data_as(C.POINTER(C.c_uint32)),
        edge_w.ctypes.data_as(_lib.c_float_p)))
    return CsrGraph(row_ptr, col_idx, edge_w)


def locality_order(g: CsrGraph, method: str = "bfs") -> np.ndarray:
    """A node order that pulls neighbours together: ``perm[p]`` = the node that moves to position ``p``.
    ``method="bfs"`` (default): :func:`bfs_order`, the reference's reorder, native and O(N + E);
    ``method="rcm"``: reverse Cuthill-McKee on the symmetrised adjacency (SciPy, small graphs only).

    HNSW ids are insertion order, so a contiguous id range is a random node sample and nearly
    every neighbour of a range partition is remote.  After this order most neighbours of a range
    are inside it ("optional locality permutation applied identically to X, CSR rows
    and col_idx; ... inverted on output").  Apply with :func:`permute_graph`, ``x[perm]`` and
    ``out[inverse_permutation(perm)]``."""
    if method == "bfs":
        return bfs_order(g)
    if method != "rcm":
        raise InvalidInput("method must be 'bfs' or 'rcm'")
    from scipy.sparse import csr_matrix
    from scipy.sparse.csgraph import reverse_cuthill_mckee

    n = g.n_nodes
    a = csr_matrix((np.ones(g.n_edges, dtype=np.int8), g.col_idx.astype(np.int64), g.row_ptr.astype(np.int64)),
                   shape=(n, n))
    return np.asarray(reverse_cuthill_mckee((a + a.T).tocsr(), symmetric_mode=True), dtype=np.int64)


def inverse_permutation(perm: np.ndarray) -> np.ndarray:
    inv = np.empty_like(perm)
    inv[perm] = np.arange(perm.size, dtype=perm.dtype)
    return inv


def permute_graph(g: CsrGraph, perm: np.ndarray) -> CsrGraph:
    """Relabels the nodes: new row ``p`` is old row ``perm[p]`` with every column ``j`` renamed to
    ``inverse(perm)[j]``.  The order of the neighbours inside a row is kept, so every row goes
    through exactly the arithmetic it had before (outputs are the old ones, permuted)."""
    perm = np.asarray(perm, dtype=np.int64)
    n = g.n_nodes
    if perm.shape != (n,) or not np.array_equal(np.sort(perm), np.arange(n)):
        raise InvalidInput("perm must be a permutation of the node ids")
    inv = inverse_permutation(perm)
    rp = g.row_ptr.astype(np.int64)
    deg = np.diff(rp)[perm]
    new_rp = np.zeros(n + 1, dtype=np.uint64)
    np.cumsum(deg, out=new_rp[1:].view(np.int64))
    # edge e of new row p comes from old edge rp[perm[p]] + (e - new_rp[p])
    src = np.repeat(rp[perm] - new_rp[:-1].astype(np.int64), deg) + np.arange(int(new_rp[-1]), dtype=np.int64)
    return CsrGraph(new_rp, inv[g.col_idx[src].astype(np.int64)].astype(np.uint32), g.edge_w[src].copy())


def remote_edge_fraction(g: CsrGraph, world: int) -> float:
    """Share of the edges whose endpoint lies in another part of a ``world``-way contiguous range
    partition (what the halo exchange has to serve)."""
    n = g.n_nodes
    if n == 0 or g.n_edges == 0:
        return 0.0
    bounds = np.array([(n * p) // world for p in range(world + 1)], dtype=np.int64)
    owner_of_row = np.repeat(np.searchsorted(bounds, np.arange(n), side="right") - 1, g.degrees())
    owner_of_col = np.searchsorted(bounds, g.col_idx.astype(np.int64), side="right") - 1
    return float(np.mean(owner_of_row != owner_of_col))


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
