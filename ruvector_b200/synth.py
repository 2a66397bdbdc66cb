"""Seeded synthetic inputs shared by the oracle, the tests and bench.py.

Host-side generators only (NumPy Philox streams); nothing here computes the layer.
"""
from __future__ import annotations

import numpy as np

from .graph import CsrGraph

F32 = np.float32
LINEARS = ("w_msg", "w_agg", "w_z", "u_z", "w_r", "u_r", "w_h", "u_h", "q", "k", "v", "out")


def _rng(key: int, stream: int = 0) -> np.random.Generator:
    return np.random.Generator(np.random.Philox(key=[key, stream]))


def make_features(n: int, d: int, seed: int = 0x5EED) -> np.ndarray:
    """N(0,1) rows, L2-normalised (cosine is the DB's default metric, ruvector-core types.rs:116-124)."""
    x = _rng(seed).standard_normal((n, d), dtype=F32)
    x /= np.maximum(np.linalg.norm(x, axis=1, keepdims=True), 1e-12).astype(F32)
    return x


def make_weights(input_dim: int, hidden_dim: int, seed: int = 0xC0FFEE, trivial: bool = False) -> dict:
    """Xavier-normal weights (layer.rs:25-29).  ``trivial=False`` adds non-zero biases and a
    non-trivial gamma/beta so parity runs exercise every term; ``trivial=True`` is the
    reference's own initialisation (zero biases, gamma=1, beta=0; layer.rs:31,60-62)."""
    w = {}
    for i, name in enumerate(LINEARS):
        fan_in = input_dim if name == "w_msg" else hidden_dim
        g = _rng(seed, i)
        scale = float(np.sqrt(2.0 / (fan_in + hidden_dim)))
        W = (g.standard_normal((hidden_dim, fan_in), dtype=F32) * F32(scale)).astype(F32)
        b = np.zeros(hidden_dim, dtype=F32) if trivial else g.uniform(-0.1, 0.1, hidden_dim).astype(F32)
        w[name] = (W, b)
    g = _rng(seed, 100)
    w["gamma"] = np.ones(hidden_dim, dtype=F32) if trivial else g.uniform(0.5, 1.5, hidden_dim).astype(F32)
    w["beta"] = np.zeros(hidden_dim, dtype=F32) if trivial else g.uniform(-0.1, 0.1, hidden_dim).astype(F32)
    w["eps"] = 1e-5
    return w


def hnsw_like_graph(n: int, m: int, seed: int = 0xBEEF, locality: float = 0.0, window: int = 4096,
                    iso_frac: float = 0.01, level: int = 0, n_total: int | None = None,
                    row_offset: int = 0) -> CsrGraph:
    """HNSW-shaped random adjacency (no self loops, no duplicate neighbour in a row).

    Level 0: out-degree ~ U{M..2M} (cap 2M, hnsw.rs:1169); level l>=1: U{M/2..M} over the
    ~n/M^l nodes present, the rest empty.  ``locality`` is the fraction of neighbours drawn
    from a +-``window`` id range (0 = insertion-order worst case).  ``iso_frac`` of the
    nodes get no neighbours (layer.rs:385-389 path).  ``edge_w`` is ascending U(0,1) per
    row, like HNSW's distance-sorted lists (hnsw.rs:1186).

    ``n_total`` / ``row_offset``: generate only the rows ``[row_offset, row_offset + n)`` of an
    ``n_total``-node graph, with GLOBAL column ids (one partition of a multi-GPU run).
    """
    nt = n if n_total is None else n_total
    g = _rng(seed, level)
    lo, hi = (m, 2 * m) if level == 0 else (max(m // 2, 1), m)
    hi = min(hi, max(nt - 1, 0))
    lo = min(lo, hi)
    deg = g.integers(lo, hi + 1, size=n, dtype=np.int64) if hi > 0 else np.zeros(n, dtype=np.int64)
    if level > 0:
        present = g.random(n) < float(m) ** (-level)
        deg[~present] = 0
    if iso_frac > 0:
        deg[g.random(n) < iso_frac] = 0
    width = int(deg.max()) if n else 0
    row_ptr = np.zeros(n + 1, dtype=np.uint64)
    np.cumsum(deg, out=row_ptr[1:].view(np.int64))
    if width == 0:
        return CsrGraph(row_ptr, np.zeros(0, np.uint32), np.zeros(0, F32))
    rows = row_offset + np.arange(n, dtype=np.int64)[:, None]

    def draw(shape_rows):
        r = shape_rows.shape[0]
        far = g.integers(0, nt - 1, size=(r, width), dtype=np.int64)
        far += far >= shape_rows  # skip self
        if locality <= 0:
            return far
        w = min(window, max((nt - 1) // 2, 1))
        off = g.integers(1, w + 1, size=(r, width), dtype=np.int64) * g.choice(np.array([-1, 1]), size=(r, width))
        near = (shape_rows + off) % nt
        near = np.where(near == shape_rows, (near + 1) % nt, near)
        return np.where(g.random((r, width)) < locality, near, far)

    cols = draw(rows)
    mask = np.arange(width)[None, :] < deg[:, None]
    for _ in range(64):  # re-draw rows that contain a duplicate among their first deg entries
        srt = np.sort(np.where(mask, cols, -1 - np.arange(width)[None, :]), axis=1)
        bad = np.nonzero((srt[:, 1:] == srt[:, :-1]).any(axis=1))[0]
        if bad.size == 0:
            break
        cols[bad] = draw(rows[bad])
    else:
        raise RuntimeError("could not draw duplicate-free neighbour rows")
    wts = np.sort(g.random((n, width), dtype=F32), axis=1)
    return CsrGraph(row_ptr, cols[mask].astype(np.uint32), wts[mask].astype(F32))


def knn_graph(x: np.ndarray, m: int, seed: int = 0xBEEF, iso_frac: float = 0.01) -> CsrGraph:
    """Exact cosine kNN adjacency for small n (config C1): node i keeps its deg_i ~ U{M..2M}
    nearest neighbours in ascending cosine distance, ``edge_w`` = that distance."""
    n = x.shape[0]
    g = _rng(seed, 7)
    kmax = min(2 * m, n - 1)
    deg = g.integers(min(m, kmax), kmax + 1, size=n, dtype=np.int64)
    if iso_frac > 0:
        deg[g.random(n) < iso_frac] = 0
    xn = x / np.maximum(np.linalg.norm(x, axis=1, keepdims=True), 1e-12)
    cols_out, w_out = [], []
    step = 2048
    for s in range(0, n, step):
        sim = (xn[s:s + step] @ xn.T).astype(np.float64)
        sim[np.arange(sim.shape[0]), np.arange(s, s + sim.shape[0])] = -np.inf
        part = np.argpartition(-sim, kmax - 1, axis=1)[:, :kmax]
        psim = np.take_along_axis(sim, part, axis=1)
        order = np.argsort(-psim, axis=1, kind="stable")
        idx = np.take_along_axis(part, order, axis=1)
        dist = (1.0 - np.take_along_axis(psim, order, axis=1)).astype(F32)
        for r in range(idx.shape[0]):
            d = int(deg[s + r])
            cols_out.append(idx[r, :d])
            w_out.append(dist[r, :d])
    row_ptr = np.zeros(n + 1, dtype=np.uint64)
    np.cumsum(deg, out=row_ptr[1:].view(np.int64))
    return CsrGraph(row_ptr, np.concatenate(cols_out).astype(np.uint32), np.concatenate(w_out).astype(F32))
