"""Batched mirror of ``ruvector_gnn::search`` (crates/ruvector-gnn/src/search.rs) on the GPU.

``differentiable_search(query, candidate_embeddings, k, temperature)`` (search.rs:56-94) and
``hierarchical_forward(query, layer_embeddings, gnn_layers)`` (:105-162) keep their argument
meaning; ``query`` may be one vector or a batch ``[Q, dim]`` of independent queries.
"""
from __future__ import annotations

import ctypes as C
from typing import Sequence

import numpy as np

from . import _lib
from .errors import DimensionMismatch
from .layer import RuvectorLayer, default_context

F32 = np.float32
_fp = _lib.c_float_p


def differentiable_search(query, candidate_embeddings, k: int, temperature: float, ctx=None):
    """Returns ``(indices, soft_weights)``; for a batch of queries both are ``[Q, k]`` arrays."""
    q = np.ascontiguousarray(query, dtype=F32)
    single = q.ndim == 1
    q = q.reshape(1, -1) if single else q
    cands = np.ascontiguousarray(candidate_embeddings, dtype=F32)
    if cands.size == 0:
        empty = (np.zeros((q.shape[0], 0), np.uint32), np.zeros((q.shape[0], 0), F32))
        return (empty[0][0], empty[1][0]) if single else empty
    cands = cands.reshape(-1, cands.shape[-1])
    if cands.shape[1] != q.shape[1]:
        raise DimensionMismatch(f"Tensor dimension mismatch: expected {q.shape[1]}, got {cands.shape[1]}")
    ctx = ctx or default_context(0)
    ke = min(int(k), cands.shape[0])
    idx = np.empty((q.shape[0], ke), dtype=np.uint32)
    w = np.empty((q.shape[0], ke), dtype=F32)
    kout = C.c_uint32()
    _lib.check(_lib.load().ruvgnn_differentiable_search(
        ctx.handle, q.ctypes.data_as(_fp), q.shape[0], cands.ctypes.data_as(_fp), cands.shape[0], q.shape[1],
        int(k), float(temperature), idx.ctypes.data_as(C.POINTER(C.c_uint32)), w.ctypes.data_as(_fp),
        C.byref(kout), 0))
    return (idx[0], w[0]) if single else (idx, w)


def hierarchical_forward(query, layer_embeddings: Sequence, gnn_layers: Sequence[RuvectorLayer], ctx=None):
    """Top-to-bottom pass through per-level embeddings and GNN layers (search.rs:105-162)."""
    q = np.ascontiguousarray(query, dtype=F32)
    single = q.ndim == 1
    q = q.reshape(1, -1) if single else q
    n_levels = min(len(layer_embeddings), len(gnn_layers))  # zip semantics (:118-119)
    if n_levels == 0:
        return q[0].copy() if single else q.copy()
    ctx = ctx or gnn_layers[0].ctx
    embs = []
    for e in layer_embeddings[:n_levels]:
        a = np.ascontiguousarray(e, dtype=F32)
        embs.append(a.reshape(-1, a.shape[-1]) if a.size else np.zeros((0, 1), F32))
    rows = (C.c_uint64 * n_levels)(*[e.shape[0] for e in embs])
    dims = (C.c_uint32 * n_levels)(*[e.shape[1] for e in embs])
    ptrs = (_fp * n_levels)(*[e.ctypes.data_as(_fp) for e in embs])
    lh = (C.c_void_p * n_levels)(*[l._device_layer() for l in gnn_layers[:n_levels]])
    cap = max([q.shape[1]] + [l.hidden_dim for l in gnn_layers[:n_levels]])
    od = C.c_uint32()
    # rows are written densely with the final width (returned in od)
    flat = np.empty(q.shape[0] * cap, dtype=F32)
    _lib.check(_lib.load().ruvgnn_hierarchical_forward(
        ctx.handle, q.ctypes.data_as(_fp), q.shape[0], q.shape[1], ptrs, rows, dims, lh, n_levels,
        flat.ctypes.data_as(_fp), flat.size, C.byref(od)))
    out = flat[: q.shape[0] * od.value].reshape(q.shape[0], od.value).copy()
    return out[0] if single else out
