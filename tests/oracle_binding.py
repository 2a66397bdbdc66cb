"""ctypes binding of oracle/liboracle.so -- the checker.  Imported only by tests/,
__graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs."""
from __future__ import annotations

import ctypes as C
import os
from pathlib import Path

import numpy as np

from ruvector_b200._lib import Weights, c_float_p

ROOT = Path(__file__).resolve().parents[1]
F32 = np.float32


class OrcLayer(C.Structure):
    _fields_ = [("input_dim", C.c_uint32), ("hidden_dim", C.c_uint32), ("heads", C.c_uint32),
                ("dropout", C.c_float), ("w", Weights)]


def _p(a):
    return a.ctypes.data_as(c_float_p)


class HostLayer:
    """Host-only stand-in for ruvector_b200.RuvectorLayer (dims + weights, nothing else): lets
    bench.py's --impl reference arm feed the oracle without ever loading libruvgnn.so."""

    def __init__(self, input_dim, hidden_dim, heads, dropout, weights):
        self.input_dim, self.hidden_dim, self.heads = int(input_dim), int(hidden_dim), int(heads)
        self.dropout = float(np.float32(dropout))
        self.weights = {k: ((np.ascontiguousarray(v[0], dtype=F32), np.ascontiguousarray(v[1], dtype=F32))
                            if isinstance(v, tuple) else v) for k, v in weights.items()}
        self.weights["gamma"] = np.ascontiguousarray(weights["gamma"], dtype=F32)
        self.weights["beta"] = np.ascontiguousarray(weights["beta"], dtype=F32)

    def c_weights(self) -> Weights:
        from ruvector_b200._lib import LINEAR_NAMES, Linear

        cw = Weights()
        for name in LINEAR_NAMES:
            W, b = self.weights[name]
            setattr(cw, name, Linear(_p(W), _p(b)))
        cw.gamma, cw.beta = _p(self.weights["gamma"]), _p(self.weights["beta"])
        cw.eps = float(self.weights.get("eps", 1e-5))
        return cw


class Oracle:
    def __init__(self):
        path = ROOT / "oracle" / "liboracle.so"
        if not path.exists():
            import subprocess

            subprocess.run(["make", "-C", str(ROOT / "oracle"), "liboracle.so"], check=True,
                           stdout=subprocess.DEVNULL)
        self.lib = C.CDLL(str(path))
        L = self.lib
        u64p, u32p = C.POINTER(C.c_uint64), C.POINTER(C.c_uint32)
        L.orc_layer_forward_node.restype = C.c_int
        L.orc_layer_forward_node.argtypes = [C.POINTER(OrcLayer), c_float_p, c_float_p, C.c_uint32,
                                             c_float_p, C.c_uint32, c_float_p]
        L.orc_forward_graph_faithful.restype = C.c_int
        L.orc_forward_graph_faithful.argtypes = [C.POINTER(OrcLayer), C.c_uint64, c_float_p, u64p, u32p,
                                                 c_float_p, C.c_uint64, C.c_uint64, c_float_p, C.c_int]
        L.orc_forward_graph_cached.restype = C.c_int
        L.orc_forward_graph_cached.argtypes = [C.POINTER(OrcLayer), C.c_uint64, c_float_p, u64p, u32p,
                                               c_float_p, c_float_p, C.c_int]
        L.orc_linear_forward.argtypes = [c_float_p, c_float_p, C.c_uint32, C.c_uint32, c_float_p, c_float_p]
        L.orc_layernorm_forward.argtypes = [c_float_p, c_float_p, C.c_float, C.c_uint32, c_float_p, c_float_p]
        L.orc_sigmoid.restype = C.c_float
        L.orc_sigmoid.argtypes = [C.c_float]
        L.orc_cosine_similarity.restype = C.c_float
        L.orc_cosine_similarity.argtypes = [c_float_p, c_float_p, C.c_uint32]
        L.orc_softmax_t.argtypes = [c_float_p, C.c_uint32, C.c_float, c_float_p]
        L.orc_differentiable_search.restype = C.c_uint32
        L.orc_differentiable_search.argtypes = [c_float_p, c_float_p, C.c_uint32, C.c_uint32, C.c_uint32,
                                                C.c_float, u32p, c_float_p]

    def differentiable_search(self, query, cands, k, temperature):
        q = np.ascontiguousarray(query, dtype=F32)
        c = np.ascontiguousarray(cands, dtype=F32)
        idx = np.empty(max(k, 1), dtype=np.uint32)
        w = np.empty(max(k, 1), dtype=F32)
        n = self.lib.orc_differentiable_search(_p(q), _p(c), c.shape[0], c.shape[1], k, temperature,
                                               idx.ctypes.data_as(C.POINTER(C.c_uint32)), _p(w))
        return idx[:n], w[:n]

    def hierarchical_forward(self, query, level_embs, layers) -> np.ndarray:
        n = min(len(level_embs), len(layers))
        embs = [np.ascontiguousarray(e, dtype=F32).reshape(-1, np.asarray(e).shape[-1]) for e in level_embs[:n]]
        ols = [self.c_layer(l) for l in layers[:n]]
        eptr = (c_float_p * n)(*[_p(e) for e in embs])
        rows = (C.c_uint32 * n)(*[e.shape[0] for e in embs])
        dims = (C.c_uint32 * n)(*[e.shape[1] for e in embs])
        lptr = (C.POINTER(OrcLayer) * n)(*[C.pointer(o) for o in ols])
        q = np.ascontiguousarray(query, dtype=F32)
        out = np.empty(4096, dtype=F32)
        self.lib.orc_hierarchical_forward.restype = C.c_uint32
        self.lib.orc_hierarchical_forward.argtypes = [c_float_p, C.c_uint32, C.POINTER(c_float_p), C.POINTER(C.c_uint32),
                                                      C.POINTER(C.c_uint32), C.POINTER(C.POINTER(OrcLayer)), C.c_uint32,
                                                      c_float_p]
        ln = self.lib.orc_hierarchical_forward(_p(q), q.size, eptr, rows, dims, lptr, n, _p(out))
        return out[:ln].copy()

    @staticmethod
    def c_layer(layer) -> OrcLayer:
        """layer: ruvector_b200.RuvectorLayer (only its host weights are read)."""
        ol = OrcLayer(layer.input_dim, layer.hidden_dim, layer.heads, layer.dropout, layer.c_weights())
        ol._keep = layer  # keep the numpy buffers alive
        return ol

    def forward_node(self, layer, x, nbrs, w) -> np.ndarray:
        x = np.ascontiguousarray(x, dtype=F32)
        nbrs = np.ascontiguousarray(nbrs, dtype=F32).reshape(-1, layer.input_dim)
        w = np.ascontiguousarray(w, dtype=F32)
        out = np.empty(layer.hidden_dim, dtype=F32)
        ol = self.c_layer(layer)
        rc = self.lib.orc_layer_forward_node(C.byref(ol), _p(x), _p(nbrs), nbrs.shape[0], _p(w), w.size, _p(out))
        assert rc == 0
        return out

    def forward_graph(self, layer, x, g, threads: int | None = None, faithful: bool = False,
                      node_range=None) -> np.ndarray:
        x = np.ascontiguousarray(x, dtype=F32)
        n = x.shape[0]
        out = np.zeros((n, layer.hidden_dim), dtype=F32)
        ol = self.c_layer(layer)
        threads = threads or os.cpu_count() or 1
        rp = g.row_ptr.ctypes.data_as(C.POINTER(C.c_uint64))
        ci = g.col_idx.ctypes.data_as(C.POINTER(C.c_uint32))
        if faithful or node_range is not None:
            b, e = node_range if node_range is not None else (0, n)
            rc = self.lib.orc_forward_graph_faithful(C.byref(ol), n, _p(x), rp, ci, _p(g.edge_w), b, e, _p(out), threads)
        else:
            rc = self.lib.orc_forward_graph_cached(C.byref(ol), n, _p(x), rp, ci, _p(g.edge_w), _p(out), threads)
        assert rc == 0
        return out
