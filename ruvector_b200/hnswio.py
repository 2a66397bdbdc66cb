"""hnsw_rs graph dumps -> per-level CSR.

``HnswDump(path)`` parses the ``<basename>.hnsw.graph`` file ``Hnsw::file_dump`` writes
(patches/hnsw_rs/src/hnswio.rs:878-918, 1063-1096, 1303-1340) in C
(``ruvector_b200/csrc/hnswio.cc``); ``level_csr(l)`` returns that level's adjacency with
rows indexed by ``origin_id`` (= ruvector-core's ``HnswIndex`` idx) and neighbours in the
stored (ascending distance) order -- the reference's own on-disk topology, no Rust needed.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

from . import _lib
from .graph import CsrGraph


class HnswDump:
    def __init__(self, graph_path):
        self._lib = _lib.load()
        h = C.c_void_p()
        _lib.check(self._lib.ruvgnn_hnsw_dump_open(os.fspath(graph_path).encode(), C.byref(h)))
        self.handle = h
        nb, nid, ep, dim = C.c_uint64(), C.c_uint64(), C.c_uint64(), C.c_uint64()
        nl, mc, mode = C.c_uint32(), C.c_uint32(), C.c_uint32()
        _lib.check(self._lib.ruvgnn_hnsw_dump_info(h, C.byref(nb), C.byref(nid), C.byref(nl), C.byref(mc),
                                                   C.byref(dim), C.byref(mode), C.byref(ep)))
        self.nb_point, self.n_ids, self.nb_layer = nb.value, nid.value, nl.value
        self.max_nb_connection, self.dimension = mc.value, dim.value
        self.full_mode, self.entry_point = mode.value == 1, ep.value

    def level_edges(self, level: int) -> int:
        e = C.c_uint64()
        _lib.check(self._lib.ruvgnn_hnsw_dump_level_edges(self.handle, level, C.byref(e)))
        return e.value

    def level_csr(self, level: int = 0, n_nodes: int | None = None) -> CsrGraph:
        n = self.n_ids if n_nodes is None else n_nodes
        e = self.level_edges(level)
        try:
            rp = np.empty(n + 1, dtype=np.uint64)
            ci = np.empty(e, dtype=np.uint32)
            ew = np.empty(e, dtype=np.float32)
        except (MemoryError, ValueError) as err:  # e.g. one stray origin_id of 2^31 in a small dump
            from .errors import InvalidInput

            raise InvalidInput(f"Invalid input: cannot hold a CSR of {n} rows / {e} edges: {err}") from None
        _lib.check(self._lib.ruvgnn_hnsw_dump_level_csr(
            self.handle, level, n, rp.ctypes.data_as(C.POINTER(C.c_uint64)),
            ci.ctypes.data_as(C.POINTER(C.c_uint32)), ew.ctypes.data_as(_lib.c_float_p)))
        return CsrGraph(rp, ci, ew)

    def close(self):
        if getattr(self, "handle", None):
            self._lib.ruvgnn_hnsw_dump_close(self.handle)
            self.handle = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
