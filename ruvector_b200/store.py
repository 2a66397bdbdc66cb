"""Embedding store files as the X source / output sink of the layer.

Two on-disk layouts of the reference are mapped without conversion:

* ``MmapManager`` (crates/ruvector-gnn/src/mmap.rs:118-186): headerless dense row-major
  ``[max_nodes][d_embed]`` f32; node i starts at byte ``i * d_embed * 4`` (:199-203).
* cold-tier ``FeatureStorage`` (crates/ruvector-gnn/src/cold_tier.rs:17-18,45-75): 24-byte
  little-endian header ``dim u64, num_nodes u64, block_size u64`` followed by one
  page-aligned block of ``block_size`` bytes per node, the first ``dim * 4`` of which hold
  the features (:120-149).

Two ways to run the layer on a store:

* :func:`forward_store_files` -- the product path: ``ruvgnn_layer_forward_store`` streams the rows
  file -> pinned block -> device block -> projection and the new embeddings back to the output
  file, double-buffered, so neither X nor the output is ever resident as a whole (host or HBM);
* :func:`forward_from_store` -- convenience for small stores: ``numpy.memmap`` views handed to
  ``RuvectorLayer.forward_graph``.

Host-side byte work only; nothing here computes the layer.
"""
from __future__ import annotations

import mmap as _mmap
import os
import struct

import numpy as np

from .errors import DimensionMismatch, InvalidInput

HEADER_SIZE = 24


def open_mmap_embeddings(path, d_embed: int, max_nodes: int, mode: str = "r+") -> np.memmap:
    """``MmapManager::new(path, d_embed, max_nodes)``: creates/extends the file to
    ``max_nodes * d_embed * 4`` bytes and maps it as ``[max_nodes, d_embed]`` f32."""
    if d_embed <= 0:
        raise InvalidInput("d_embed must be > 0")
    size = max_nodes * d_embed * 4
    if mode != "r":
        with open(path, "ab"):
            pass
        if os.path.getsize(path) != size:
            with open(path, "r+b") as f:
                f.truncate(size)
    elif os.path.getsize(path) < size:
        raise InvalidInput(f"mmap file holds fewer than {max_nodes} x {d_embed} f32 values")
    if max_nodes == 0:
        return np.zeros((0, d_embed), dtype=np.float32)
    return np.memmap(path, dtype=np.float32, mode=mode, shape=(max_nodes, d_embed))


def create_feature_storage(path, dim: int, num_nodes: int) -> np.ndarray:
    """``FeatureStorage::create``: writes the header, sizes the file, returns the strided view."""
    if dim <= 0:
        raise InvalidInput("dim must be > 0")
    page = _mmap.PAGESIZE
    block = (dim * 4 + page - 1) // page * page
    with open(path, "wb") as f:
        f.write(struct.pack("<QQQ", dim, num_nodes, block))
        f.truncate(HEADER_SIZE + num_nodes * block)
    return open_feature_storage(path, mode="r+")


def open_feature_storage(path, mode: str = "r") -> np.ndarray:
    """``FeatureStorage::open``: a ``[num_nodes, dim]`` f32 view with row stride ``block_size``."""
    with open(path, "rb") as f:
        head = f.read(HEADER_SIZE)
    if len(head) != HEADER_SIZE:
        raise InvalidInput("feature storage file is shorter than its header")
    dim, num_nodes, block = struct.unpack("<QQQ", head)
    if dim == 0 or block < dim * 4 or block % 4:
        raise InvalidInput("feature storage header is inconsistent")
    if os.path.getsize(path) < HEADER_SIZE + num_nodes * block:
        raise InvalidInput("feature storage file is truncated")
    if num_nodes == 0:
        return np.zeros((0, dim), dtype=np.float32)
    raw = np.memmap(path, dtype=np.uint8, mode=mode, offset=HEADER_SIZE, shape=(num_nodes * block,))
    rows = raw.view(np.float32).reshape(num_nodes, block // 4)
    return rows[:, :dim]


def describe_mmap(path, d_embed: int, max_nodes: int):
    """``ruvgnn_store`` descriptor of a ``MmapManager`` file (size checked by the library)."""
    import ctypes as C

    from . import _lib

    st = _lib.Store()
    st._path = os.fsencode(str(path))
    _lib.check(_lib.load().ruvgnn_store_describe_mmap(st._path, int(d_embed), int(max_nodes), C.byref(st)))
    return st


def describe_feature_storage(path):
    """``ruvgnn_store`` descriptor of a cold-tier ``FeatureStorage`` file (header read by the library)."""
    import ctypes as C

    from . import _lib

    st = _lib.Store()
    st._path = os.fsencode(str(path))
    _lib.check(_lib.load().ruvgnn_store_describe_feature_storage(st._path, C.byref(st)))
    return st


def forward_store_files(layer, graph, x_store, out_store, block_rows: int = 0):
    """``ruvgnn_layer_forward_store``: X streamed from ``x_store``, the new embeddings streamed into
    ``out_store`` (descriptors from :func:`describe_mmap` / :func:`describe_feature_storage`; distinct
    files).  Blocks until the output file holds the result."""
    import ctypes as C

    from . import _lib

    _lib.check(_lib.load().ruvgnn_layer_forward_store(layer.ctx.handle, layer._device_layer(), graph.handle,
                                                      C.byref(x_store), C.byref(out_store), int(block_rows)))


def forward_from_store(layer, x_store: np.ndarray, graph, out_store: np.ndarray | None = None) -> np.ndarray:
    """Runs the batched forward with X read from a store view and, optionally, the new
    embeddings written back into another store view (dense or page-strided)."""
    if x_store.shape[1] != layer.input_dim:
        raise DimensionMismatch(f"Tensor dimension mismatch: expected {layer.input_dim}, got {x_store.shape[1]}")
    x = np.ascontiguousarray(x_store[: graph.n_nodes], dtype=np.float32)  # packs page-strided rows
    out = layer.forward_graph(x, graph)
    if out_store is not None:
        if out_store.shape[1] != layer.hidden_dim or out_store.shape[0] < out.shape[0]:
            raise DimensionMismatch(
                f"Tensor dimension mismatch: expected {(out.shape[0], layer.hidden_dim)}, got {out_store.shape}")
        out_store[: out.shape[0]] = out
        base = out_store
        while base is not None and not isinstance(base, np.memmap):
            base = getattr(base, "base", None)
        if base is not None:
            base.flush()
    return out
