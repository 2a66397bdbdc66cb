"""Host-side mirror of ``ruvector_gnn::layer::RuvectorLayer`` over libruvgnn.so.

Same constructor arguments, per-node ``forward`` signature, error behaviour and
serde-JSON weight format as the reference
(crates/ruvector-gnn/src/layer.rs:315-457; N-API surface
crates/ruvector-gnn-node/src/lib.rs:45-117), plus the batched
``forward_graph`` the reference's callers emulate with a per-node loop
(crates/ruvector-graph-transformer/src/verified_training.rs:422-436).

All arithmetic happens in the CUDA library; there is no CPU path here.
"""
from __future__ import annotations

import ctypes as C
import json
from typing import Optional, Sequence

import numpy as np

from . import _lib
from .errors import DimensionMismatch, InvalidInput, LayerConfig

F32 = np.float32
_fp = _lib.c_float_p


def _ptr(a: np.ndarray):
    return a.ctypes.data_as(_fp)


class Context:
    """``ruvgnn_ctx``: device, stream and activation workspace."""

    def __init__(self, device: int = 0, stream: Optional[int] = None):
        self._lib = _lib.load()
        h = C.c_void_p()
        _lib.check(self._lib.ruvgnn_ctx_create(device, C.c_void_p(stream or 0), C.byref(h)))
        self.handle = h
        self.device = device

    def sync(self):
        _lib.check(self._lib.ruvgnn_sync(self.handle))

    @property
    def stream(self) -> int:
        """The ``cudaStream_t`` every kernel of this ctx is enqueued on.  A ctx created without a
        stream owns a non-blocking one, which does NOT synchronise with torch's current stream:
        device buffers handed to the ``*_device`` entries must be ordered against THIS stream
        (``torch_stream()``), or the ctx must be created on the caller's stream."""
        return int(self._lib.ruvgnn_ctx_stream(self.handle) or 0)

    def torch_stream(self):
        """``torch.cuda.ExternalStream`` view of :attr:`stream` (``with torch.cuda.stream(ctx.torch_stream())``)."""
        import torch

        return torch.cuda.ExternalStream(self.stream, device=torch.device("cuda", self.device))

    @property
    def launch_count(self) -> int:
        return int(self._lib.ruvgnn_ctx_launch_count(self.handle))

    def last_timing(self) -> dict:
        ms = (C.c_float * 6)()
        _lib.check(self._lib.ruvgnn_ctx_last_timing(self.handle, ms))
        return dict(zip(("project_ms", "aggregate_ms", "update_ms", "h2d_ms", "d2h_ms", "exchange_ms"), map(float, ms)))

    def set_timing(self, on: bool):
        _lib.check(self._lib.ruvgnn_ctx_set_timing(self.handle, 1 if on else 0))

    def close(self):
        if getattr(self, "handle", None):
            self._lib.ruvgnn_ctx_destroy(self.handle)
            self.handle = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


_default_ctx: dict = {}


def default_context(device: int = 0) -> Context:
    if device not in _default_ctx:
        _default_ctx[device] = Context(device)
    return _default_ctx[device]


class DeviceGraph:
    """``ruvgnn_graph``: one HNSW level's CSR on the device."""

    def __init__(self, ctx: Context, row_ptr, col_idx, edge_w, n_cols: Optional[int] = None, flags: int = 0):
        self._lib = _lib.load()
        self.ctx = ctx
        h = C.c_void_p()
        if hasattr(row_ptr, "data_ptr"):
            # device tensors (torch): row_ptr int64/uint64 [n+1], col_idx int32/uint32 [E], edge_w float32 [E],
            # contiguous, produced on (or synchronised with) the ctx stream; the library copies them
            n, n_edges = int(row_ptr.numel()) - 1, int(col_idx.numel())
            if row_ptr.element_size() != 8 or col_idx.element_size() != 4 or edge_w.element_size() != 4:
                raise InvalidInput("device CSR needs 64-bit row_ptr, 32-bit col_idx and float32 edge_w")
            if int(edge_w.numel()) != n_edges or not (row_ptr.is_contiguous() and col_idx.is_contiguous() and edge_w.is_contiguous()):
                raise InvalidInput("len(col_idx) and len(edge_w) must agree and the tensors be contiguous")
            _lib.check(self._lib.ruvgnn_graph_create(
                ctx.handle, n, n if n_cols is None else n_cols, C.c_void_p(row_ptr.data_ptr()),
                C.c_void_p(col_idx.data_ptr()), C.c_void_p(edge_w.data_ptr()), flags | _lib.GRAPH_DEVICE_PTRS, C.byref(h)))
        else:
            row_ptr = np.ascontiguousarray(row_ptr, dtype=np.uint64)
            col_idx = np.ascontiguousarray(col_idx, dtype=np.uint32)
            edge_w = np.ascontiguousarray(edge_w, dtype=F32)
            if row_ptr.ndim != 1 or row_ptr.size < 1:
                raise InvalidInput("row_ptr must be a 1-d array of n_nodes + 1 entries")
            n, n_edges = row_ptr.size - 1, int(col_idx.size)
            if col_idx.size != edge_w.size or int(row_ptr[-1]) != col_idx.size:
                raise InvalidInput("row_ptr[-1], len(col_idx) and len(edge_w) must agree")
            _lib.check(self._lib.ruvgnn_graph_create(
                ctx.handle, n, n if n_cols is None else n_cols, C.c_void_p(row_ptr.ctypes.data),
                C.c_void_p(col_idx.ctypes.data), C.c_void_p(edge_w.ctypes.data), flags, C.byref(h)))
        self.handle = h
        self.n_nodes = n
        self.n_cols = n if n_cols is None else n_cols
        self.n_edges = n_edges

    def download(self):
        rp = np.empty(self.n_nodes + 1, dtype=np.uint64)
        ci = np.empty(self.n_edges, dtype=np.uint32)
        ew = np.empty(self.n_edges, dtype=F32)
        _lib.check(
            self._lib.ruvgnn_graph_download(
                self.handle, rp.ctypes.data_as(C.POINTER(C.c_uint64)),
                ci.ctypes.data_as(C.POINTER(C.c_uint32)), _ptr(ew),
            )
        )
        return rp, ci, ew

    def close(self):
        if getattr(self, "handle", None):
            self._lib.ruvgnn_graph_destroy(self.handle)
            self.handle = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


LINEARS = _lib.LINEAR_NAMES  # serde declaration order
_JSON_PATH = {  # weight name -> path inside the serde JSON
    "w_msg": ("w_msg",), "w_agg": ("w_agg",),
    "w_z": ("w_update", "w_z"), "u_z": ("w_update", "u_z"), "w_r": ("w_update", "w_r"),
    "u_r": ("w_update", "u_r"), "w_h": ("w_update", "w_h"), "u_h": ("w_update", "u_h"),
    "q": ("attention", "q_linear"), "k": ("attention", "k_linear"),
    "v": ("attention", "v_linear"), "out": ("attention", "out_linear"),
}


def _f32_list_json(a: np.ndarray) -> str:
    # shortest decimal that round-trips the f32 (what serde_json/ryu prints)
    return "[" + ",".join(np.ascontiguousarray(a, dtype=F32).ravel().astype("U32").tolist()) + "]"


def _ndarray_json(a: np.ndarray) -> str:
    return '{"v":1,"dim":[%s],"data":%s}' % (",".join(str(int(d)) for d in a.shape), _f32_list_json(a))


class RuvectorLayer:
    """``RuvectorLayer::new(input_dim, hidden_dim, heads, dropout)`` (layer.rs:347-368).

    Raises :class:`LayerConfig` with the reference's messages when ``dropout``
    is outside ``[0, 1]`` or ``hidden_dim % heads != 0``.  Weights are Xavier
    normal ``N(0, 2/(in+out))`` with zero biases, gamma=1, beta=0 as in the
    reference (layer.rs:21-34,59-64) unless ``weights`` injects them.
    """

    def __init__(self, input_dim: int, hidden_dim: int, heads: int, dropout: float, *,
                 weights: Optional[dict] = None, rng: Optional[np.random.Generator] = None,
                 ctx: Optional[Context] = None, mode: Optional[int] = None):
        lib = _lib.load()
        st = lib.ruvgnn_layer_validate(int(input_dim), int(hidden_dim), int(heads), float(dropout))
        if st != 0:
            _lib.check(st)
        self.input_dim, self.hidden_dim, self.heads = int(input_dim), int(hidden_dim), int(heads)
        self.dropout = float(np.float32(dropout))
        self._ctx = ctx
        self._handle = None
        self._mode = mode
        if weights is None:
            rng = rng or np.random.default_rng()
            weights = {}
            for name in LINEARS:
                fan_in = self.input_dim if name == "w_msg" else self.hidden_dim
                scale = np.sqrt(F32(2.0) / F32(fan_in + self.hidden_dim))
                w = rng.normal(0.0, float(scale), size=(self.hidden_dim, fan_in)).astype(F32)
                weights[name] = (w, np.zeros(self.hidden_dim, dtype=F32))
            weights["gamma"] = np.ones(self.hidden_dim, dtype=F32)
            weights["beta"] = np.zeros(self.hidden_dim, dtype=F32)
            weights["eps"] = 1e-5
        self.weights = self._check_weights(weights)

    # ---- weights -----------------------------------------------------------
    def _check_weights(self, w: dict) -> dict:
        out = {}
        for name in LINEARS:
            fan_in = self.input_dim if name == "w_msg" else self.hidden_dim
            W, b = w[name]
            W = np.ascontiguousarray(W, dtype=F32)
            b = np.ascontiguousarray(b, dtype=F32)
            if W.shape != (self.hidden_dim, fan_in) or b.shape != (self.hidden_dim,):
                raise DimensionMismatch(
                    f"Tensor dimension mismatch: expected {(self.hidden_dim, fan_in)}, got {W.shape} ({name})")
            out[name] = (W, b)
        out["gamma"] = np.ascontiguousarray(w["gamma"], dtype=F32)
        out["beta"] = np.ascontiguousarray(w["beta"], dtype=F32)
        if out["gamma"].shape != (self.hidden_dim,) or out["beta"].shape != (self.hidden_dim,):
            raise DimensionMismatch(f"Tensor dimension mismatch: expected {self.hidden_dim}, got {out['gamma'].shape}")
        out["eps"] = float(np.float32(w.get("eps", 1e-5)))
        return out

    def c_weights(self) -> "_lib.Weights":
        """The ``ruvgnn_weights`` view of this layer (pointers into the host arrays)."""
        cw = _lib.Weights()
        for name in LINEARS:
            W, b = self.weights[name]
            setattr(cw, name, _lib.Linear(_ptr(W), _ptr(b)))
        cw.gamma = _ptr(self.weights["gamma"])
        cw.beta = _ptr(self.weights["beta"])
        cw.eps = self.weights["eps"]
        return cw

    # ---- serde JSON (ruvector-gnn-node/src/lib.rs:97-117) --------------------
    def to_json(self) -> str:
        w = self.weights

        def lin(n):
            return '{"weights":%s,"bias":%s}' % (_ndarray_json(w[n][0]), _ndarray_json(w[n][1]))

        return (
            '{"w_msg":%s,"w_agg":%s,"w_update":{"w_z":%s,"u_z":%s,"w_r":%s,"u_r":%s,"w_h":%s,"u_h":%s},'
            '"attention":{"num_heads":%d,"head_dim":%d,"q_linear":%s,"k_linear":%s,"v_linear":%s,"out_linear":%s},'
            '"norm":{"gamma":%s,"beta":%s,"eps":%s},"dropout":%s}'
            % (lin("w_msg"), lin("w_agg"), lin("w_z"), lin("u_z"), lin("w_r"), lin("u_r"), lin("w_h"),
               lin("u_h"), self.heads, self.hidden_dim // self.heads, lin("q"), lin("k"), lin("v"), lin("out"),
               _ndarray_json(w["gamma"]), _ndarray_json(w["beta"]), str(np.float32(w["eps"])),
               str(np.float32(self.dropout)))
        )

    @classmethod
    def from_json(cls, text: str, *, ctx: Optional[Context] = None) -> "RuvectorLayer":
        try:
            doc = json.loads(text)

            def arr(node, ndim):
                dim = node["dim"]
                dim = [dim] if isinstance(dim, int) else list(dim)
                a = np.asarray(node["data"], dtype=np.float64).astype(F32)
                if len(dim) != ndim or int(np.prod(dim)) != a.size:
                    raise ValueError("data and dimension must match in size")
                return a.reshape(dim)

            weights = {}
            for name, path in _JSON_PATH.items():
                node = doc
                for p in path:
                    node = node[p]
                weights[name] = (arr(node["weights"], 2), arr(node["bias"], 1))
            weights["gamma"] = arr(doc["norm"]["gamma"], 1)
            weights["beta"] = arr(doc["norm"]["beta"], 1)
            weights["eps"] = float(doc["norm"]["eps"])
            heads = int(doc["attention"]["num_heads"])
            dropout = float(doc["dropout"])
            hidden, in_dim = weights["w_msg"][0].shape
        except (KeyError, TypeError, ValueError) as e:
            raise InvalidInput(f"Deserialization error: {e}") from e
        return cls(in_dim, hidden, heads, dropout, weights=weights, ctx=ctx)

    # ---- device ------------------------------------------------------------
    @property
    def ctx(self) -> Context:
        if self._ctx is None:
            self._ctx = default_context(0)
        return self._ctx

    def _device_layer(self):
        if self._handle is None:
            lib = _lib.load()
            h = C.c_void_p()
            cw = self.c_weights()
            _lib.check(lib.ruvgnn_layer_create(self.ctx.handle, self.input_dim, self.hidden_dim, self.heads,
                                               self.dropout, C.byref(cw), C.byref(h)))
            self._handle = h
            if self._mode is not None:  # None: the library's default (tensor cores where the shape allows)
                _lib.check(lib.ruvgnn_layer_set_mode(h, self._mode))
        return self._handle

    @property
    def mode(self) -> int:
        """The arithmetic mode the device layer runs in (``MODE_FP32_TC`` by default where the shape
        allows it, else ``MODE_FP32_SIMT``)."""
        return int(_lib.load().ruvgnn_layer_mode(self._device_layer()))

    @property
    def update_launches(self) -> int:
        """Kernel launches of the update phase per row block: 1 = the fused kernel (hidden_dim 128,
        tensor-core modes), 5 = four GEMMs + LayerNorm."""
        return int(_lib.load().ruvgnn_layer_update_launches(self._device_layer()))

    def set_mode(self, mode: int):
        self._mode = mode
        if self._handle is not None:
            _lib.check(_lib.load().ruvgnn_layer_set_mode(self._handle, mode))

    def close(self):
        if self._handle is not None:
            _lib.load().ruvgnn_layer_destroy(self._handle)
            self._handle = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ---- forward -----------------------------------------------------------
    def forward(self, node_embedding, neighbor_embeddings: Sequence, edge_weights) -> np.ndarray:
        """``forward(&self, node_embedding, neighbor_embeddings, edge_weights)`` (layer.rs:379-384)."""
        x = np.ascontiguousarray(node_embedding, dtype=F32).ravel()
        k = len(neighbor_embeddings)
        if k:
            nb = np.ascontiguousarray(np.stack([np.asarray(v, dtype=F32).ravel() for v in neighbor_embeddings]))
            if nb.shape[1] != x.size:
                raise DimensionMismatch(f"Tensor dimension mismatch: expected {x.size}, got {nb.shape[1]}")
        else:
            nb = np.zeros((0, x.size), dtype=F32)
        w = np.ascontiguousarray(edge_weights, dtype=F32).ravel()
        out = np.empty(self.hidden_dim, dtype=F32)
        _lib.check(_lib.load().ruvgnn_layer_forward_node(
            self.ctx.handle, self._device_layer(), _ptr(x), x.size, _ptr(nb), k, _ptr(w), w.size, _ptr(out)))
        return out

    def forward_graph(self, x, graph: DeviceGraph, out: Optional[np.ndarray] = None, timing: bool = False) -> np.ndarray:
        """Batched forward over every node of ``graph`` with host buffers (H2D/D2H inside)."""
        x = np.ascontiguousarray(x, dtype=F32)
        if x.ndim != 2 or x.shape[1] != self.input_dim:
            raise DimensionMismatch(f"Tensor dimension mismatch: expected {self.input_dim}, got {x.shape[-1]}")
        if x.shape[0] != graph.n_nodes:
            raise DimensionMismatch(f"Tensor dimension mismatch: expected {graph.n_nodes} rows, got {x.shape[0]}")
        if out is None:
            out = np.empty((x.shape[0], self.hidden_dim), dtype=F32)
        flags = _lib.FWD_TIMING if timing else 0
        _lib.check(_lib.load().ruvgnn_layer_forward(
            self.ctx.handle, self._device_layer(), graph.handle, C.c_void_p(x.ctypes.data),
            C.c_void_p(out.ctypes.data), flags))
        return out

    def forward_graph_device(self, x_ptr: int, out_ptr: int, graph: DeviceGraph, *, asynchronous: bool = True,
                             timing: bool = False):
        """Batched forward on device pointers (e.g. ``torch.Tensor.data_ptr()``), enqueue only.

        Stream contract: the kernels run on ``self.ctx.stream``.  The buffers must have been produced
        on (or synchronised with) that stream, and readers must wait for it -- create the ctx on the
        caller's stream (``Context(dev, torch.cuda.current_stream().cuda_stream)``) or work under
        ``with torch.cuda.stream(ctx.torch_stream())``."""
        flags = _lib.FWD_DEVICE_PTRS | (_lib.FWD_ASYNC if asynchronous else 0) | (_lib.FWD_TIMING if timing else 0)
        _lib.check(_lib.load().ruvgnn_layer_forward(
            self.ctx.handle, self._device_layer(), graph.handle, C.c_void_p(x_ptr), C.c_void_p(out_ptr), flags))


def forward_stack_device(layers: Sequence[RuvectorLayer], x_ptr: int, out_ptr: int, graphs: Sequence[DeviceGraph],
                         asynchronous: bool = True, timing: bool = False):
    """Stacked forward on device pointers (``ruvgnn_stack_forward``), enqueue only by default; same stream
    contract as :meth:`RuvectorLayer.forward_graph_device`.  ``timing``: one timing slot per layer."""
    if len(layers) != len(graphs) or not layers:
        raise InvalidInput("one graph per layer is required")
    n = len(layers)
    lh = (C.c_void_p * n)(*[l._device_layer() for l in layers])
    gh = (C.c_void_p * n)(*[g.handle for g in graphs])
    flags = _lib.FWD_DEVICE_PTRS | (_lib.FWD_ASYNC if asynchronous else 0) | (_lib.FWD_TIMING if timing else 0)
    _lib.check(_lib.load().ruvgnn_stack_forward(layers[0].ctx.handle, lh, gh, n, C.c_void_p(x_ptr), C.c_void_p(out_ptr), flags))


def forward_stack(layers: Sequence[RuvectorLayer], x, graphs: Sequence[DeviceGraph], out: Optional[np.ndarray] = None) -> np.ndarray:
    """Stacked forward, one layer per HNSW level graph, applied in the order given (config 3).

    Batched analogue of the per-level loop of ``hierarchical_forward``
    (crates/ruvector-gnn/src/search.rs:117-159): the output of layer l is the input of
    layer l+1, so ``layers[l+1].input_dim`` must equal ``layers[l].hidden_dim``.
    Host buffers in / out; intermediates stay on the device.
    """
    if len(layers) != len(graphs) or not layers:
        raise InvalidInput("one graph per layer is required")
    x = np.ascontiguousarray(x, dtype=F32)
    if x.ndim != 2 or x.shape[1] != layers[0].input_dim:
        raise DimensionMismatch(f"Tensor dimension mismatch: expected {layers[0].input_dim}, got {x.shape[-1]}")
    if x.shape[0] != graphs[0].n_nodes:
        raise DimensionMismatch(f"Tensor dimension mismatch: expected {graphs[0].n_nodes} rows, got {x.shape[0]}")
    ctx = layers[0].ctx
    if out is None:
        out = np.empty((x.shape[0], layers[-1].hidden_dim), dtype=F32)
    n = len(layers)
    lh = (C.c_void_p * n)(*[l._device_layer() for l in layers])
    gh = (C.c_void_p * n)(*[g.handle for g in graphs])
    _lib.check(_lib.load().ruvgnn_stack_forward(ctx.handle, lh, gh, n, C.c_void_p(x.ctypes.data),
                                                C.c_void_p(out.ctypes.data), 0))
    return out
