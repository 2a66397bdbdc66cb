"""ctypes binding of libruvgnn.so (include/ruvgnn.h).  No fallback: a missing
library or a missing GPU raises."""
from __future__ import annotations

import ctypes as C
from pathlib import Path

from .errors import from_status

_LIB_PATH = Path(__file__).resolve().parent / "libruvgnn.so"
_lib = None

c_float_p = C.POINTER(C.c_float)


class Linear(C.Structure):
    _fields_ = [("w", c_float_p), ("b", c_float_p)]


LINEAR_NAMES = ("w_msg", "w_agg", "w_z", "u_z", "w_r", "u_r", "w_h", "u_h", "q", "k", "v", "out")


class Weights(C.Structure):
    """ruvgnn_weights (include/ruvgnn.h)."""

    _fields_ = [(n, Linear) for n in LINEAR_NAMES] + [("gamma", c_float_p), ("beta", c_float_p), ("eps", C.c_float)]


# every symbol include/ruvgnn.h declares: name -> (restype, argtypes)
_vp = C.c_void_p
_u64p = C.POINTER(C.c_uint64)
_u32p = C.POINTER(C.c_uint32)
SYMBOLS = {
    "ruvgnn_last_error": (C.c_char_p, []),
    "ruvgnn_abi_version": (C.c_uint32, []),
    "ruvgnn_ctx_create": (C.c_int32, [C.c_int32, _vp, C.POINTER(_vp)]),
    "ruvgnn_ctx_destroy": (None, [_vp]),
    "ruvgnn_sync": (C.c_int32, [_vp]),
    "ruvgnn_ctx_stream": (_vp, [_vp]),
    "ruvgnn_ctx_launch_count": (C.c_uint64, [_vp]),
    "ruvgnn_ctx_last_timing": (C.c_int32, [_vp, c_float_p]),
    "ruvgnn_ctx_set_timing": (C.c_int32, [_vp, C.c_int32]),
    "ruvgnn_ctx_set_sm_limit": (C.c_int32, [_vp, C.c_int32]),
    "ruvgnn_layer_create": (C.c_int32, [_vp, C.c_uint32, C.c_uint32, C.c_uint32, C.c_float, C.POINTER(Weights), C.POINTER(_vp)]),
    "ruvgnn_layer_validate": (C.c_int32, [C.c_uint32, C.c_uint32, C.c_uint32, C.c_float]),
    "ruvgnn_layer_from_json": (C.c_int32, [_vp, C.c_char_p, C.c_size_t, C.POINTER(_vp)]),
    "ruvgnn_layer_to_json": (C.c_int32, [_vp, C.c_char_p, C.c_size_t, C.POINTER(C.c_size_t)]),
    "ruvgnn_layer_json_normalize": (C.c_int32, [C.c_char_p, C.c_size_t, C.c_char_p, C.c_size_t, C.POINTER(C.c_size_t)]),
    "ruvgnn_layer_dims": (C.c_int32, [_vp, _u32p, _u32p, _u32p, c_float_p]),
    "ruvgnn_layer_set_mode": (C.c_int32, [_vp, C.c_uint32]),
    "ruvgnn_layer_mode": (C.c_uint32, [_vp]),
    "ruvgnn_layer_destroy": (None, [_vp]),
    "ruvgnn_pack_neighbours": (C.c_int32, [C.c_uint64, C.POINTER(_vp), _u64p, C.c_size_t, C.c_size_t, C.c_size_t, _u64p, _u32p, c_float_p]),
    "ruvgnn_flatten_levels": (C.c_int32, [C.c_uint64, C.c_uint32, C.POINTER(_vp), C.POINTER(_vp), C.POINTER(_vp), _u64p, _u32p, c_float_p]),
    "ruvgnn_bfs_order": (C.c_int32, [C.c_uint64, _u64p, _u32p, _u64p]),
    "ruvgnn_hnsw_dump_open": (C.c_int32, [C.c_char_p, C.POINTER(_vp)]),
    "ruvgnn_hnsw_dump_close": (None, [_vp]),
    "ruvgnn_hnsw_dump_info": (C.c_int32, [_vp, _u64p, _u64p, _u32p, _u32p, _u64p, _u32p, _u64p]),
    "ruvgnn_hnsw_dump_level_edges": (C.c_int32, [_vp, C.c_uint32, _u64p]),
    "ruvgnn_hnsw_dump_level_csr": (C.c_int32, [_vp, C.c_uint32, C.c_uint64, _u64p, _u32p, c_float_p]),
    "ruvgnn_graph_create": (C.c_int32, [_vp, C.c_uint64, C.c_uint64, _vp, _vp, _vp, C.c_uint32, C.POINTER(_vp)]),
    "ruvgnn_graph_dims": (C.c_int32, [_vp, _u64p, _u64p, _u64p]),
    "ruvgnn_graph_download": (C.c_int32, [_vp, _u64p, _u32p, c_float_p]),
    "ruvgnn_graph_destroy": (None, [_vp]),
    "ruvgnn_layer_forward": (C.c_int32, [_vp, _vp, _vp, _vp, _vp, C.c_uint32]),
    "ruvgnn_stack_forward": (C.c_int32, [_vp, C.POINTER(_vp), C.POINTER(_vp), C.c_uint32, _vp, _vp, C.c_uint32]),
    "ruvgnn_layer_forward_node": (C.c_int32, [_vp, _vp, c_float_p, C.c_uint32, c_float_p, C.c_uint32, c_float_p, C.c_uint32, c_float_p]),
    "ruvgnn_differentiable_search": (C.c_int32, [_vp, c_float_p, C.c_uint64, c_float_p, C.c_uint64, C.c_uint32, C.c_uint32, C.c_float, _u32p, c_float_p, _u32p, C.c_uint32]),
    "ruvgnn_hierarchical_forward": (C.c_int32, [_vp, c_float_p, C.c_uint64, C.c_uint32, C.POINTER(c_float_p), _u64p, _u32p, C.POINTER(_vp), C.c_uint32, c_float_p, C.c_uint64, _u32p]),
    "ruvgnn_layer_project": (C.c_int32, [_vp, _vp, C.c_uint64, _vp]),
    "ruvgnn_layer_project_rows": (C.c_int32, [_vp, _vp, C.c_uint64, _vp, C.c_uint64, C.c_uint64]),
    "ruvgnn_layer_halo_width": (C.c_uint32, [_vp]),
    "ruvgnn_layer_update_launches": (C.c_uint32, [_vp]),
    "ruvgnn_halo_pack": (C.c_int32, [_vp, _vp, _vp, C.c_uint64, _vp]),
    "ruvgnn_layer_messages": (C.c_int32, [_vp, _vp, C.c_uint64, C.c_uint64, _vp]),
    "ruvgnn_layer_expand_halo": (C.c_int32, [_vp, _vp, _vp, C.c_uint64, _vp]),
    "ruvgnn_layer_aggregate_update": (C.c_int32, [_vp, _vp, _vp, _vp, C.c_uint64, C.c_uint64, _vp]),
    "ruvgnn_store_describe_mmap": (C.c_int32, [C.c_char_p, C.c_uint32, C.c_uint64, _vp]),
    "ruvgnn_store_describe_feature_storage": (C.c_int32, [C.c_char_p, _vp]),
    "ruvgnn_layer_forward_store": (C.c_int32, [_vp, _vp, _vp, _vp, _vp, C.c_uint64]),
    "ruvgnn_layer_forward_split": (C.c_int32, [_vp, _vp, _vp, _vp, _vp, _u64p, C.c_uint32]),
    "ruvgnn_dist_unique_id": (C.c_int32, [C.POINTER(C.c_uint8)]),
    "ruvgnn_dist_create": (C.c_int32, [_vp, C.POINTER(C.c_uint8), C.c_int32, C.c_int32, C.POINTER(_vp)]),
    "ruvgnn_dist_create_all": (C.c_int32, [C.POINTER(_vp), C.c_int32, C.POINTER(_vp)]),
    "ruvgnn_dist_destroy": (None, [_vp]),
    "ruvgnn_dist_info": (C.c_int32, [_vp, C.POINTER(C.c_int32), C.POINTER(C.c_int32)]),
    "ruvgnn_dist_set_chunks": (C.c_int32, [_vp, C.c_uint32]),
    "ruvgnn_partition_create": (C.c_int32, [_vp, _u64p, _vp, _vp, _vp, C.c_uint32, C.POINTER(_vp)]),
    "ruvgnn_partition_destroy": (None, [_vp]),
    "ruvgnn_partition_stats": (C.c_int32, [_vp, _vp]),
    "ruvgnn_partition_download": (C.c_int32, [_vp, _u32p, _u32p, _u32p, _u64p, _u64p]),
    "ruvgnn_dist_forward": (C.c_int32, [_vp, _vp, _vp, _vp, _vp, C.c_uint32]),
    "ruvgnn_dist_last_stats": (C.c_int32, [_vp, _vp]),
}


class Store(C.Structure):
    """ruvgnn_store (include/ruvgnn.h)."""

    _fields_ = [("path", C.c_char_p), ("offset_bytes", C.c_uint64), ("row_stride_bytes", C.c_uint64),
                ("dim", C.c_uint32), ("rows", C.c_uint64)]


class PartitionInfo(C.Structure):
    """ruvgnn_partition_info (include/ruvgnn.h)."""

    _fields_ = [("n_local", C.c_uint64), ("n_total", C.c_uint64), ("n_edges", C.c_uint64), ("n_halo", C.c_uint64),
                ("n_send", C.c_uint64), ("remote_edges", C.c_uint64), ("interior_rows", C.c_uint64),
                ("dense", C.c_uint32), ("chunks", C.c_uint32)]


class DistStats(C.Structure):
    """ruvgnn_dist_stats (include/ruvgnn.h)."""

    _fields_ = [("bytes_sent", C.c_double), ("bytes_received", C.c_double), ("exposed_wait_ms", C.c_float),
                ("exchange_ms", C.c_float), ("passes", C.c_uint32), ("exact", C.c_uint32), ("peer_direct", C.c_uint32),
                ("reserved", C.c_uint32)]

FWD_DEVICE_PTRS, FWD_ASYNC, FWD_TIMING = 1, 2, 4
DIST_EXACT, DIST_MSG_HALO, DIST_PEER_DIRECT, DIST_OVERLAP, DIST_ROW_HALO = 16, 32, 64, 128, 256
PART_DEVICE_PTRS, PART_FORCE_SPARSE, PART_FORCE_DENSE = 1, 2, 4
NCCL_ID_BYTES = 128
GRAPH_PRENORMALIZED, GRAPH_DEVICE_PTRS, GRAPH_NO_ACTIVE_LIST = 1, 2, 4
MODE_FP32_SIMT, MODE_FP32_TC, MODE_TF32 = 0, 1, 2


def lib_path() -> Path:
    return _LIB_PATH


def load():
    """Loads libruvgnn.so, failing loudly when it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not _LIB_PATH.exists():
        raise ImportError(
            f"{_LIB_PATH} is missing: build it with `python -m ruvector_b200.build` "
            "(or __graft_entry__.build()). ruvector_b200 has no CPU fallback."
        )
    lib = C.CDLL(str(_LIB_PATH))
    for name, (res, args) in SYMBOLS.items():
        fn = getattr(lib, name)  # AttributeError if the .so does not export it
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def last_error() -> str:
    msg = load().ruvgnn_last_error()
    return msg.decode("utf-8", "replace") if msg else ""


def check(status: int):
    if status != 0:
        raise from_status(status, last_error())
