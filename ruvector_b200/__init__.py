"""ruvector_b200 -- B200 (sm_100a) backend for ruvector-gnn's RuvectorLayer forward
over an HNSW neighbourhood graph.  Host-side mirror of the reference interface over
the C ABI in include/ruvgnn.h; all arithmetic runs in libruvgnn.so (no CPU fallback).
"""
from . import errors
from .errors import GnnError, LayerConfig, DimensionMismatch, InvalidInput, CudaError
from .graph import (CsrGraph, pack_neighbour_lists, pack_id_distance_lists, batch_to_graph, NEIGHBOUR_DTYPE,
                    locality_order, bfs_order, flatten_levels, permute_graph, inverse_permutation, remote_edge_fraction)
from .layer import Context, DeviceGraph, RuvectorLayer, default_context, forward_stack, forward_stack_device
from .search import differentiable_search, hierarchical_forward
from ._lib import MODE_FP32_SIMT, MODE_FP32_TC, MODE_TF32

__all__ = [
    "RuvectorLayer", "forward_stack", "forward_stack_device", "differentiable_search", "hierarchical_forward", "Context", "DeviceGraph", "CsrGraph", "default_context",
    "pack_neighbour_lists", "pack_id_distance_lists", "batch_to_graph", "NEIGHBOUR_DTYPE",
    "locality_order", "bfs_order", "flatten_levels", "permute_graph", "inverse_permutation", "remote_edge_fraction",
    "GnnError", "LayerConfig", "DimensionMismatch", "InvalidInput", "CudaError", "errors",
    "MODE_FP32_SIMT", "MODE_FP32_TC", "MODE_TF32",
]
