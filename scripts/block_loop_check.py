#!/usr/bin/env python
"""Runs a fixed set of forwards and saves the outputs (np.savez).  tests/test_gpu_parity.py runs it
twice -- default row-block size and RUVGNN_UPD_BLOCK=777 -- and compares the files bit for bit:
the K-agg -> K-upd block loop must not change a single bit."""
import sys
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))

import ruvector_b200 as rv  # noqa: E402
from ruvector_b200 import synth  # noqa: E402


def main(path):
    ctx = rv.Context(0)
    out = {}
    for tag, (n, d, heads, m, mode) in {"r3_tc": (5000, 384, 8, 16, rv.MODE_FP32_TC), "r1_tc": (4100, 128, 4, 8, rv.MODE_FP32_TC),
                                        "r3_simt": (3001, 96, 3, 8, rv.MODE_FP32_SIMT)}.items():
        layer = rv.RuvectorLayer(d, d, heads, 0.1, weights=synth.make_weights(d, d, seed=n), ctx=ctx, mode=mode)
        x = synth.make_features(n, d, seed=n + 1)
        g = synth.hnsw_like_graph(n, m, seed=n + 2, iso_frac=0.03)
        dg = g.to_device(ctx)
        out[tag + "_host"] = layer.forward_graph(x, dg, timing=True)  # host-buffer row-block pipeline
        xd = torch.from_numpy(x).cuda()
        od = torch.empty(n, d, device="cuda")
        layer.forward_graph_device(xd.data_ptr(), od.data_ptr(), dg, timing=True)
        ctx.sync()
        out[tag + "_dev"] = od.cpu().numpy()
        t = ctx.last_timing()
        assert t["aggregate_ms"] > 0 and t["update_ms"] > 0 and t["project_ms"] > 0, t
    # three stacked layers over three level graphs (config 3 shape)
    n, d = 6000, 128
    layers = [rv.RuvectorLayer(d, d, 4, 0.1, weights=synth.make_weights(d, d, seed=70 + l), ctx=ctx) for l in range(3)]
    graphs = [synth.hnsw_like_graph(n, 8, seed=80, level=l).to_device(ctx) for l in (2, 1, 0)]
    out["stack"] = rv.forward_stack(layers, synth.make_features(n, d, seed=9), graphs)
    np.savez(path, **out)


if __name__ == "__main__":
    main(sys.argv[1])
