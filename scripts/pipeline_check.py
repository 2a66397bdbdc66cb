#!/usr/bin/env python
"""Runs forwards large enough for the row-block pipeline of phase B (K-upd of block b on a second stream
beside K-agg of block b+1) and saves sampled outputs.  tests/test_gpu_parity.py runs it with the pipeline
on (default) and off (RUVGNN_UPD_SMS=0) and compares bit for bit."""
import sys
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))

import ruvector_b200 as rv  # noqa: E402
from ruvector_b200 import synth  # noqa: E402


def main(path):
    stream = torch.cuda.Stream()
    torch.cuda.set_stream(stream)
    ctx = rv.Context(0, stream.cuda_stream)
    out = {}
    for tag, (n, d, heads, m) in {"r3": (300_001, 384, 8, 8), "r1": (400_003, 128, 4, 8)}.items():
        layer = rv.RuvectorLayer(d, d, heads, 0.1, weights=synth.make_weights(d, d, seed=n), ctx=ctx)
        x = synth.make_features(n, d, seed=n + 1)
        g = synth.hnsw_like_graph(n, m, seed=n + 2, iso_frac=0.02)
        dg = g.to_device(ctx)
        xd = torch.from_numpy(x).cuda()
        od = torch.empty(n, d, device="cuda")
        for _ in range(2):  # twice: the buffer halves are reused across calls
            od.fill_(float("nan"))
            layer.forward_graph_device(xd.data_ptr(), od.data_ptr(), dg, asynchronous=False)
        out[tag + "_dev"] = od[::7].cpu().numpy()
        out[tag + "_tail"] = od[-300:].cpu().numpy()
        out[tag + "_host"] = layer.forward_graph(x, dg)[::7]  # host-buffer path: pipelined inside each H2D block
    np.savez(path, **out)


if __name__ == "__main__":
    main(sys.argv[1])
