#!/usr/bin/env python
"""Small end-to-end runs of every kernel family for compute-sanitizer (memcheck):
  compute-sanitizer --tool memcheck python scripts/sanitize_small.py"""
import sys
from pathlib import Path

import numpy as np

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import ruvector_b200 as rv  # noqa: E402
from ruvector_b200 import synth  # noqa: E402

ctx = rv.default_context(0)
for (n, d, h, heads, m, mode) in [(300, 128, 128, 4, 8, rv.MODE_FP32_TC), (257, 384, 384, 8, 16, rv.MODE_FP32_TC),
                                  (200, 128, 256, 4, 8, rv.MODE_FP32_SIMT), (150, 20, 24, 3, 4, rv.MODE_FP32_SIMT),
                                  (130, 64, 128, 4, 4, rv.MODE_TF32)]:
    layer = rv.RuvectorLayer(d, h, heads, 0.1, weights=synth.make_weights(d, h, seed=n), ctx=ctx, mode=mode)
    x = synth.make_features(n, d, seed=n + 1)
    g = synth.hnsw_like_graph(n, m, seed=n + 2, iso_frac=0.05)
    out = layer.forward_graph(x, g.to_device(ctx))
    assert np.isfinite(out).all()
    one = layer.forward(x[0], list(x[1:4]), [0.2, 0.3, 0.5])
    assert np.isfinite(one).all()
idx, w = rv.differentiable_search(synth.make_features(3, 32, seed=1), synth.make_features(100, 32, seed=2), 5, 1.0, ctx=ctx)
l1 = rv.RuvectorLayer(32, 32, 2, 0.1, ctx=ctx, rng=np.random.default_rng(0))
o = rv.hierarchical_forward(synth.make_features(3, 32, seed=3), [synth.make_features(50, 32, seed=4)], [l1], ctx=ctx)
assert np.isfinite(o).all()
print("sanitize_small ok")
