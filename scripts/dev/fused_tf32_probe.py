"""Which stage of the fused K-upd goes wrong in single-pass TF32 mode?  (development probe)
RUVGNN_FUSED_DEBUG bits 8-10 select the intermediate the kernel writes: 1 a, 2 z, 3 rh, 4 d."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import ruvector_b200 as rv
from ruvector_b200 import synth

ctx = rv.default_context(0)
def mk(mode, seed=11):
    os.environ["RUVGNN_FUSED_UPD"] = "1"
    l = rv.RuvectorLayer(128, 128, 4, 0.1, weights=synth.make_weights(128, 128, seed=seed), ctx=ctx, mode=mode)
    l._device_layer()
    return l
n = 3000
x = synth.make_features(n, 128, seed=12)
g = synth.hnsw_like_graph(n, 16, seed=13)
dg = g.to_device(ctx)
ref = mk(rv.MODE_FP32_TC).forward_graph(x, dg)
out = mk(rv.MODE_TF32).forward_graph(x, dg)
nz = np.linalg.norm(ref, axis=1) > 0
rel = np.linalg.norm(out - ref, axis=1)[nz] / np.linalg.norm(ref, axis=1)[nz]
print(f"debug {os.environ.get('RUVGNN_FUSED_DEBUG')}: single vs 3-pass: max {np.nanmax(rel):.3e} mean {np.nanmean(rel):.3e} nan rows {int(np.isnan(out).any(axis=1).sum())}"
      f" bad rows {(rel > 2e-2).sum()} of {nz.sum()}; ref absmax {np.abs(ref).max():.3e}", flush=True)
