"""Runs one fused forward in the given mode under RUVGNN_FUSED_DEBUG and reports finiteness (development probe)."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import ruvector_b200 as rv
from ruvector_b200 import synth
mode = rv.MODE_TF32 if sys.argv[1] == "tf32" else rv.MODE_FP32_TC
ctx = rv.default_context(0)
os.environ["RUVGNN_FUSED_UPD"] = "1"
l = rv.RuvectorLayer(128, 128, 4, 0.1, weights=synth.make_weights(128, 128, seed=11), ctx=ctx, mode=mode)
n = 3000
x = synth.make_features(n, 128, seed=12)
g = synth.hnsw_like_graph(n, 16, seed=13)
try:
    out = l.forward_graph(x, g.to_device(ctx))
    print(sys.argv[1], "debug", os.environ.get("RUVGNN_FUSED_DEBUG"), "nan rows", int(np.isnan(out).any(axis=1).sum()), "absmax", float(np.nanmax(np.abs(out))))
except Exception as e:
    print(sys.argv[1], "debug", os.environ.get("RUVGNN_FUSED_DEBUG"), "ERROR", str(e)[:120])
