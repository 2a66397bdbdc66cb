"""tcgen05 projection path (RUVGNN_MODE_FP32_TC = 3xTF32, RUVGNN_MODE_TF32 = single pass).

FP32_TC is a parity mode: same 1e-4 bound as the SIMT path (expected error ~1e-6).
TF32 is NOT a parity mode; its own stated tolerance is 2e-2 row-relative (reported separately).
"""
import numpy as np
import pytest

import ruvector_b200 as rv
from ruvector_b200 import synth

pytestmark = pytest.mark.gpu
F32 = np.float32


def row_rel(out, ref):
    return np.linalg.norm(out - ref, axis=1) / np.maximum(np.linalg.norm(ref, axis=1), 1e-6)


@pytest.fixture(scope="module")
def ctx(built):
    return rv.default_context(0)


SHAPES = [
    (3000, 384, 384, 8, 32),   # config 2 layer shape; 3000 rows = 23 full tiles + a 56-row tail
    (5000, 128, 128, 4, 16),   # config 1/3/5 layer shape
    (1000, 768, 768, 8, 16),   # config 4 layer shape
    (700, 128, 256, 4, 8),     # D != H
    (300, 96, 96, 8, 6),       # tensor-core GEMMs with the generic gather (H % 128 != 0)
    (130, 64, 128, 4, 4),      # barely more than one row tile
    (64, 32, 32, 2, 4),        # less than one row tile, smallest k-block count
]


@pytest.mark.parametrize("n,d,hdim,heads,m", SHAPES)
def test_fp32_tc_matches_oracle(oracle, ctx, n, d, hdim, heads, m):
    w = synth.make_weights(d, hdim, seed=n + 7)
    layer = rv.RuvectorLayer(d, hdim, heads, 0.1, weights=w, ctx=ctx, mode=rv.MODE_FP32_TC)
    x = synth.make_features(n, d, seed=n + 8)
    g = synth.hnsw_like_graph(n, m, seed=n + 9, iso_frac=0.03)
    out = layer.forward_graph(x, g.to_device(ctx))
    ref = oracle.forward_graph(layer, x, g)
    rel = row_rel(out, ref)
    assert np.isfinite(out).all()
    assert rel.max() <= 1e-4, f"3xTF32 row-relative error {rel.max():.3e}"
    assert np.abs(out - ref).max() <= 1e-4 * max(1.0, float(np.abs(ref).max()))
    # and it agrees with the SIMT FP32 path of the same library far inside the tolerance
    layer.set_mode(rv.MODE_FP32_SIMT)
    simt = layer.forward_graph(x, g.to_device(ctx))
    assert row_rel(out, simt).max() <= 2e-5


@pytest.mark.parametrize("n,d,hdim,heads,m", SHAPES[:3])
def test_tf32_single_pass_own_tolerance(oracle, ctx, n, d, hdim, heads, m):
    w = synth.make_weights(d, hdim, seed=n + 17)
    layer = rv.RuvectorLayer(d, hdim, heads, 0.1, weights=w, ctx=ctx, mode=rv.MODE_TF32)
    x = synth.make_features(n, d, seed=n + 18)
    g = synth.hnsw_like_graph(n, m, seed=n + 19)
    out = layer.forward_graph(x, g.to_device(ctx))
    ref = oracle.forward_graph(layer, x, g)
    rel = row_rel(out, ref)
    print(f"TF32 single-pass row-relative error: max {rel.max():.2e}, mean {rel.mean():.2e}")
    assert np.isfinite(out).all() and rel.max() <= 2e-2


def test_tc_mode_is_refused_for_unsupported_shapes(ctx):
    layer = rv.RuvectorLayer(100, 60, 5, 0.1, ctx=ctx)
    layer._device_layer()
    with pytest.raises(rv.errors.InvalidShape):
        layer.set_mode(rv.MODE_FP32_TC)


def test_tc_large_row_count_and_determinism(ctx):
    """Many row tiles per CTA (persistent scheduler wrap-around, both accumulator buffers,
    all pipeline phases) and bit-identical repeat."""
    n, d = 60_000, 128
    layer = rv.RuvectorLayer(d, d, 4, 0.1, weights=synth.make_weights(d, d, seed=3), ctx=ctx, mode=rv.MODE_FP32_TC)
    x = synth.make_features(n, d, seed=4)
    g = synth.hnsw_like_graph(n, 16, seed=5)
    dg = g.to_device(ctx)
    a = layer.forward_graph(x, dg)
    b = layer.forward_graph(x, dg)
    np.testing.assert_array_equal(a, b)
    layer.set_mode(rv.MODE_FP32_SIMT)
    s = layer.forward_graph(x, dg)
    assert row_rel(a, s).max() <= 2e-5
