"""Fused K-upd (gemm_upd_fused.cu): out-proj + combine + w_agg + GRU + dropout scale + LayerNorm of an
H = 128 layer in ONE tcgen05 kernel (layer.rs:156-159, 407-417, 255-275, 68-81).  Same parity bound as
the unfused tensor-core path (1e-4 row-relative against the oracle); it folds W_agg into the
out-projection, so against the unfused path it agrees to rounding, not to the bit."""
import os

import numpy as np
import pytest

import ruvector_b200 as rv
from ruvector_b200 import _lib, synth

pytestmark = pytest.mark.gpu


def row_rel(out, ref):
    return np.linalg.norm(out - ref, axis=1) / np.maximum(np.linalg.norm(ref, axis=1), 1e-6)


def make_layer(d, heads, seed, ctx, fused, mode=None):
    old = os.environ.get("RUVGNN_FUSED_UPD")
    os.environ["RUVGNN_FUSED_UPD"] = "1" if fused else "0"
    try:
        layer = rv.RuvectorLayer(d, 128, heads, 0.1, weights=synth.make_weights(d, 128, seed=seed), ctx=ctx,
                                 mode=rv.MODE_FP32_TC if mode is None else mode)
        layer._device_layer()  # the knob is read when the device image is built
    finally:
        if old is None:
            os.environ.pop("RUVGNN_FUSED_UPD", None)
        else:
            os.environ["RUVGNN_FUSED_UPD"] = old
    return layer


@pytest.fixture(scope="module")
def ctx(built):
    return rv.default_context(0)


SHAPES = [
    (5000, 128, 4, 16),    # config 1/3/5 layer shape (R = 1 gather, K1 = 640): 39 tiles + an 8-row tail
    (40000, 128, 4, 8),    # 313 tiles over 148 persistent CTAs: barrier phases wrap, Y reused tile after tile
    (130, 128, 4, 4),      # one full tile and a 2-row tail
    (100, 128, 4, 4),      # less than one tile
    (3000, 64, 8, 16),     # 8 heads at H = 128: K|V|M gather (R = 3), K1 = 256, message rows with stride 4H
    (2000, 128, 2, 8),     # K1 = 384
    (1000, 256, 1, 8),     # K1 = 256 through the R = 1 gather, D != H
]


@pytest.mark.parametrize("n,d,heads,m", SHAPES)
def test_fused_update_matches_oracle_and_unfused(oracle, ctx, n, d, heads, m):
    fused = make_layer(d, heads, n + 1, ctx, True)
    plain = make_layer(d, heads, n + 1, ctx, False)
    assert fused.update_launches == 1 and plain.update_launches == 5
    x = synth.make_features(n, d, seed=n + 2)
    g = synth.hnsw_like_graph(n, m, seed=n + 3, iso_frac=0.05)
    dg = g.to_device(ctx)
    l0 = ctx.launch_count
    out = fused.forward_graph(x, dg)
    l1 = ctx.launch_count
    ref_dev = plain.forward_graph(x, dg)
    l2 = ctx.launch_count
    assert np.isfinite(out).all()
    ref = oracle.forward_graph(fused, x, g)
    rel = row_rel(out, ref)
    assert rel.max() <= 1e-4, f"fused K-upd row-relative error {rel.max():.3e}"
    assert np.abs(out - ref).max() <= 1e-4 * max(1.0, float(np.abs(ref).max()))
    assert row_rel(out, ref_dev).max() <= 2e-5
    # one launch instead of five per row block
    assert (l2 - l1) - (l1 - l0) == 4
    # deterministic
    assert np.array_equal(out, fused.forward_graph(x, dg))


def test_fused_update_no_neighbour_rows_take_layernorm_of_the_message(oracle, ctx):
    """layer.rs:385-389 inside the fused epilogue: a row without neighbours is LN(W_msg x + b)."""
    n, d, heads = 700, 128, 4
    fused = make_layer(d, heads, 5, ctx, True)
    x = synth.make_features(n, d, seed=6)
    g = synth.hnsw_like_graph(n, 8, seed=7, iso_frac=0.5)
    assert (g.degrees() == 0).sum() > 100
    out = fused.forward_graph(x, g.to_device(ctx))
    ref = oracle.forward_graph(fused, x, g)
    assert row_rel(out, ref).max() <= 1e-4
    iso = g.degrees() == 0
    assert row_rel(out[iso], ref[iso]).max() <= 1e-5


def test_fused_update_sparse_level_scatter(oracle, ctx):
    """Active-row form of an upper HNSW level: the fused kernel runs on the compacted rows and
    scatters its LayerNorm rows over the LN(m) rows; bits equal to the dense path with the same kernel."""
    import torch

    c2 = rv.Context(0)
    n, d, heads = 30000, 128, 4
    g = synth.hnsw_like_graph(n, 8, seed=108, level=1)
    x = synth.make_features(n, d, seed=5)
    fused = make_layer(d, heads, 9, c2, True)
    dense = fused.forward_graph(x, g.to_device(c2, flags=_lib.GRAPH_NO_ACTIVE_LIST))
    with torch.cuda.stream(c2.torch_stream()):
        xd = torch.from_numpy(x).cuda()
        od = torch.full((n, d), float("nan"), device="cuda")
        fused.forward_graph_device(xd.data_ptr(), od.data_ptr(), g.to_device(c2), asynchronous=False)
        sparse = od.cpu().numpy()
    assert np.array_equal(sparse, dense)
    assert row_rel(sparse, oracle.forward_graph(fused, x, g)).max() <= 1e-4


def test_fused_update_tf32_single_pass_own_tolerance(oracle, ctx):
    n, d, heads = 3000, 128, 4
    fused = make_layer(d, heads, 11, ctx, True, mode=rv.MODE_TF32)
    x = synth.make_features(n, d, seed=12)
    g = synth.hnsw_like_graph(n, 16, seed=13)
    out = fused.forward_graph(x, g.to_device(ctx))
    rel = row_rel(out, oracle.forward_graph(fused, x, g))
    assert np.isfinite(out).all() and rel.max() <= 2e-2


def test_fused_update_simt_mode_is_untouched(oracle, ctx):
    """FP32_SIMT stays the unfused on-device reference whatever the knob says."""
    n, d, heads = 500, 128, 4
    fused = make_layer(d, heads, 21, ctx, True)
    plain = make_layer(d, heads, 21, ctx, False)
    fused.set_mode(rv.MODE_FP32_SIMT)
    plain.set_mode(rv.MODE_FP32_SIMT)
    assert fused.update_launches == 5
    x = synth.make_features(n, d, seed=22)
    g = synth.hnsw_like_graph(n, 8, seed=23)
    dg = g.to_device(ctx)
    assert np.array_equal(fused.forward_graph(x, dg), plain.forward_graph(x, dg))
