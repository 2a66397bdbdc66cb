"""Parity of the CUDA path (through the C ABI) against the CPU oracle.

Tolerance (BASELINE.json north_star): embeddings within 1e-4 relative in the FP32 path --
per row ||gpu-ref||2 / max(||ref||2, 1e-6) <= 1e-4 and max-abs <= 1e-4 * max(1, ||ref||inf)
(LayerNorm outputs are zero-mean, so element-wise relative error is ill-posed near 0).
CSR / indexing: bit-exact.
"""
import ctypes as C

import numpy as np
import pytest

import ruvector_b200 as rv
from ruvector_b200 import _lib, synth

pytestmark = pytest.mark.gpu
F32 = np.float32
TOL = 1e-4


def assert_parity(out, ref, tol=TOL):
    assert out.shape == ref.shape and np.isfinite(out).all()
    rel = np.linalg.norm(out - ref, axis=1) / np.maximum(np.linalg.norm(ref, axis=1), 1e-6)
    assert rel.max() <= tol, f"row-relative error {rel.max():.3e} > {tol}"
    assert np.abs(out - ref).max() <= tol * max(1.0, float(np.abs(ref).max()))
    return float(rel.max())


@pytest.fixture(scope="module")
def ctx(built):
    return rv.default_context(0)


def run_case(oracle, ctx, n, d, hdim, heads, m, seed, dropout=0.1, trivial=False, graph=None, mode=0):
    w = synth.make_weights(d, hdim, seed=seed, trivial=trivial)
    layer = rv.RuvectorLayer(d, hdim, heads, dropout, weights=w, ctx=ctx, mode=mode)
    x = synth.make_features(n, d, seed=seed + 1)
    g = graph if graph is not None else synth.hnsw_like_graph(n, m, seed=seed + 2, iso_frac=0.03)
    out = layer.forward_graph(x, g.to_device(ctx))
    ref = oracle.forward_graph(layer, x, g)
    return assert_parity(out, ref), layer, x, g, out


def test_config1_reference_cpu_case(oracle, ctx):
    """BASELINE config 1: 10k random 128-d vectors, HNSW M=16 (exact cosine kNN lists,
    distances as edge weights), 4 heads, 1 layer."""
    x = synth.make_features(10_000, 128)
    g = synth.knn_graph(x, 16)
    for trivial in (False, True):  # trivial = the reference's own initialisation
        layer = rv.RuvectorLayer(128, 128, 4, 0.1, weights=synth.make_weights(128, 128, trivial=trivial), ctx=ctx)
        out = layer.forward_graph(x, g.to_device(ctx))
        assert_parity(out, oracle.forward_graph(layer, x, g))


@pytest.mark.parametrize("n,d,hdim,heads,m", [
    (64, 4, 8, 2, 2),          # the reference's unit-test layer (layer.rs:513): generic kernels
    (500, 128, 256, 4, 8),     # D != H (gnn_performance_test.rs:28), fast path P=2
    (700, 48, 96, 8, 6),       # H % 128 != 0 -> generic path
    (900, 100, 60, 5, 5),      # nothing aligned at all
    (3000, 384, 384, 8, 32),   # config 2's layer shape (d_h = 48), P=3
    (1200, 768, 768, 8, 32),   # config 4's layer shape (d_h = 96), P=6
    (2000, 128, 128, 1, 16), (2000, 128, 128, 32, 16), (2000, 128, 128, 16, 16),
    (1500, 256, 512, 8, 16),   # P=4
    (600, 64, 1024, 16, 8),    # P=8
    (800, 128, 384, 6, 8),     # heads do not divide 32 -> generic path at H=384
])
def test_layer_shapes(oracle, ctx, n, d, hdim, heads, m):
    run_case(oracle, ctx, n, d, hdim, heads, m, seed=n + hdim)


def test_edge_cases_degree_and_weights(oracle, ctx):
    n, d = 600, 128
    rng = np.random.default_rng(7)
    # ragged degrees incl. 0, 1, > 32 (several warp chunks) and > 64
    deg = rng.choice([0, 1, 2, 31, 32, 33, 64, 65, 100], size=n)
    row_ptr = np.concatenate([[0], np.cumsum(deg)]).astype(np.uint64)
    cols = np.concatenate([rng.choice(n, size=k, replace=False) for k in deg]).astype(np.uint32)
    w = rng.random(cols.size).astype(F32)
    # negative / zero weight rows -> uniform rule (layer.rs:428-432)
    for i in range(0, n, 5):
        s, e = int(row_ptr[i]), int(row_ptr[i + 1])
        w[s:e] = -w[s:e] if i % 10 else 0.0
    g = rv.CsrGraph(row_ptr, cols, w)
    run_case(oracle, ctx, n, d, d, 4, 0, seed=31, graph=g)
    # duplicate neighbours and self loops are legal inputs for the layer
    cols2 = cols.copy()
    cols2[::3] = np.repeat(np.arange(n), deg)[::3].astype(np.uint32)
    run_case(oracle, ctx, n, d, d, 4, 0, seed=32, graph=rv.CsrGraph(row_ptr, cols2, np.abs(w) + F32(0.01)))


def test_empty_graph_and_single_node(oracle, ctx):
    n, d = 50, 128
    g = rv.CsrGraph(np.zeros(n + 1, np.uint64), np.zeros(0, np.uint32), np.zeros(0, F32))
    _, layer, x, _, out = run_case(oracle, ctx, n, d, d, 4, 0, seed=41, graph=g)
    assert np.abs(out.mean(axis=1) - layer.weights["beta"].mean()).max() < 0.2  # LN(m) rows
    g1 = rv.CsrGraph(np.zeros(2, np.uint64), np.zeros(0, np.uint32), np.zeros(0, F32))
    run_case(oracle, ctx, 1, d, d, 4, 0, seed=42, graph=g1)


@pytest.mark.parametrize("d,hdim,heads", [(4, 8, 2), (128, 128, 4), (128, 256, 4)])
def test_per_node_forward_matches_reference_semantics(oracle, ctx, d, hdim, heads):
    layer = rv.RuvectorLayer(d, hdim, heads, 0.1, weights=synth.make_weights(d, hdim, seed=51), ctx=ctx)
    x = synth.make_features(6, d, seed=52)
    cases = [
        (x[1:3], [0.3, 0.7]),
        (x[1:6], [0.1, 0.2, 0.3, 0.4, 0.5]),
        (x[1:1], []),                     # no neighbours: LN(W_msg x + b)  (layer.rs:385-389)
        (x[1:4], [0.2, 0.6]),             # fewer weights than neighbours: zip truncation (:438)
        (x[1:3], [0.2, 0.6, 0.9, 0.4]),   # more weights than neighbours: extra ones still normalise
        (x[1:4], []),                     # no weights: zero aggregation, attention still runs
        (x[1:4], [-1.0, -2.0, -0.5]),     # non-positive sum: uniform (:428-432)
    ]
    for nb, w in cases:
        out = layer.forward(x[0], list(nb), w)
        ref = oracle.forward_node(layer, x[0], nb, w)
        assert out.shape == (hdim,)
        assert_parity(out[None, :], ref[None, :])


def test_per_node_forward_dimension_mismatch(ctx):
    layer = rv.RuvectorLayer(8, 8, 2, 0.1, ctx=ctx)
    with pytest.raises(rv.DimensionMismatch) as e:
        layer.forward(np.ones(7, F32), [], [])
    assert "expected 8, got 7" in str(e.value)
    with pytest.raises(rv.DimensionMismatch):
        layer.forward(np.ones(8, F32), [np.ones(5, F32)], [1.0])


def test_json_roundtrip_same_forward_on_device(ctx):  # basic.test.js:54-73
    layer = rv.RuvectorLayer(4, 8, 2, 0.1, ctx=ctx, rng=np.random.default_rng(5))
    node = np.array([1, 2, 3, 4], dtype=F32)
    nb = [np.array([0.5, 1.0, 1.5, 2.0], F32), np.array([2.0, 3.0, 4.0, 5.0], F32)]
    a = layer.forward(node, nb, [0.3, 0.7])
    assert a.shape == (8,)
    # through the C serde path: to_json on the device layer, from_json into a new device layer
    lib = _lib.load()
    need = C.c_size_t()
    _lib.check(lib.ruvgnn_layer_to_json(layer._device_layer(), None, 0, C.byref(need)))
    buf = C.create_string_buffer(need.value + 1)
    _lib.check(lib.ruvgnn_layer_to_json(layer._device_layer(), buf, need.value + 1, C.byref(need)))
    h = C.c_void_p()
    _lib.check(lib.ruvgnn_layer_from_json(ctx.handle, buf.value, len(buf.value), C.byref(h)))
    out = np.empty(8, dtype=F32)
    x2 = np.ascontiguousarray(np.stack(nb))
    w = np.array([0.3, 0.7], dtype=F32)
    fp = _lib.c_float_p
    _lib.check(lib.ruvgnn_layer_forward_node(ctx.handle, h, node.ctypes.data_as(fp), 4, x2.ctypes.data_as(fp), 2,
                                             w.ctypes.data_as(fp), 2, out.ctypes.data_as(fp)))
    lib.ruvgnn_layer_destroy(h)
    np.testing.assert_array_equal(a, out)
    b = rv.RuvectorLayer.from_json(buf.value.decode(), ctx=ctx).forward(node, nb, [0.3, 0.7])
    np.testing.assert_array_equal(a, b)


def test_csr_upload_is_bit_exact_and_validated(ctx):
    g = synth.hnsw_like_graph(5000, 16, seed=61)
    dg = g.to_device(ctx)
    rp, ci, ew = dg.download()
    assert rp.tobytes() == g.row_ptr.tobytes() and ci.tobytes() == g.col_idx.tobytes()
    assert ew.tobytes() == g.edge_w.tobytes()
    bad = g.col_idx.copy()
    bad[123] = 5000
    with pytest.raises(rv.InvalidInput) as e:
        rv.DeviceGraph(ctx, g.row_ptr, bad, g.edge_w)
    assert "out of range" in str(e.value)
    rp2 = g.row_ptr.copy()
    rp2[10], rp2[11] = rp2[11], rp2[10]
    if rp2[10] != rp2[11]:
        with pytest.raises(rv.InvalidInput) as e:
            rv.DeviceGraph(ctx, rp2, g.col_idx, g.edge_w)
        assert "monotone" in str(e.value)


def test_device_pointer_entry_with_torch_tensors(oracle, ctx):
    import torch

    n, d = 4000, 128
    layer = rv.RuvectorLayer(d, d, 4, 0.1, weights=synth.make_weights(d, d, seed=71), ctx=ctx)
    x = synth.make_features(n, d, seed=72)
    g = synth.hnsw_like_graph(n, 16, seed=73)
    dg = g.to_device(ctx)
    xt = torch.from_numpy(x).cuda()
    ot = torch.empty(n, d, device="cuda")
    torch.cuda.synchronize()
    layer.forward_graph_device(xt.data_ptr(), ot.data_ptr(), dg, timing=True)
    ctx.sync()
    t = ctx.last_timing()
    assert t["aggregate_ms"] > 0 and t["project_ms"] > 0 and t["update_ms"] > 0
    ref = oracle.forward_graph(layer, x, g)
    assert_parity(ot.cpu().numpy(), ref)
    # deterministic: a second run is bit-identical
    o2 = torch.empty_like(ot)
    layer.forward_graph_device(xt.data_ptr(), o2.data_ptr(), dg)
    ctx.sync()
    assert torch.equal(ot, o2)
    with pytest.raises(rv.InvalidInput):
        layer.forward_graph_device(xt.data_ptr(), xt.data_ptr(), dg)  # out aliases x


def test_batch_request_equals_per_node_calls(oracle, ctx):  # gnn_cache.rs:325-377
    d = 128
    layer = rv.RuvectorLayer(d, d, 4, 0.1, weights=synth.make_weights(d, d, seed=81), ctx=ctx)
    x = synth.make_features(40, d, seed=82)
    ops = [(x[0], list(x[1:4]), [0.2, 0.3, 0.5]), (x[5], [], []), (x[6], list(x[7:20]), list(np.linspace(0.1, 1, 13)))]
    xb, g, q = rv.batch_to_graph(ops)
    out = layer.forward_graph(xb, g.to_device(ctx))[q]
    for o, (node, nb, w) in zip(out, ops):
        ref = oracle.forward_node(layer, node, np.asarray(nb, F32).reshape(-1, d), w)
        assert_parity(o[None, :], ref[None, :])


def test_large_graph_properties_without_the_oracle(ctx):
    """Size-independent properties at a size the oracle would not finish quickly:
    relabelling equivariance, isolated rows = LN(projection), finite zero-mean-ish rows."""
    n, d, heads = 200_000, 128, 4
    layer = rv.RuvectorLayer(d, d, heads, 0.1, weights=synth.make_weights(d, d, seed=91, trivial=True), ctx=ctx)
    x = synth.make_features(n, d, seed=92)
    g = synth.hnsw_like_graph(n, 16, seed=93)
    out = layer.forward_graph(x, g.to_device(ctx))
    assert np.isfinite(out).all()
    assert np.abs(out.mean(axis=1)).max() < 1e-4  # gamma=1, beta=0: LayerNorm rows are zero-mean
    assert np.abs(out.var(axis=1) - 1).max() < 1e-2
    # relabel nodes: new id = perm[old id]
    perm = np.random.default_rng(94).permutation(n)
    inv = np.argsort(perm)
    deg = g.degrees()
    new_deg = deg[inv]
    rp = np.concatenate([[0], np.cumsum(new_deg)]).astype(np.uint64)
    starts = g.row_ptr[:-1].astype(np.int64)[inv]
    take = np.repeat(starts - rp[:-1].astype(np.int64), new_deg) + np.arange(int(rp[-1]))
    g2 = rv.CsrGraph(rp, perm[g.col_idx[take]].astype(np.uint32), g.edge_w[take])
    out2 = layer.forward_graph(x[inv], g2.to_device(ctx))
    np.testing.assert_array_equal(out2, out[inv])  # same per-row arithmetic, same neighbour order
    iso = np.nonzero(deg == 0)[0][:50]
    m = x[iso] @ layer.weights["w_msg"][0].T
    ln = (m - m.mean(1, keepdims=True)) / np.sqrt(m.var(1, keepdims=True) + 1e-5)
    assert np.abs(out[iso] - ln).max() < 2e-4


def test_config3_three_stacked_layers_over_hnsw_levels(oracle, ctx):
    """BASELINE config 3 (scaled down): 3 RuvectorLayers with distinct weights, one per HNSW
    level, top-down (level 2, 1, 0).  Nodes absent from an upper level take the LN(W_msg x)
    branch there.  Checked against the oracle applied layer by layer."""
    n, d, heads, m = 20_000, 128, 4, 8
    x = synth.make_features(n, d, seed=301)
    levels = [2, 1, 0]
    graphs = [synth.hnsw_like_graph(n, m, seed=302, level=lv, iso_frac=0.01 if lv == 0 else 0.0) for lv in levels]
    assert graphs[0].n_edges < graphs[1].n_edges < graphs[2].n_edges
    layers = [rv.RuvectorLayer(d, d, heads, 0.1, weights=synth.make_weights(d, d, seed=310 + i), ctx=ctx)
              for i in range(3)]
    for mode in (rv.MODE_FP32_SIMT, rv.MODE_FP32_TC):
        for l in layers:
            l.set_mode(mode)
        out = rv.forward_stack(layers, x, [g.to_device(ctx) for g in graphs])
        ref = x
        for l, g in zip(layers, graphs):
            ref = oracle.forward_graph(l, ref, g)
        assert_parity(out, ref)


def test_stack_rejects_mismatched_dimensions(ctx):
    a = rv.RuvectorLayer(8, 16, 2, 0.1, ctx=ctx)
    b = rv.RuvectorLayer(8, 8, 2, 0.1, ctx=ctx)  # expects 16 inputs
    g = rv.CsrGraph(np.zeros(5, np.uint64), np.zeros(0, np.uint32), np.zeros(0, F32)).to_device(ctx)
    with pytest.raises(rv.DimensionMismatch) as e:
        rv.forward_stack([a, b], np.zeros((4, 8), F32), [g, g])
    assert "expected 16, got 8" in str(e.value)
    # widths may change between layers when they chain correctly (search.rs:232-241)
    c = rv.RuvectorLayer(16, 8, 2, 0.1, ctx=ctx)
    out = rv.forward_stack([a, c], np.ones((4, 8), F32), [g, g])
    assert out.shape == (4, 8) and np.isfinite(out).all()


def test_nine_million_nodes_64bit_offsets_sampled_against_oracle(oracle, ctx):
    """Config-3/5 scale on one GPU: 9M nodes x 128 puts the projected block beyond 2^32 floats,
    so every row offset in the kernels must be 64-bit.  A sample of rows (incl. the last ones)
    is checked against the oracle's per-node forward; all rows must be finite."""
    n, d, heads, m = 9_000_000, 128, 4, 8
    rng = np.random.default_rng(77)
    x = rng.standard_normal((n, d), dtype=F32)
    x *= F32(1.0 / np.sqrt(d))
    deg = rng.integers(m, 2 * m + 1, size=n, dtype=np.int64)
    deg[rng.random(n) < 0.01] = 0
    row_ptr = np.zeros(n + 1, dtype=np.uint64)
    np.cumsum(deg, out=row_ptr[1:].view(np.int64))
    e = int(row_ptr[-1])
    col = rng.integers(0, n, size=e, dtype=np.int64).astype(np.uint32)  # duplicates / self loops are legal
    w = rng.random(e, dtype=F32)
    g = rv.CsrGraph(row_ptr, col, w)
    layer = rv.RuvectorLayer(d, d, heads, 0.1, weights=synth.make_weights(d, d, seed=78), ctx=ctx,
                             mode=rv.MODE_FP32_TC)
    out = layer.forward_graph(x, g.to_device(ctx))
    assert out.shape == (n, d)
    assert np.isfinite(out[::997]).all() and np.isfinite(out[-4096:]).all()
    sample = np.concatenate([rng.integers(0, n, size=60), np.arange(n - 8, n), np.arange(8),
                             np.nonzero(deg == 0)[0][:4]])
    for i in sample:
        s, t = int(row_ptr[i]), int(row_ptr[i + 1])
        ref = oracle.forward_node(layer, x[i], x[col[s:t]], w[s:t])
        rel = np.linalg.norm(out[i] - ref) / max(np.linalg.norm(ref), 1e-6)
        assert rel <= 1e-4, (int(i), rel)


def test_config2_full_size_sampled_against_oracle(oracle, ctx):
    """BASELINE config 2 at full size (the bench workload): 1M nodes, d=384, HNSW M=32
    (level-0 degree U{32..64}), 8 heads, tcgen05 3xTF32 projections.  Rows sampled uniformly,
    plus the first/last rows and isolated rows, are checked against the oracle's per-node
    forward (the reference's own call pattern)."""
    n, d, heads, m = 1_000_000, 384, 8, 32
    x = synth.make_features(n, d)
    g = synth.hnsw_like_graph(n, m)
    layer = rv.RuvectorLayer(d, d, heads, 0.1, weights=synth.make_weights(d, d), ctx=ctx, mode=rv.MODE_FP32_TC)
    out = layer.forward_graph(x, g.to_device(ctx))
    assert np.isfinite(out[::101]).all()
    deg = g.degrees()
    rng = np.random.default_rng(5)
    sample = np.concatenate([rng.integers(0, n, size=40), np.arange(4), np.arange(n - 4, n),
                             np.nonzero(deg == 0)[0][:3], [int(np.argmax(deg))]])
    worst = 0.0
    for i in sample:
        s, t = int(g.row_ptr[i]), int(g.row_ptr[i + 1])
        ref = oracle.forward_node(layer, x[i], x[g.col_idx[s:t]], g.edge_w[s:t])
        worst = max(worst, float(np.linalg.norm(out[i] - ref) / max(np.linalg.norm(ref), 1e-6)))
    assert worst <= 1e-4, worst


def test_seeded_fuzz_over_shapes_degrees_and_modes(oracle, ctx):
    """Random layer shapes (aligned and unaligned), head counts, degree distributions (incl.
    empty rows, hubs), weight signs and dropout values; every arithmetic mode the shape admits."""
    rng = np.random.default_rng(20261018)
    for case in range(24):
        heads = int(rng.choice([1, 2, 3, 4, 8]))
        if case % 3 == 0:   # tensor-core eligible, fast gather layouts
            hdim = int(rng.choice([128, 256, 384])); heads = int(rng.choice([1, 2, 4, 8]))
            d = int(rng.choice([32, 64, 128, 384]))
        elif case % 3 == 1:  # tensor-core eligible, generic gather
            hdim = int(rng.choice([32, 64, 96, 160])); heads = int(rng.choice([h for h in (1, 2, 4, 8) if hdim % h == 0]))
            d = int(rng.choice([32, 96]))
        else:                # nothing aligned
            hdim = heads * int(rng.integers(1, 12)); d = int(rng.integers(1, 70))
        n = int(rng.integers(1, 700))
        deg = rng.choice([0, 1, 2, 5, 17, 33, 70], size=n, p=[.1, .15, .2, .25, .15, .1, .05])
        deg = np.minimum(deg, 4 * n)
        row_ptr = np.concatenate([[0], np.cumsum(deg)]).astype(np.uint64)
        cols = rng.integers(0, n, size=int(row_ptr[-1])).astype(np.uint32)
        w = (rng.random(cols.size) * rng.choice([1.0, 1.0, -1.0])).astype(F32)
        g = rv.CsrGraph(row_ptr, cols, w)
        dropout = float(rng.choice([0.0, 0.1, 0.5, 1.0]))
        layer = rv.RuvectorLayer(d, hdim, heads, dropout, weights=synth.make_weights(d, hdim, seed=1000 + case), ctx=ctx)
        x = (rng.standard_normal((n, d)) * rng.choice([0.1, 1.0, 3.0])).astype(F32)
        ref = oracle.forward_graph(layer, x, g)
        dg = g.to_device(ctx)
        modes = [rv.MODE_FP32_SIMT] + ([rv.MODE_FP32_TC] if d % 32 == 0 and hdim % 32 == 0 else [])
        for mode in modes:
            layer.set_mode(mode)
            out = layer.forward_graph(x, dg)
            if dropout == 1.0:  # everything scales to 0 before LayerNorm: output = beta (:414-417)
                assert np.isfinite(out).all()
                continue
            rel = np.linalg.norm(out - ref, axis=1) / np.maximum(np.linalg.norm(ref, axis=1), 1e-6)
            assert rel.max() <= 1e-4, (case, n, d, hdim, heads, mode, float(rel.max()))


def test_row_block_loop_changes_no_bit(tmp_path):
    """K-agg -> K-upd runs per block of at most 2^20 rows on block-sized buffers; with a 777-row block
    (ragged last block, many blocks) every output must equal the single-block run bit for bit."""
    import os
    import subprocess
    import sys
    from pathlib import Path

    root = Path(__file__).resolve().parents[1]
    outs = []
    for block in (None, "777"):
        env = dict(os.environ)
        env.pop("RUVGNN_UPD_BLOCK", None)
        if block:
            env["RUVGNN_UPD_BLOCK"] = block
        path = tmp_path / f"out_{block}.npz"
        res = subprocess.run([sys.executable, str(root / "scripts" / "block_loop_check.py"), str(path)], env=env,
                             capture_output=True, text=True, timeout=600)
        assert res.returncode == 0, res.stderr[-2000:]
        outs.append(np.load(path))
    assert set(outs[0].files) == set(outs[1].files) and len(outs[0].files) == 7
    for k in outs[0].files:
        assert np.isfinite(outs[0][k]).all()
        assert np.array_equal(outs[0][k], outs[1][k]), k


def test_locality_permutation_permutes_the_outputs_bit_exactly(ctx):
    """graph.locality_order / permute_graph relabel nodes only: the device output of the permuted
    problem is the original output, permuted, bit for bit (rows keep their neighbour order)."""
    n, d, heads = 20000, 128, 4
    layer = rv.RuvectorLayer(d, d, heads, 0.1, weights=synth.make_weights(d, d, seed=31), ctx=ctx,
                             mode=rv.MODE_FP32_TC)
    x = synth.make_features(n, d, seed=32)
    g = synth.hnsw_like_graph(n, 8, seed=33, locality=0.7, window=200)
    perm = rv.locality_order(g)
    ref = layer.forward_graph(x, g.to_device(ctx))
    got = layer.forward_graph(np.ascontiguousarray(x[perm]), rv.permute_graph(g, perm).to_device(ctx))
    assert np.array_equal(got, ref[perm])
    out = np.empty_like(got)
    out[perm] = got  # == got[inverse_permutation(perm)]: back in the caller's node order
    assert np.array_equal(out, ref) and np.array_equal(out, got[rv.inverse_permutation(perm)])

