"""Multi-pass gather (the overlapped form of the partitioned forward) on ONE GPU, and the
in-library partition plan against its NumPy model.

ruvgnn_layer_forward_split runs the layer with every node's neighbour list processed in several
passes over column ranges, the online-softmax state carried through HBM between passes -- the very
kernels ruvgnn_dist_forward launches per halo chunk.  It must equal the single-pass forward up to
f32 summation order (bound 2e-5 row-relative, measured ~1e-6) and the oracle within the 1e-4
parity bound (layer.rs:174-210, 421-445)."""
import ctypes as C

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def row_rel(a, b):
    return float((np.linalg.norm(a - b, axis=1) / np.maximum(np.linalg.norm(b, axis=1), 1e-6)).max())


def _torch_ctx():
    import torch

    import ruvector_b200 as rv

    stream = torch.cuda.Stream()
    torch.cuda.set_stream(stream)
    return rv.Context(0, stream.cuda_stream), stream


def _split_forward(ctx, layer, dg, xd, od, splits):
    from ruvector_b200 import _lib

    arr = np.asarray(splits, dtype=np.uint64)
    _lib.check(_lib.load().ruvgnn_layer_forward_split(
        ctx.handle, layer._device_layer(), dg.handle, C.c_void_p(xd.data_ptr()), C.c_void_p(od.data_ptr()),
        arr.ctypes.data_as(C.POINTER(C.c_uint64)), arr.size))


@pytest.mark.parametrize("d,heads", [(128, 4), (128, 2), (128, 1), (384, 8), (256, 8), (768, 8), (512, 4)])
def test_split_passes_equal_the_single_pass(built, oracle, d, heads):
    import torch

    import ruvector_b200 as rv
    from ruvector_b200 import synth

    ctx, _ = _torch_ctx()
    n, m = 3000, 16
    g = synth.hnsw_like_graph(n, m, seed=d + heads, iso_frac=0.03)
    # rows of degree 1, 33, 65, 100 (more than one 32-edge chunk, ragged tails)
    deg = g.degrees()
    assert deg.max() >= 17
    x = synth.make_features(n, d, seed=d)
    layer = rv.RuvectorLayer(d, d, heads, 0.1, weights=synth.make_weights(d, d, seed=7 * d), ctx=ctx)
    dg = g.to_device(ctx)
    xd = torch.from_numpy(x).cuda()
    single = torch.empty(n, d, device="cuda")
    layer.forward_graph_device(xd.data_ptr(), single.data_ptr(), dg)
    ctx.sync()
    single = single.cpu().numpy()
    ref = oracle.forward_graph(layer, x, g)
    assert row_rel(single, ref) <= 1e-4
    for splits in ([], [n // 2], [n // 8, n // 2, n // 2, n - 5], [0, n], [1, 2, 3, 4, 5, 6, 7]):
        od = torch.full((n, d), float("nan"), device="cuda")
        _split_forward(ctx, layer, dg, xd, od, splits)
        got = od.cpu().numpy()
        assert np.isfinite(got).all(), splits
        if not splits:
            assert np.array_equal(got, single)  # one pass over everything IS the single-pass kernel
        assert row_rel(got, single) <= 2e-5, (splits, row_rel(got, single))
        assert row_rel(got, ref) <= 1e-4, splits
    torch.cuda.set_stream(torch.cuda.default_stream())


def test_split_passes_on_ragged_degrees_and_nonpositive_weights(built, oracle):
    """Degrees 0, 1, 31, 32, 33, 65, 100; a row whose weights sum to <= 0 (uniform rule, layer.rs:426-432)
    must take the same rule in every pass (the sum runs over ALL of the row's edges, not the pass's)."""
    import torch

    import ruvector_b200 as rv
    from ruvector_b200 import synth
    from ruvector_b200.graph import CsrGraph

    ctx, _ = _torch_ctx()
    d, heads, n = 128, 4, 400
    rng = np.random.default_rng(5)
    degs = np.array([0, 1, 31, 32, 33, 65, 100, 7] * (n // 8))
    rp = np.zeros(n + 1, np.uint64)
    rp[1:] = np.cumsum(degs)
    cols = np.concatenate([rng.choice(n, size=k, replace=False) for k in degs]).astype(np.uint32)
    w = rng.random(cols.size).astype(np.float32)
    w[int(rp[5]):int(rp[6])] *= -1.0  # node 5: negative sum -> uniform weights
    g = CsrGraph(rp, cols, w)
    x = synth.make_features(n, d, seed=11)
    layer = rv.RuvectorLayer(d, d, heads, 0.1, weights=synth.make_weights(d, d, seed=3), ctx=ctx)
    ref = oracle.forward_graph(layer, x, g)
    dg = g.to_device(ctx)
    xd = torch.from_numpy(x).cuda()
    for splits in ([n // 3], [10, 200, 390]):
        od = torch.empty(n, d, device="cuda")
        _split_forward(ctx, layer, dg, xd, od, splits)
        assert row_rel(od.cpu().numpy(), ref) <= 1e-4
    torch.cuda.set_stream(torch.cuda.default_stream())


def test_split_rejects_generic_layouts_and_bad_splits(built):
    import torch

    import ruvector_b200 as rv
    from ruvector_b200 import synth

    ctx, _ = _torch_ctx()
    g = synth.hnsw_like_graph(64, 4, seed=1)
    dg = g.to_device(ctx)
    xd = torch.zeros(64, 96, device="cuda")
    od = torch.zeros(64, 96, device="cuda")
    generic = rv.RuvectorLayer(96, 96, 3, 0.1, weights=synth.make_weights(96, 96), ctx=ctx)
    with pytest.raises(rv.errors.InvalidShape):
        _split_forward(ctx, generic, dg, xd, od, [10])
    fast = rv.RuvectorLayer(128, 128, 4, 0.1, weights=synth.make_weights(128, 128), ctx=ctx)
    xd, od = torch.zeros(64, 128, device="cuda"), torch.zeros(64, 128, device="cuda")
    with pytest.raises(rv.InvalidInput):
        _split_forward(ctx, fast, dg, xd, od, [20, 10])
    with pytest.raises(rv.InvalidInput):
        _split_forward(ctx, fast, dg, xd, od, [65])
    torch.cuda.set_stream(torch.cuda.default_stream())


@pytest.mark.parametrize("world", [2, 3, 8])
@pytest.mark.parametrize("locality", [0.0, 0.9])
def test_device_plan_equals_the_numpy_model(built, world, locality):
    """ruvgnn_partition_create on a plan-only dist (no communicator): the de-duplicated halo, the
    renamed columns (stored order kept) and the per-peer counts of every rank against
    ruvector_b200.dist.HaloPlan, for both data planes."""
    import ruvector_b200 as rv
    from ruvector_b200 import dist as rvdist, synth
    from ruvector_b200.graph import CsrGraph

    ctx = rv.Context(0)
    n = 2400
    g = synth.hnsw_like_graph(n, 8, seed=world, locality=locality, window=100, iso_frac=0.05)
    bounds = rvdist.even_bounds(n, world)
    for rank in range(world):
        lo, hi = int(bounds[rank]), int(bounds[rank + 1])
        s, e = int(g.row_ptr[lo]), int(g.row_ptr[hi])
        rows = CsrGraph((g.row_ptr[lo:hi + 1] - g.row_ptr[lo]).astype(np.uint64), g.col_idx[s:e].copy(), g.edge_w[s:e].copy())
        model = rvdist.HaloPlan.build(rows, rank, bounds)
        d_ = rvdist.Dist(ctx, rank, world)  # plan-only
        d_.set_chunks(3)
        for plane in ("sparse", "dense"):
            part = rvdist.Partition(d_, bounds, rows.row_ptr, rows.col_idx, rows.edge_w, plane=plane)
            got = part.download()
            assert part.n_halo == model.n_halo
            assert part.info["remote_edges"] == int((model.graph.col_idx >= model.n_local).sum())
            assert np.array_equal(got["recv_counts"], model.recv_counts.astype(np.uint64))
            if plane == "sparse":
                assert np.array_equal(got["col_idx"], model.graph.col_idx)
                assert np.array_equal(got["need_ids"], model.need_ids.astype(np.uint32))
            else:
                assert np.array_equal(got["col_idx"], rvdist.dense_columns(model, rvdist.chunk_starts(hi - lo, 3)))
            interior, _ = rvdist.interior_runs(model)
            assert part.interior_rows == sum(b - a for a, b in interior)
            with pytest.raises(rv.errors.NcclError):
                part.forward(rv.RuvectorLayer(128, 128, 4, 0.1, ctx=ctx), 0x1000, 0x2000)
            part.close()
        d_.close()


def test_world_of_one_is_the_single_gpu_forward(built):
    """ruvgnn_dist_forward with world == 1 (no communicator, no halo) returns the bits of
    ruvgnn_layer_forward, host and device buffers."""
    import torch

    import ruvector_b200 as rv
    from ruvector_b200 import dist as rvdist, synth

    ctx, _ = _torch_ctx()
    n, d, heads = 5000, 128, 4
    g = synth.hnsw_like_graph(n, 16, seed=2)
    x = synth.make_features(n, d, seed=4)
    layer = rv.RuvectorLayer(d, d, heads, 0.1, weights=synth.make_weights(d, d), ctx=ctx)
    want = layer.forward_graph(x, g.to_device(ctx))
    d_ = rvdist.Dist(ctx, 0, 1)
    part = rvdist.Partition.from_graph(d_, g, n)
    assert part.n_halo == 0 and not part.dense and part.interior_rows == n
    out = np.empty((n, d), np.float32)
    part.forward_host(layer, x, out)
    assert np.array_equal(out, want)
    xd, od = torch.from_numpy(x).cuda(), torch.empty(n, d, device="cuda")
    part.forward(layer, xd, od, asynchronous=False)
    assert np.array_equal(od.cpu().numpy(), want)
    assert d_.last_stats()["passes"] == 1
    torch.cuda.set_stream(torch.cuda.default_stream())


@pytest.mark.parametrize("level,m", [(1, 8), (2, 8), (1, 32)])
def test_sparse_level_equals_the_dense_path_bit_for_bit(built, oracle, level, m):
    """Upper HNSW levels: only ~n / M^level rows have neighbours.  The active-row form (Q projection, gather
    and update on those rows alone, LN(W_msg x + b) for the rest -- layer.rs:385-389) must return the bits of
    the dense path, which is the oracle-checked one."""
    import ruvector_b200 as rv
    from ruvector_b200 import _lib, synth

    ctx = rv.Context(0)
    n, d, heads = 30000, 128, 4
    g = synth.hnsw_like_graph(n, m, seed=level * 100 + m, level=level)
    assert 0 < (g.degrees() > 0).sum() < n // 2
    x = synth.make_features(n, d, seed=5)
    layer = rv.RuvectorLayer(d, d, heads, 0.1, weights=synth.make_weights(d, d, seed=level), ctx=ctx)
    dense = layer.forward_graph(x, g.to_device(ctx, flags=_lib.GRAPH_NO_ACTIVE_LIST))
    # host buffers go through the row-block pipeline (dense); device buffers take the active-row form
    import torch

    with torch.cuda.stream(ctx.torch_stream()):
        xd = torch.from_numpy(x).cuda()
        od = torch.full((n, d), float("nan"), device="cuda")
        dg = g.to_device(ctx)
        layer.forward_graph_device(xd.data_ptr(), od.data_ptr(), dg, asynchronous=False)
        sparse = od.cpu().numpy()
    assert np.array_equal(sparse, dense)
    assert row_rel(sparse, oracle.forward_graph(layer, x, g)) <= 1e-4


def test_stack_over_levels_uses_sparse_levels_and_matches_the_oracle(built, oracle):
    """Config-3 shape in small: three layers over levels 2, 1, 0 (search.rs:116-159 order); levels 2 and 1 run
    the active-row form inside ruvgnn_stack_forward."""
    import ruvector_b200 as rv
    from ruvector_b200 import synth

    ctx = rv.Context(0)
    n, d, heads, m = 40000, 128, 4, 8
    levels = [2, 1, 0]
    gs = [synth.hnsw_like_graph(n, m, seed=77, level=lv) for lv in levels]
    x = synth.make_features(n, d, seed=6)
    layers = [rv.RuvectorLayer(d, d, heads, 0.1, weights=synth.make_weights(d, d, seed=40 + l), ctx=ctx) for l in range(3)]
    got = rv.forward_stack(layers, x, [g.to_device(ctx) for g in gs])
    ref = x
    for l, g in zip(layers, gs):
        ref = oracle.forward_graph(l, ref, g)
    assert row_rel(got, ref) <= 1e-4
