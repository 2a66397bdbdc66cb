"""Batched differentiable_search / hierarchical_forward on the device (row f2) against the
oracle's restatement of crates/ruvector-gnn/src/search.rs and the reference's own test
properties (search.rs:164-247)."""
import numpy as np
import pytest

import ruvector_b200 as rv
from ruvector_b200 import synth

pytestmark = pytest.mark.gpu
F32 = np.float32


@pytest.fixture(scope="module")
def ctx(built):
    return rv.default_context(0)


def test_differentiable_search_reference_example(ctx):  # search.rs:206-225
    q = np.array([1.0, 0.0, 0.0], dtype=F32)
    cands = np.array([[1.0, 0.0, 0.0], [0.9, 0.1, 0.0], [0.0, 1.0, 0.0]], dtype=F32)
    idx, w = rv.differentiable_search(q, cands, 2, 1.0, ctx=ctx)
    assert idx.tolist() == [0, 1] and w[0] > w[1] and w.shape == (2,)
    idx, w = rv.differentiable_search(q, cands, 10, 1.0, ctx=ctx)  # k clamps to the candidate count
    assert len(idx) == 3 and abs(float(w.sum()) - 1.0) < 1e-6
    idx, w = rv.differentiable_search(q, np.zeros((0, 3), F32), 2, 1.0, ctx=ctx)
    assert len(idx) == 0 and len(w) == 0


@pytest.mark.parametrize("n,dim,k,temp", [(1000, 128, 5, 1.0), (37, 16, 5, 0.5), (4, 8, 5, 1.0), (5000, 384, 8, 2.0)])
def test_differentiable_search_matches_oracle(oracle, ctx, n, dim, k, temp):
    rng = np.random.default_rng(n)
    cands = synth.make_features(n, dim, seed=n)
    cands[n // 2] = 0.0  # zero-norm candidate -> similarity 0 (search.rs:21-25)
    qs = rng.standard_normal((9, dim)).astype(F32)
    qs[3] = 0.0  # zero query: all similarities 0 -> uniform weights, first k indices
    idx, w = rv.differentiable_search(qs, cands, k, temp, ctx=ctx)
    for i in range(qs.shape[0]):
        ri, rw = oracle.differentiable_search(qs[i], cands, k, temp)
        assert idx[i].tolist() == ri.tolist()
        np.testing.assert_allclose(w[i], rw, rtol=2e-5, atol=1e-9)


def test_hierarchical_forward_matches_oracle(oracle, ctx):
    """Three levels (top to bottom: 40, 300, 2000 embeddings), widths 64 -> 128 -> 128 -> 64."""
    dims = [(64, 128), (128, 128), (128, 64)]
    layers = [rv.RuvectorLayer(a, b, 4, 0.1, weights=synth.make_weights(a, b, seed=40 + i), ctx=ctx)
              for i, (a, b) in enumerate(dims)]
    embs = [synth.make_features(n, d[0], seed=50 + i) for i, (n, d) in enumerate(zip((40, 300, 2000), dims))]
    qs = synth.make_features(33, 64, seed=60)
    out = rv.hierarchical_forward(qs, embs, layers, ctx=ctx)
    assert out.shape == (33, 64)
    for i in range(qs.shape[0]):
        ref = oracle.hierarchical_forward(qs[i], embs, layers)
        rel = np.linalg.norm(out[i] - ref) / max(np.linalg.norm(ref), 1e-6)
        assert rel <= 1e-4, (i, rel)
    one = rv.hierarchical_forward(qs[0], embs, layers, ctx=ctx)
    np.testing.assert_array_equal(one, out[0])


def test_hierarchical_forward_reference_properties(ctx):  # search.rs:228-246 and :110-112,121-123
    layer = rv.RuvectorLayer(4, 4, 2, 0.1, ctx=ctx, rng=np.random.default_rng(1))
    q = np.array([1.0, 0.0, 0.0, 0.0], dtype=F32)
    embs = [np.array([[1.0, 0.0, 0.0, 0.0], [0.0, 1.0, 0.0, 0.0]], dtype=F32)]
    out = rv.hierarchical_forward(q, embs, [layer], ctx=ctx)
    assert out.shape == (4,)
    np.testing.assert_array_equal(rv.hierarchical_forward(q, [], [layer], ctx=ctx), q)
    np.testing.assert_array_equal(rv.hierarchical_forward(q, embs, [], ctx=ctx), q)
    # an empty level is skipped
    out2 = rv.hierarchical_forward(q, [np.zeros((0, 4), F32), embs[0]], [layer, layer], ctx=ctx)
    np.testing.assert_array_equal(out2, out)
    with pytest.raises(rv.DimensionMismatch):
        rv.hierarchical_forward(q, [np.ones((3, 5), F32)], [layer], ctx=ctx)
