"""Embedding store layouts (row f3): byte-exact mapping of the reference's files."""
import mmap
import struct

import numpy as np
import pytest

import ruvector_b200 as rv
from ruvector_b200 import store, synth

F32 = np.float32


def test_mmap_manager_layout(tmp_path):  # mmap.rs:118-203
    p = tmp_path / "emb.bin"
    m = store.open_mmap_embeddings(p, 6, 10)
    assert p.stat().st_size == 10 * 6 * 4 and m.shape == (10, 6)
    m[3] = np.arange(6, dtype=F32)
    m.flush()
    raw = p.read_bytes()
    off = 3 * 6 * 4  # embedding_offset(node_id) = node_id * d_embed * 4
    assert np.frombuffer(raw[off:off + 24], dtype=F32).tolist() == [0, 1, 2, 3, 4, 5]
    again = store.open_mmap_embeddings(p, 6, 10, mode="r")
    assert again[3].tolist() == [0, 1, 2, 3, 4, 5]
    with pytest.raises(rv.InvalidInput):
        store.open_mmap_embeddings(p, 6, 11, mode="r")


def test_feature_storage_layout(tmp_path):  # cold_tier.rs:45-149
    p = tmp_path / "feat.bin"
    v = store.create_feature_storage(p, 5, 7)
    block = (5 * 4 + mmap.PAGESIZE - 1) // mmap.PAGESIZE * mmap.PAGESIZE
    assert p.stat().st_size == 24 + 7 * block and v.shape == (7, 5)
    v[2] = [1.5, 2.5, 3.5, 4.5, 5.5]
    base = v
    while not isinstance(base, np.memmap):
        base = base.base
    base.flush()
    raw = p.read_bytes()
    assert struct.unpack("<QQQ", raw[:24]) == (5, 7, block)
    off = 24 + 2 * block  # HEADER_SIZE + node_id * block_size
    assert np.frombuffer(raw[off:off + 20], dtype=F32).tolist() == [1.5, 2.5, 3.5, 4.5, 5.5]
    r = store.open_feature_storage(p)
    assert r.shape == (7, 5) and r[2].tolist() == [1.5, 2.5, 3.5, 4.5, 5.5]
    bad = tmp_path / "bad.bin"
    bad.write_bytes(raw[:24 + block])
    with pytest.raises(rv.InvalidInput):
        store.open_feature_storage(bad)


@pytest.mark.gpu
def test_forward_reads_and_writes_store_files(tmp_path, oracle):
    n, d = 3000, 128
    ctx = rv.default_context(0)
    layer = rv.RuvectorLayer(d, d, 4, 0.1, weights=synth.make_weights(d, d, seed=9), ctx=ctx)
    x = synth.make_features(n, d, seed=10)
    g = synth.hnsw_like_graph(n, 8, seed=11)
    dense = store.open_mmap_embeddings(tmp_path / "x.bin", d, n + 50)  # max_nodes > n
    dense[:n] = x
    paged = store.create_feature_storage(tmp_path / "x.feat", d, n)
    paged[:] = x
    out_file = store.open_mmap_embeddings(tmp_path / "out.bin", d, n)
    dg = g.to_device(ctx)
    a = store.forward_from_store(layer, dense, dg, out_store=out_file)
    b = store.forward_from_store(layer, paged, dg)
    ref = oracle.forward_graph(layer, x, g)
    np.testing.assert_array_equal(a, b)
    rel = np.linalg.norm(a - ref, axis=1) / np.maximum(np.linalg.norm(ref, axis=1), 1e-6)
    assert rel.max() <= 1e-4
    np.testing.assert_array_equal(store.open_mmap_embeddings(tmp_path / "out.bin", d, n, mode="r"), a)


@pytest.mark.gpu
@pytest.mark.parametrize("block_rows", [0, 1000, 4096])
def test_streamed_store_forward_equals_the_in_memory_forward(tmp_path, block_rows):
    """ruvgnn_layer_forward_store: X streamed from a MmapManager file / a page-strided FeatureStorage file and
    the new embeddings streamed into store files, in double-buffered row blocks -- the bits of
    ruvgnn_layer_forward, with neither X nor the output resident as a whole."""
    n, d, h = 9000, 128, 256
    ctx = rv.default_context(0)
    layer = rv.RuvectorLayer(d, h, 4, 0.1, weights=synth.make_weights(d, h, seed=19), ctx=ctx)
    x = synth.make_features(n, d, seed=20)
    g = synth.hnsw_like_graph(n, 8, seed=21)
    dg = g.to_device(ctx)
    want = layer.forward_graph(x, dg)
    dense = store.open_mmap_embeddings(tmp_path / "x.bin", d, n + 7)
    dense[:n] = x
    dense.flush()
    paged = store.create_feature_storage(tmp_path / "x.feat", d, n)
    paged[:] = x
    base = paged
    while not isinstance(base, np.memmap):
        base = base.base
    base.flush()
    store.open_mmap_embeddings(tmp_path / "o.bin", h, n).flush()
    store.create_feature_storage(tmp_path / "o.feat", h, n)
    # dense in -> dense out
    store.forward_store_files(layer, dg, store.describe_mmap(tmp_path / "x.bin", d, n + 7),
                              store.describe_mmap(tmp_path / "o.bin", h, n), block_rows)
    np.testing.assert_array_equal(store.open_mmap_embeddings(tmp_path / "o.bin", h, n, mode="r"), want)
    # page-strided in -> page-strided out
    store.forward_store_files(layer, dg, store.describe_feature_storage(tmp_path / "x.feat"),
                              store.describe_feature_storage(tmp_path / "o.feat"), block_rows)
    np.testing.assert_array_equal(store.open_feature_storage(tmp_path / "o.feat"), want)
    # errors: wrong width, too few rows
    with pytest.raises(rv.DimensionMismatch):
        store.forward_store_files(layer, dg, store.describe_mmap(tmp_path / "o.bin", h, n),
                                  store.describe_mmap(tmp_path / "o.bin", h, n))
    with pytest.raises(rv.InvalidInput):
        store.forward_store_files(layer, dg, store.describe_mmap(tmp_path / "x.bin", d, n - 1),
                                  store.describe_mmap(tmp_path / "o.bin", h, n))


def test_store_descriptors_without_a_device(tmp_path):
    """The descriptor helpers are host-only byte work (header parsing, size checks)."""
    p = tmp_path / "feat.bin"
    store.create_feature_storage(p, 5, 7)
    st = store.describe_feature_storage(p)
    block = (5 * 4 + mmap.PAGESIZE - 1) // mmap.PAGESIZE * mmap.PAGESIZE
    assert (st.offset_bytes, st.row_stride_bytes, st.dim, st.rows) == (24, block, 5, 7)
    store.open_mmap_embeddings(tmp_path / "m.bin", 6, 10).flush()
    sm = store.describe_mmap(tmp_path / "m.bin", 6, 10)
    assert (sm.offset_bytes, sm.row_stride_bytes, sm.dim, sm.rows) == (0, 24, 6, 10)
    with pytest.raises(rv.InvalidInput):
        store.describe_mmap(tmp_path / "m.bin", 6, 11)
    (tmp_path / "short.bin").write_bytes(b"\0" * 10)
    with pytest.raises(rv.InvalidInput):
        store.describe_feature_storage(tmp_path / "short.bin")
