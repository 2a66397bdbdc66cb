"""CPU tests of the host logic and the C-ABI surface (no compute calls)."""
import ctypes as C
import re
from pathlib import Path

import numpy as np
import pytest

import ruvector_b200 as rv
from ruvector_b200 import _lib, synth

ROOT = Path(__file__).resolve().parents[1]
F32 = np.float32


def test_library_exports_every_declared_symbol(built):
    header = (ROOT / "include" / "ruvgnn.h").read_text()
    declared = set(re.findall(r"\b(ruvgnn_[a-z_0-9]+)\s*\(", header))
    lib = C.CDLL(str(_lib.lib_path()))
    missing = [s for s in sorted(declared) if not hasattr(lib, s)]
    assert not missing, missing
    assert declared == set(_lib.SYMBOLS), declared ^ set(_lib.SYMBOLS)
    assert _lib.load().ruvgnn_abi_version() == 2


def test_constructor_rejects_bad_dropout_with_reference_message(built):  # layer.rs:353-358,536-539
    with pytest.raises(rv.LayerConfig) as e:
        rv.RuvectorLayer(4, 8, 2, 1.5)
    assert "Dropout must be between 0.0 and 1.0, got 1.5" in str(e.value)
    assert str(e.value).startswith("Layer configuration error: ")  # error.rs:25
    with pytest.raises(rv.LayerConfig):
        rv.RuvectorLayer(4, 8, 2, -0.1)
    with pytest.raises(rv.LayerConfig):
        rv.RuvectorLayer(4, 8, 2, float("nan"))
    rv.RuvectorLayer(4, 8, 2, 0.0)
    rv.RuvectorLayer(4, 8, 2, 1.0)


def test_constructor_rejects_indivisible_heads(built):  # layer.rs:101-106,541-545,494-499
    with pytest.raises(rv.LayerConfig) as e:
        rv.RuvectorLayer(4, 7, 3, 0.1)
    assert "Embedding dimension (7) must be divisible by number of heads (3)" in str(e.value)
    assert "divisible" in str(e.value)
    with pytest.raises(rv.LayerConfig):
        rv.RuvectorLayer(10, 10, 3, 0.1)


def test_default_init_matches_reference_initialisation(built):  # layer.rs:21-34,59-64
    layer = rv.RuvectorLayer(128, 256, 4, 0.1, rng=np.random.default_rng(0))
    W, b = layer.weights["w_msg"]
    assert W.shape == (256, 128) and not b.any()
    assert abs(float(W.std()) - np.sqrt(2.0 / (128 + 256))) < 2e-3
    assert (layer.weights["gamma"] == 1).all() and not layer.weights["beta"].any()
    assert layer.weights["eps"] == pytest.approx(1e-5)
    assert layer.weights["u_h"][0].shape == (256, 256)


def test_json_has_serde_shape_and_roundtrips_bit_exact(built):
    layer = rv.RuvectorLayer(4, 8, 2, 0.1, weights=synth.make_weights(4, 8, seed=5))
    text = layer.to_json()
    import json

    doc = json.loads(text)
    assert list(doc) == ["w_msg", "w_agg", "w_update", "attention", "norm", "dropout"]  # layer.rs:315-333
    assert list(doc["w_update"]) == ["w_z", "u_z", "w_r", "u_r", "w_h", "u_h"]  # :215-227
    assert list(doc["attention"]) == ["num_heads", "head_dim", "q_linear", "k_linear", "v_linear", "out_linear"]
    assert doc["attention"]["num_heads"] == 2 and doc["attention"]["head_dim"] == 4
    assert doc["w_msg"]["weights"]["dim"] == [8, 4] and doc["w_msg"]["weights"]["v"] == 1
    assert doc["norm"]["eps"] == pytest.approx(1e-5) and doc["dropout"] == pytest.approx(0.1)
    again = rv.RuvectorLayer.from_json(text)
    for n in synth.LINEARS:
        np.testing.assert_array_equal(again.weights[n][0], layer.weights[n][0])
        np.testing.assert_array_equal(again.weights[n][1], layer.weights[n][1])
    np.testing.assert_array_equal(again.weights["gamma"], layer.weights["gamma"])
    assert again.dropout == layer.dropout and again.heads == 2


def test_c_json_parser_and_emitter_agree_with_python(built):
    """ruvgnn_layer_json_normalize = the C parse + emit used by from_json/to_json, device-free."""
    layer = rv.RuvectorLayer(12, 24, 3, 0.25, weights=synth.make_weights(12, 24, seed=8))
    text = layer.to_json().encode()
    lib = _lib.load()
    need = C.c_size_t()
    _lib.check(lib.ruvgnn_layer_json_normalize(text, len(text), None, 0, C.byref(need)))
    buf = C.create_string_buffer(need.value + 1)
    _lib.check(lib.ruvgnn_layer_json_normalize(text, len(text), buf, need.value + 1, C.byref(need)))
    again = rv.RuvectorLayer.from_json(buf.value.decode())
    for n in synth.LINEARS:
        np.testing.assert_array_equal(again.weights[n][0], layer.weights[n][0])
        np.testing.assert_array_equal(again.weights[n][1], layer.weights[n][1])
    np.testing.assert_array_equal(again.weights["beta"], layer.weights["beta"])
    assert again.weights["eps"] == layer.weights["eps"] and again.dropout == layer.dropout
    # whitespace / unknown keys are tolerated, a truncated file is not
    spaced = text.replace(b",", b" ,\n ").replace(b"{\"w_msg\"", b"{ \"extra\": [1, {\"a\": null}], \"w_msg\"")
    _lib.check(lib.ruvgnn_layer_json_normalize(spaced, len(spaced), None, 0, C.byref(need)))
    st = lib.ruvgnn_layer_json_normalize(text[:-40], len(text) - 40, None, 0, C.byref(need))
    assert st == 7 and "Deserialization error" in _lib.last_error()
    bad = text.replace(b'"dim":[24,12]', b'"dim":[24,13]', 1)
    st = lib.ruvgnn_layer_json_normalize(bad, len(bad), None, 0, C.byref(need))
    assert st == 7 and "data and dimension must match in size" in _lib.last_error()


def test_packer_is_bit_exact_and_keeps_order(built):
    rng = np.random.default_rng(4)
    n = 500
    lists, ids, dists = [], [], []
    for i in range(n):
        k = int(rng.integers(0, 33)) if i % 17 else 0  # ragged, some empty
        a = rng.choice(n, size=k, replace=False).astype(np.uint64)
        d = np.sort(rng.random(k).astype(F32))
        r = np.zeros(k, dtype=rv.NEIGHBOUR_DTYPE)
        r["d_id"], r["distance"] = a, d
        r["p_layer"], r["p_rank"] = rng.integers(0, 16, k), rng.integers(0, 1 << 20, k)  # must be ignored
        lists.append(r)
        ids.append(a)
        dists.append(d)
    g = rv.pack_neighbour_lists(lists)
    assert g.row_ptr.dtype == np.uint64 and g.col_idx.dtype == np.uint32 and g.edge_w.dtype == F32
    np.testing.assert_array_equal(g.row_ptr, np.concatenate([[0], np.cumsum([len(a) for a in ids])]).astype(np.uint64))
    np.testing.assert_array_equal(g.col_idx, np.concatenate(ids).astype(np.uint32))
    assert g.edge_w.tobytes() == np.concatenate(dists).astype(F32).tobytes()
    g2 = rv.pack_id_distance_lists([a.tolist() for a in ids], [d.tolist() for d in dists])
    assert g2.col_idx.tobytes() == g.col_idx.tobytes() and g2.edge_w.tobytes() == g.edge_w.tobytes()


def test_packer_rejects_out_of_range_ids_and_handles_empty(built):
    r = np.zeros(1, dtype=rv.NEIGHBOUR_DTYPE)
    r["d_id"] = 3
    with pytest.raises(rv.InvalidInput) as e:
        rv.pack_neighbour_lists([r, r[:0], r[:0]])
    assert "out of range" in str(e.value)
    g = rv.pack_neighbour_lists([r[:0], r[:0]])
    assert g.n_nodes == 2 and g.n_edges == 0 and g.row_ptr.tolist() == [0, 0, 0]
    g = rv.pack_neighbour_lists([])
    assert g.n_nodes == 0 and g.row_ptr.tolist() == [0]


def test_batch_request_packing(built):  # gnn_cache.rs:295-313
    ops = [
        (np.ones(4), [np.zeros(4), np.full(4, 2.0)], [0.3, 0.7]),
        (np.full(4, 5.0), [], []),
        (np.full(4, 6.0), [np.full(4, 7.0)], [1.0]),
    ]
    x, g, q = rv.batch_to_graph(ops)
    assert x.shape == (6, 4) and q.tolist() == [0, 3, 4]
    assert g.row_ptr.tolist() == [0, 2, 2, 2, 2, 3, 3] and g.col_idx.tolist() == [1, 2, 5]
    assert g.edge_w.tolist() == pytest.approx([0.3, 0.7, 1.0])
    with pytest.raises(rv.InvalidInput):
        rv.batch_to_graph([(np.ones(4), [np.zeros(4)], [])])


def test_synthetic_graph_invariants(built):
    g = synth.hnsw_like_graph(3000, 8, seed=2, locality=0.5, window=64, iso_frac=0.02)
    deg = g.degrees()
    assert deg.max() <= 16 and (deg[deg > 0] >= 8).all() and (deg == 0).sum() > 0
    for i in (0, 17, 2999):
        s, e = int(g.row_ptr[i]), int(g.row_ptr[i + 1])
        row = g.col_idx[s:e]
        assert len(set(row.tolist())) == len(row) and i not in row
        assert (np.diff(g.edge_w[s:e]) >= 0).all()
    g1 = synth.hnsw_like_graph(3000, 8, seed=2, level=1)
    assert 0 < (g1.degrees() > 0).sum() < 3000 * 0.25
    x = synth.make_features(400, 16)
    k = synth.knn_graph(x, 4, iso_frac=0.0)
    s, e = int(k.row_ptr[5]), int(k.row_ptr[6])
    sims = x[k.col_idx[s:e]] @ x[5]
    assert (np.diff(sims) <= 1e-6).all() and np.allclose(k.edge_w[s:e], 1 - sims, atol=1e-5)


def test_compute_entry_fails_loudly_without_a_gpu(built):
    import torch

    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    layer = rv.RuvectorLayer(4, 8, 2, 0.1)
    with pytest.raises(rv.CudaError) as e:
        layer.forward(np.ones(4, F32), [np.ones(4, F32)], [1.0])
    assert "no CPU fallback" in str(e.value)


def test_product_package_never_touches_the_oracle():
    """The product path may not import, link or execute anything under oracle/."""
    pat = re.compile(r"import\s+oracle|from\s+oracle|liboracle|oracle/|orc_[a-z_]+\s*\(|tests\.oracle_binding")
    for p in (ROOT / "ruvector_b200").rglob("*"):
        if p.suffix in {".py", ".cu", ".cuh", ".cc", ".h"}:
            assert not pat.search(p.read_text()), p


def test_locality_order_recovers_hidden_locality_and_keeps_every_row(built):
    """Insertion-order ids hide the graph's locality: relabel a window graph at random, then let
    locality_order() find an order in which range partitions keep their neighbours."""
    from ruvector_b200.graph import inverse_permutation, locality_order, permute_graph, remote_edge_fraction

    n = 20000
    g = synth.hnsw_like_graph(n, 8, seed=21, locality=1.0, window=40, iso_frac=0.01)
    rng = np.random.default_rng(3)
    shuffle = rng.permutation(n)
    hidden = permute_graph(g, shuffle)
    assert remote_edge_fraction(g, 8) < 0.02 and remote_edge_fraction(hidden, 8) > 0.8
    perm = locality_order(hidden)
    assert np.array_equal(np.sort(perm), np.arange(n))
    found = permute_graph(hidden, perm)
    assert remote_edge_fraction(found, 8) < 0.1
    # rows are relabelled, never rewritten: row p of the permuted graph is row perm[p] with renamed columns
    inv = inverse_permutation(perm)
    for p in (0, 1, n // 2, n - 1):
        s, t = int(found.row_ptr[p]), int(found.row_ptr[p + 1])
        s0, t0 = int(hidden.row_ptr[perm[p]]), int(hidden.row_ptr[perm[p] + 1])
        assert np.array_equal(found.col_idx[s:t], inv[hidden.col_idx[s0:t0]])
        assert np.array_equal(found.edge_w[s:t], hidden.edge_w[s0:t0])
    back = permute_graph(found, inv)
    assert np.array_equal(back.row_ptr, hidden.row_ptr) and np.array_equal(back.col_idx, hidden.col_idx)
    with pytest.raises(rv.GnnError):
        permute_graph(g, np.zeros(n, dtype=np.int64))


def test_header_is_plain_c_and_cxx(tmp_path):
    """include/ruvgnn.h is the drop-in boundary: it must compile as C99 (what cgo / bindgen / ctypes
    generators read) and as C++, with no dependency beyond <stdint.h>/<stddef.h>."""
    import subprocess
    from pathlib import Path

    root = Path(__file__).resolve().parents[1]
    src = tmp_path / "use.c"
    src.write_text('#include "ruvgnn.h"\nint main(void) { ruvgnn_weights w; (void)w; return RUVGNN_OK; }\n')
    subprocess.run(["gcc", "-std=c99", "-pedantic", "-Wall", "-Werror", "-fsyntax-only", "-I", str(root / "include"), str(src)],
                   check=True)
    subprocess.run(["g++", "-std=c++17", "-Wall", "-Werror", "-fsyntax-only", "-x", "c++", "-I", str(root / "include"), str(src)],
                   check=True)
    hpp = tmp_path / "use.cpp"
    hpp.write_text('#include "ruvgnn.hpp"\nint main() { return 0; }\n')
    subprocess.run(["g++", "-std=c++17", "-Wall", "-Wextra", "-Werror", "-fsyntax-only", "-I", str(root / "include"), str(hpp)],
                   check=True)


def _reorder_bfs_reference(adjacency, num_nodes):
    """Python restatement of DataLoader::reorder_bfs (crates/ruvector-gnn/src/cold_tier.rs:307-342)."""
    from collections import deque

    adj = [[] for _ in range(num_nodes)]
    for u, v in adjacency:
        if u < num_nodes and v < num_nodes:
            adj[u].append(v)
            adj[v].append(u)
    visited, order, queue = [False] * num_nodes, [], deque()
    for start in range(num_nodes):
        if visited[start]:
            continue
        visited[start] = True
        queue.append(start)
        while queue:
            node = queue.popleft()
            order.append(node)
            for nb in adj[node]:
                if not visited[nb]:
                    visited[nb] = True
                    queue.append(nb)
    return order


@pytest.mark.parametrize("seed,loc", [(1, 0.0), (2, 0.7), (3, 1.0)])
def test_bfs_order_is_the_references_reorder_bfs(built, seed, loc):
    """ruvgnn_bfs_order against the restated reference on the CSR's edge list (stored order), incl.
    isolated nodes, several components and a self loop; and it is the scalable default of
    locality_order (the SciPy RCM stays selectable)."""
    from ruvector_b200.graph import CsrGraph, bfs_order, locality_order, permute_graph, remote_edge_fraction

    n = 3000
    g = synth.hnsw_like_graph(n, 6, seed=seed, locality=loc, window=25, iso_frac=0.2)
    col = g.col_idx.copy()
    if col.size:
        col[int(g.row_ptr[5])] = 5 if g.row_ptr[6] > g.row_ptr[5] else col[int(g.row_ptr[5])]  # a self loop
    g = CsrGraph(g.row_ptr, col, g.edge_w)
    rows = np.repeat(np.arange(n), g.degrees())
    want = _reorder_bfs_reference(list(zip(rows.tolist(), g.col_idx.tolist())), n)
    got = bfs_order(g)
    assert got.tolist() == want
    assert np.array_equal(locality_order(g), got)
    assert np.array_equal(np.sort(locality_order(g, method="rcm")), np.arange(n))
    if loc == 1.0:  # a shuffled window graph: BFS pulls the neighbours back together
        rng = np.random.default_rng(seed)
        hidden = permute_graph(synth.hnsw_like_graph(20000, 8, seed=9, locality=1.0, window=40, iso_frac=0.0), rng.permutation(20000))
        assert remote_edge_fraction(hidden, 8) > 0.8
        assert remote_edge_fraction(permute_graph(hidden, bfs_order(hidden)), 8) < 0.25
    assert bfs_order(CsrGraph(np.zeros(1, np.uint64), np.zeros(0, np.uint32), np.zeros(0, np.float32))).size == 0


def test_flatten_levels_mirrors_flatten_point(built):
    """Union-of-levels CSR = hnsw_rs flatten_point (patches/hnsw_rs/src/flatten.rs:74-89): per node the
    concatenation of all level lists sorted by distance (stable), duplicates kept; NaN is an error."""
    from ruvector_b200.graph import CsrGraph, flatten_levels

    n = 500
    levels = [synth.hnsw_like_graph(n, 8, seed=4, level=l, iso_frac=0.05) for l in range(3)]
    flat = flatten_levels(levels)
    assert flat.n_edges == sum(l.n_edges for l in levels)
    for i in range(n):
        recs = []
        for l in levels:
            s, e = int(l.row_ptr[i]), int(l.row_ptr[i + 1])
            recs += list(zip(l.edge_w[s:e].tolist(), l.col_idx[s:e].tolist()))
        recs.sort(key=lambda r: r[0])  # Python's sort is stable, like the native one
        s, e = int(flat.row_ptr[i]), int(flat.row_ptr[i + 1])
        assert flat.edge_w[s:e].tolist() == [r[0] for r in recs]
        assert flat.col_idx[s:e].tolist() == [r[1] for r in recs]
    assert np.array_equal(flatten_levels(levels[:1]).col_idx, levels[0].col_idx)  # level 0 alone is already sorted
    bad = CsrGraph(np.array([0, 1], np.uint64), np.array([0], np.uint32), np.array([np.nan], np.float32))
    with pytest.raises(rv.InvalidInput):
        flatten_levels([bad])
