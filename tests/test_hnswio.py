"""hnsw_rs dump reader (row f1): bit-exact CSR from the reference's on-disk graph format."""
import numpy as np
import pytest

import ruvector_b200 as rv
from ruvector_b200.hnswio import HnswDump
from tests.hnsw_dump_writer import NB_LAYER_MAX, write_graph_dump


def make_points(n, m, seed):
    """HNSW-shaped lists: level l lists only reference points whose own level >= l."""
    rng = np.random.default_rng(seed)
    level = np.minimum(np.floor(-np.log(rng.random(n)) / np.log(m)).astype(int), 3)
    perm = rng.permutation(n)  # file order != origin order
    rank_in = {}
    pts = []
    for lv in range(int(level.max()) + 1):
        members = [int(i) for i in perm if level[i] == lv]
        for j, i in enumerate(members):
            rank_in[i] = (lv, j)
    for lv in range(int(level.max()) + 1):
        for i in [int(i) for i in perm if level[i] == lv]:
            nbrs = []
            for l in range(NB_LAYER_MAX):
                cand = np.nonzero(level >= l)[0]
                cand = cand[cand != i]
                k = 0
                if l <= level[i] and cand.size:
                    k = int(rng.integers(1, min(2 * m if l == 0 else m, cand.size) + 1))
                elif l == level[i] + 1 and cand.size and rng.random() < 0.2:
                    k = 1  # reverse link filed above the point's own level (SURVEY 8d note)
                ids = rng.choice(cand, size=k, replace=False) if k else np.zeros(0, int)
                d = np.sort(rng.random(k).astype(np.float32))
                nbrs.append([(int(a), float(b), *rank_in[int(a)]) for a, b in zip(ids, d)])
            pts.append({"origin_id": i, "layer": lv, "rank": rank_in[i][1], "neighbours": nbrs})
    return pts, level


@pytest.mark.parametrize("full,version", [(True, 4), (False, 4), (True, 3)])
def test_dump_to_csr_is_bit_exact(tmp_path, built, full, version):
    n, m = 300, 6
    pts, level = make_points(n, m, seed=5)
    path = tmp_path / "idx.hnsw.graph"
    write_graph_dump(path, pts, full=full, version=version, max_nb_connection=m, dimension=24)
    d = HnswDump(path)
    assert d.nb_point == n and d.n_ids == n and d.nb_layer == NB_LAYER_MAX
    assert d.max_nb_connection == m and d.dimension == 24 and d.full_mode == full
    by_id = {p["origin_id"]: p for p in pts}
    for l in (0, 1, 2):
        g = d.level_csr(l)
        assert g.n_nodes == n and g.n_edges == d.level_edges(l)
        for i in range(n):
            s, e = int(g.row_ptr[i]), int(g.row_ptr[i + 1])
            want = by_id[i]["neighbours"][l]
            assert g.col_idx[s:e].tolist() == [w[0] for w in want]
            assert g.edge_w[s:e].tobytes() == np.array([w[1] for w in want], dtype=np.float32).tobytes()
    # a larger id space leaves holes as empty rows (deleted ids)
    g = d.level_csr(0, n_nodes=n + 10)
    assert g.n_nodes == n + 10 and (g.degrees()[n:] == 0).all()
    with pytest.raises(rv.InvalidInput):
        d.level_csr(0, n_nodes=n - 1)
    with pytest.raises(rv.InvalidInput):
        d.level_csr(NB_LAYER_MAX)


def test_corrupt_dumps_are_rejected(tmp_path, built):
    pts, _ = make_points(50, 4, seed=6)
    path = tmp_path / "ok.hnsw.graph"
    write_graph_dump(path, pts)
    raw = path.read_bytes()
    bad = tmp_path / "bad.hnsw.graph"
    bad.write_bytes(b"\x00\x00\x00\x00" + raw[4:])
    with pytest.raises(rv.InvalidInput) as e:
        HnswDump(bad)
    assert "bad magic" in str(e.value)
    bad.write_bytes(raw[: len(raw) // 2])
    with pytest.raises(rv.InvalidInput):
        HnswDump(bad)
    with pytest.raises(rv.InvalidInput):
        HnswDump(tmp_path / "missing.hnsw.graph")
    # an absurd neighbour count (first list of the first point) must be rejected, not allocated
    import struct
    hdr = 4 + 1 + 1 + 8 + 1 + 8 * 3 + 8 + len("DistCosine") + 8 + len("f32")  # Description (v4)
    pos = hdr + 1 + 4 + 8 + 4 + 8 + 5  # nb_layer, MAGICLAYER, nb_point, MAGICPOINT, origin_id, p_id
    huge = bytearray(raw)
    huge[pos:pos + 8] = struct.pack("=Q", 1 << 61)
    bad.write_bytes(bytes(huge))
    with pytest.raises(rv.InvalidInput) as e:
        HnswDump(bad)
    assert "truncated neighbour list" in str(e.value)
    empty = tmp_path / "empty.hnsw.graph"
    write_graph_dump(empty, [])
    d = HnswDump(empty)
    assert d.nb_point == 0 and d.level_csr(0).n_nodes == 0


def test_mutated_files_never_crash_the_native_parsers(tmp_path, built):
    """600 mutations (truncate / flip / extreme 8-byte fields / duplicate / delete slices) of a valid
    dump and of a layer JSON: each must parse or be rejected with a GnnError status.  Runs in a child
    process (scripts/fuzz_parsers.py) so that a crash is an exit code, not a dead test session."""
    import subprocess
    import sys
    from pathlib import Path

    root = Path(__file__).resolve().parents[1]
    res = subprocess.run([sys.executable, str(root / "scripts" / "fuzz_parsers.py"), str(tmp_path), "300"],
                         capture_output=True, text=True, timeout=600)
    assert res.returncode == 0 and "FUZZ OK" in res.stdout, res.stdout[-1500:] + res.stderr[-1500:]


def test_absurd_origin_id_is_rejected_at_open(tmp_path, built):
    import struct

    pts, _ = make_points(30, 4, seed=7)
    path = tmp_path / "ok.hnsw.graph"
    write_graph_dump(path, pts)
    raw = bytearray(path.read_bytes())
    hdr = 4 + 1 + 1 + 8 + 1 + 8 * 3 + 8 + len("DistCosine") + 8 + len("f32")
    pos = hdr + 1 + 4 + 8 + 4  # nb_layer, MAGICLAYER, nb_point, MAGICPOINT -> origin_id of the first point
    for value in (2**64 - 1, 2**32 - 1, 2**40):
        raw[pos:pos + 8] = struct.pack("=Q", value)
        bad = tmp_path / "bad.hnsw.graph"
        bad.write_bytes(bytes(raw))
        with pytest.raises(rv.InvalidInput) as e:
            HnswDump(bad)
        assert "origin_id exceeds" in str(e.value)
