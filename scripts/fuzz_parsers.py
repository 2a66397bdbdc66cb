#!/usr/bin/env python
"""Mutation fuzz of the two native parsers that read files a user hands over: the hnsw_rs dump
reader (hnswio.cc) and the serde-JSON layer parser (json.cc).  Every mutated input must either
parse or come back as a GnnError -- never crash, hang or allocate absurdly.  Run in a child
process by tests/test_hnswio.py so that a crash shows up as a non-zero exit, not a dead pytest.
  python scripts/fuzz_parsers.py <workdir> [cases]"""
import resource
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))

import ruvector_b200 as rv  # noqa: E402
from ruvector_b200.hnswio import HnswDump  # noqa: E402
from tests.hnsw_dump_writer import write_graph_dump  # noqa: E402
from tests.test_hnswio import make_points  # noqa: E402


def mutate(raw: bytes, rng) -> bytes:
    b = bytearray(raw)
    kind = int(rng.integers(0, 5))
    if kind == 0:  # truncate
        return bytes(b[: int(rng.integers(0, len(b)))])
    if kind == 1:  # flip a few bytes
        for _ in range(int(rng.integers(1, 6))):
            b[int(rng.integers(0, len(b)))] = int(rng.integers(0, 256))
        return bytes(b)
    if kind == 2:  # overwrite an aligned 8-byte field with an extreme value
        pos = int(rng.integers(0, max(len(b) - 8, 1)))
        extremes = [0, 1, 2**31, 2**32 - 1, 2**61, 2**64 - 1]
        b[pos:pos + 8] = extremes[int(rng.integers(0, len(extremes)))].to_bytes(8, "little")
        return bytes(b)
    if kind == 3:  # duplicate a slice
        a, c = sorted(int(v) for v in rng.integers(0, len(b), size=2))
        return bytes(b[:c] + b[a:c] + b[c:])
    a, c = sorted(int(v) for v in rng.integers(0, len(b), size=2))  # delete a slice
    return bytes(b[:a] + b[c:])


def main(workdir, cases):
    resource.setrlimit(resource.RLIMIT_AS, (8 << 30, 8 << 30))  # an absurd allocation fails instead of thrashing
    work = Path(workdir)
    rng = np.random.default_rng(1234)
    pts, _ = make_points(40, 4, seed=9)
    good = work / "good.hnsw.graph"
    write_graph_dump(good, pts)
    raw = good.read_bytes()
    ok = rejected = 0
    for i in range(cases):
        p = work / "m.hnsw.graph"
        p.write_bytes(mutate(raw, rng))
        try:
            d = HnswDump(p)
            for level in range(min(d.nb_layer, 3)):
                g = d.level_csr(level)
                assert g.row_ptr[-1] == g.col_idx.size == g.edge_w.size
                assert g.col_idx.size == 0 or int(g.col_idx.max()) < g.n_nodes
            d.close()
            ok += 1
        except rv.GnnError:
            rejected += 1
    print(f"hnsw dump: {ok} parsed, {rejected} rejected")

    import ctypes as C

    from ruvector_b200 import _lib

    lib = _lib.load()
    layer = rv.RuvectorLayer(4, 8, 2, 0.1, rng=np.random.default_rng(2))
    text = layer.to_json().encode()
    ok = rejected = 0
    need = C.c_size_t()
    for i in range(cases):
        m = mutate(text, rng)
        st = lib.ruvgnn_layer_json_normalize(m, len(m), None, 0, C.byref(need))  # the C parser + emitter
        if st == 0:
            buf = C.create_string_buffer(need.value + 1)
            assert lib.ruvgnn_layer_json_normalize(m, len(m), buf, need.value + 1, C.byref(need)) == 0
            rv.RuvectorLayer.from_json(buf.value.decode())  # what the C side accepts, the Python side reads
            ok += 1
        else:
            assert st in (1, 2, 3, 7), st
            rejected += 1
    print(f"layer json: {ok} parsed, {rejected} rejected")
    print("FUZZ OK")


if __name__ == "__main__":
    main(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 400)
