#!/usr/bin/env python
"""Generates the known-answer fixtures under tests/golden/ with the CPU oracle.

The reference ships no golden vectors for this path, and it cannot be
built in this image, so these vectors are ORACLE-generated: they pin the oracle against drift
and the CUDA path against the oracle.  Each case file carries the layer in the reference's own
serde-JSON shape, so a maintainer with a Rust toolchain can pin the oracle itself:

    let layer: RuvectorLayer = serde_json::from_str(&case["layer_json"])?;
    let out = layer.forward(&case.node, &case.neighbors, &case.edge_weights);
    // compare with case["expected"] (f32, printed as shortest round-trip decimals)

Run from the repo root:  python tests/golden/make_golden.py
"""
import json
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT))
import ruvector_b200 as rv  # noqa: E402
from ruvector_b200 import synth  # noqa: E402
from tests.oracle_binding import Oracle  # noqa: E402

F32 = np.float32


def f32_list(a):
    return [float(s) for s in np.asarray(a, dtype=F32).ravel().astype("U32")]


def main():
    orc = Oracle()
    out_dir = Path(__file__).resolve().parent
    cases = []
    # 1. the reference's own unit-test call (layer.rs:512-520) with seeded weights
    l1 = rv.RuvectorLayer(4, 8, 2, 0.1, weights=synth.make_weights(4, 8, seed=1001))
    cases.append(("unit_test_4_8_2", l1, [1.0, 2.0, 3.0, 4.0],
                  [[0.5, 1.0, 1.5, 2.0], [2.0, 3.0, 4.0, 5.0]], [0.3, 0.7]))
    # 2. no neighbours (layer.rs:523-533): LN(W_msg x + b)
    cases.append(("no_neighbors_4_8_2", l1, [1.0, 2.0, 3.0, 4.0], [], []))
    # 3. reference initialisation (zero biases, gamma = 1), 5 neighbours, 16 -> 32, 4 heads
    l3 = rv.RuvectorLayer(16, 32, 4, 0.25, weights=synth.make_weights(16, 32, seed=1003, trivial=True))
    x3 = synth.make_features(6, 16, seed=1004)
    cases.append(("trivial_init_16_32_4", l3, x3[0], x3[1:], [0.9, 0.1, 0.4, 0.4, 0.2]))
    # 4. negative weights (DotProduct-style distances): uniform rule (layer.rs:428-432)
    cases.append(("nonpositive_weights_16_32_4", l3, x3[0], x3[1:4], [-0.5, -0.25, 0.0]))
    # 5. wider layer, 12 neighbours with ascending distances as edge weights (HNSW list order)
    l5 = rv.RuvectorLayer(48, 64, 4, 0.1, weights=synth.make_weights(48, 64, seed=1005))
    x5 = synth.make_features(13, 48, seed=1006)
    cases.append(("hnsw_list_48_64_4", l5, x5[0], x5[1:], np.sort(np.linspace(0.05, 0.6, 12))))
    for name, layer, node, nbrs, w in cases:
        nb = np.asarray(nbrs, dtype=F32).reshape(-1, layer.input_dim)
        expected = orc.forward_node(layer, np.asarray(node, F32), nb, np.asarray(w, F32))
        doc = {
            "name": name,
            "reference": "crates/ruvector-gnn/src/layer.rs:379-418 (RuvectorLayer::forward)",
            "input_dim": layer.input_dim, "hidden_dim": layer.hidden_dim, "heads": layer.heads,
            "dropout": float(np.float32(layer.dropout)),
            "layer_json": layer.to_json(),
            "node": f32_list(node), "neighbors": [f32_list(r) for r in nb], "edge_weights": f32_list(w),
            "expected": f32_list(expected),
        }
        (out_dir / f"{name}.json").write_text(json.dumps(doc))
        print(name, expected[:3])


if __name__ == "__main__":
    main()
