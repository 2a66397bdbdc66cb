"""NumPy float32 twin of the CPU oracle (TEST INFRASTRUCTURE ONLY).

An implementation of ruvector-gnn's ``RuvectorLayer::forward`` that shares no
code with ``ruvgnn_oracle.c``: whole-graph, vectorised, float32, written from
the maths of ``crates/ruvector-gnn/src/layer.rs`` rather than its loop
structure.  Its only job is to catch a mistake in the C restatement: the two
must agree to float32 rounding (tests/test_oracle.py, <= 2e-6 relative) on
random layers and graphs.  Summation order differs from the reference (NumPy
pairwise / BLAS), so this twin is NOT the parity oracle itself.

Only tests/ may import this module.
"""
from __future__ import annotations

import numpy as np

F = np.float32


def linear(w, b, x, F=F):
    """layer.rs:37-41 -- y = W x + b for a batch of rows x[N, in]; w is [out, in]."""
    return (x.astype(F) @ w.astype(F).T + b.astype(F)).astype(F)


def layer_norm(gamma, beta, eps, x, F=F):
    """layer.rs:68-81 -- biased variance, (x-mean)/sqrt(var+eps)*gamma+beta."""
    x = x.astype(F)
    mean = x.mean(axis=-1, keepdims=True, dtype=F)
    var = ((x - mean) ** 2).mean(axis=-1, keepdims=True, dtype=F)
    return ((x - mean) / np.sqrt(var + F(eps)) * gamma + beta).astype(F)


def sigmoid(x, F=F):
    """layer.rs:278-285 -- the two branches are the same function."""
    x = x.astype(F)
    pos = F(1) / (F(1) + np.exp(-np.abs(x)))
    return np.where(x > 0, pos, F(1) - pos).astype(F)


def forward_graph(weights: dict, input_dim, hidden_dim, heads, dropout, x, row_ptr, col_idx, edge_w, dtype=np.float32):
    """out[i] = RuvectorLayer.forward(x[i], x[N(i)], w[i]) for every node (Jacobi).

    ``weights`` maps the 12 linear names to (W[out,in], b[out]) plus
    ``gamma``, ``beta``, ``eps`` -- the serde field names of layer.rs:315-333.
    ``dtype=np.float64`` evaluates the same formulas in double precision: the yardstick for how
    much of the 1e-4 tolerance the float32 arithmetic of the reference itself uses up.
    """
    F = np.dtype(dtype).type

    def linear(w, b, x_):
        return globals()["linear"](w, b, x_, F)

    def sigmoid(x_):
        return globals()["sigmoid"](x_, F)

    def layer_norm(g_, b_, e_, x_):
        return globals()["layer_norm"](np.asarray(g_, dtype=F), np.asarray(b_, dtype=F), e_, x_, F)

    x = np.asarray(x, dtype=F)
    n = x.shape[0]
    H, h = hidden_dim, heads
    dh = H // h
    row_ptr = np.asarray(row_ptr, dtype=np.int64)
    col_idx = np.asarray(col_idx, dtype=np.int64)
    edge_w = np.asarray(edge_w, dtype=F)
    deg = np.diff(row_ptr)
    src = np.repeat(np.arange(n), deg)  # edge -> owning node

    m = linear(*weights["w_msg"], x)  # :392 (and :393-396 for neighbours)
    q = linear(*weights["q"], m)  # :135
    k = linear(*weights["k"], m)  # :136
    v = linear(*weights["v"], m)  # :137

    # per-edge, per-head scaled scores :187-193
    qe = q[src].reshape(-1, h, dh)
    ke = k[col_idx].reshape(-1, h, dh)
    s = (qe * ke).sum(axis=2, dtype=F) / np.sqrt(F(dh))
    # segment softmax :196-199
    smax = np.full((n, h), -np.inf, dtype=F)
    np.maximum.at(smax, src, s)
    e = np.exp(s - smax[src]).astype(F)
    ssum = np.zeros((n, h), dtype=F)
    np.add.at(ssum, src, e)
    ssum = np.maximum(ssum, F(1e-10))
    a = (e / ssum[src]).astype(F)
    # weighted values :202-207
    ve = v[col_idx].reshape(-1, h, dh)
    concat = np.zeros((n, h, dh), dtype=F)
    np.add.at(concat, src, a[:, :, None] * ve)
    attn = linear(*weights["out"], concat.reshape(n, H))  # :159

    # edge-weight aggregation :421-445
    wsum = np.zeros(n, dtype=F)
    np.add.at(wsum, src, edge_w)
    safe = np.where(wsum > 0, wsum, F(1))
    nw = np.where(wsum[src] > 0, edge_w / safe[src], F(1) / np.maximum(deg[src], 1).astype(F)).astype(F)
    agg = np.zeros((n, H), dtype=F)
    np.add.at(agg, src, nw[:, None] * m[col_idx])

    aggregated = linear(*weights["w_agg"], attn + agg)  # :407-408
    # GRU :255-275, input = aggregated, hidden = m
    z = sigmoid(linear(*weights["w_z"], aggregated) + linear(*weights["u_z"], m))
    r = sigmoid(linear(*weights["w_r"], aggregated) + linear(*weights["u_r"], m))
    ht = np.tanh(linear(*weights["w_h"], aggregated) + linear(*weights["u_h"], r * m)).astype(F)
    upd = (F(1) - z) * m + z * ht
    dropped = upd * (F(1) - F(dropout))  # :448-451
    out = layer_norm(weights["gamma"], weights["beta"], weights["eps"], dropped)  # :417
    iso = deg == 0  # :385-389
    if iso.any():
        out[iso] = layer_norm(weights["gamma"], weights["beta"], weights["eps"], m[iso])
    return out.astype(F)
