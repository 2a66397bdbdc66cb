"""Full-size north-star configurations on one B200, checked against the oracle on sampled rows.

Config 3 (BASELINE configs[2]): 10M nodes x 128, 4 heads, three stacked layers over HNSW levels
2, 1, 0 (search.rs:116-159 order).  Graphs and features are generated on the device
(ruvector_b200/synth_cuda.py); for each sampled output row the dependency closure through the three
layers (its neighbours, their neighbours on the level above, ...) is pulled back to the host and
pushed through the oracle layer by layer -- the reference's own per-node arithmetic on exactly the
rows the device used."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _rows_of(rp, col, w, ids):
    """Host copies of the CSR rows `ids` of a device CSR (row_ptr int64, col int32, w f32)."""
    import torch

    ids_t = torch.as_tensor(ids, device=rp.device, dtype=torch.int64)
    beg, end = rp[ids_t].cpu().numpy(), rp[ids_t + 1].cpu().numpy()
    out = {}
    for i, b, e in zip(ids, beg, end):
        out[int(i)] = (col[b:e].cpu().numpy().astype(np.int64), w[b:e].cpu().numpy())
    return out


def _oracle_stack_on_closure(oracle, layers, level_csrs, x_dev, sample):
    """Outputs of the stacked forward at `sample`, computed by the oracle on the dependency closure."""
    import torch

    from ruvector_b200.graph import CsrGraph

    nl = len(layers)
    need_out = [None] * nl  # rows whose layer-l output is needed
    rows = [None] * nl
    need = np.unique(np.asarray(sample, dtype=np.int64))
    for l in range(nl - 1, -1, -1):
        need_out[l] = need
        rows[l] = _rows_of(*level_csrs[l], need)
        nb = [rows[l][int(i)][0] for i in need]
        need = np.unique(np.concatenate([need] + nb)) if nb else need
    cur_ids = need  # rows whose INPUT embedding is needed by layer 0
    cur = x_dev[torch.as_tensor(cur_ids, device=x_dev.device)].cpu().numpy()
    for l in range(nl):
        pos = {int(v): k for k, v in enumerate(cur_ids)}
        rp = np.zeros(cur_ids.size + 1, dtype=np.uint64)
        cols, ws = [], []
        active = set(int(i) for i in need_out[l])
        for k, v in enumerate(cur_ids):
            if int(v) in active:
                c, w = rows[l][int(v)]
                cols.append(np.array([pos[int(j)] for j in c], dtype=np.uint32))
                ws.append(w)
                rp[k + 1] = rp[k] + c.size
            else:
                rp[k + 1] = rp[k]
        g = CsrGraph(rp, np.concatenate(cols) if cols else np.zeros(0, np.uint32),
                     np.concatenate(ws).astype(np.float32) if ws else np.zeros(0, np.float32))
        out = oracle.forward_graph(layers[l], cur, g)
        keep = np.array([pos[int(i)] for i in need_out[l]], dtype=np.int64)
        cur_ids, cur = need_out[l], out[keep]
    return cur_ids, cur


def test_config3_full_size_sampled_against_oracle(built, oracle):
    import torch

    import ruvector_b200 as rv
    from ruvector_b200 import synth, synth_cuda

    free, _ = torch.cuda.mem_get_info()
    if free < 90e9:
        pytest.skip("needs ~60 GB of free HBM")
    n, d, heads, m = 10_000_000, 128, 4, 32
    stream = torch.cuda.Stream()
    torch.cuda.set_stream(stream)
    ctx = rv.Context(0, stream.cuda_stream)
    levels = [2, 1, 0]
    x = synth_cuda.make_features(n, d, seed=0x5EED)
    csrs = [synth_cuda.hnsw_like_graph(n, m, seed=0xBEEF, level=lv) for lv in levels]
    graphs = [rv.DeviceGraph(ctx, *c) for c in csrs]
    layers = [rv.RuvectorLayer(d, d, heads, 0.1, weights=synth.make_weights(d, d, seed=0xC0FFEE + 17 * l), ctx=ctx)
              for l in range(3)]
    out = torch.empty(n, d, device="cuda")
    rv.forward_stack_device(layers, x.data_ptr(), out.data_ptr(), graphs, asynchronous=False)
    assert bool(torch.isfinite(out[::1009]).all())
    rng = np.random.default_rng(7)
    deg0 = (csrs[2][0][1:] - csrs[2][0][:-1])
    present1 = torch.nonzero((csrs[1][0][1:] - csrs[1][0][:-1]) > 0).squeeze(1)[:6].cpu().numpy()  # rows active on level 1
    present2 = torch.nonzero((csrs[0][0][1:] - csrs[0][0][:-1]) > 0).squeeze(1)[:4].cpu().numpy()  # ... and on level 2
    iso = torch.nonzero(deg0 == 0).squeeze(1)[:3].cpu().numpy()
    sample = np.unique(np.concatenate([rng.integers(0, n, size=24), np.arange(3), np.arange(n - 3, n), present1, present2, iso]))
    ids, ref = _oracle_stack_on_closure(oracle, layers, csrs, x, sample)
    got = out[torch.as_tensor(ids, device="cuda")].cpu().numpy()
    rel = np.linalg.norm(got - ref, axis=1) / np.maximum(np.linalg.norm(ref, axis=1), 1e-6)
    assert rel.max() <= 1e-4, float(rel.max())
    torch.cuda.set_stream(torch.cuda.default_stream())
