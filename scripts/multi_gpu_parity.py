#!/usr/bin/env python
"""Partitioned forward (ruvgnn_dist_forward, NCCL inside the library) vs the single-graph CPU
oracle and vs the single-GPU run.  Launch with torchrun (one rank per GPU):
  python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 \
      --master-port 29533 scripts/multi_gpu_parity.py
Every rank builds the same global graph and keeps its row range.  torch.distributed (gloo) only
ships the NCCL id and gathers the results for the check; the data plane is the library's.
Checked per case: the device plan against the NumPy model (ruvector_b200.dist.HaloPlan); the
exact mode bit-identical to the single-GPU run; the overlapped multi-pass mode within 2e-5 of it
(f32 summation order) and within 1e-4 of the oracle; the peer-direct plane (sparse plans) bit-identical
to the single-GPU run; the message-row halo and host-buffer variants equal to their counterparts."""
import os
import sys
from pathlib import Path

import numpy as np
import torch
import torch.distributed as dist

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))

import ruvector_b200 as rv  # noqa: E402
from ruvector_b200 import dist as rvdist, synth  # noqa: E402
from ruvector_b200.graph import CsrGraph  # noqa: E402


def row_rel(a, b):
    return float((np.linalg.norm(a - b, axis=1) / np.maximum(np.linalg.norm(b, axis=1), 1e-6)).max())


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("gloo")
    stream = torch.cuda.Stream()
    torch.cuda.set_stream(stream)
    ctx = rv.Context(local, stream.cuda_stream)
    d_ = rvdist.Dist.from_env(ctx)
    ok = True
    cases = [(6000, 128, 4, 16, rv.MODE_FP32_SIMT, 0.0, None), (5000, 384, 8, 16, rv.MODE_FP32_TC, 0.5, None),
             (4001, 128, 4, 8, rv.MODE_FP32_TC, 0.9, None),
             (5004, 256, 8, 16, rv.MODE_FP32_SIMT, 0.0, None),  # dense plane, K|V|M halo rows
             (5004, 128, 2, 16, rv.MODE_FP32_TC, 0.0, "sparse"),  # the same kind of graph through the all-to-all-v plane
             (6000, 96, 3, 8, rv.MODE_FP32_SIMT, 0.3, None),  # generic layout: exact mode only
             # pure id-window graph: most rows are interior and run under the exchange in exact mode
             (12000, 384, 8, 16, rv.MODE_FP32_TC, 1.0, None)]
    for (n, d, heads, m, mode, loc, plane) in cases:
        layer = rv.RuvectorLayer(d, d, heads, 0.1, weights=synth.make_weights(d, d, seed=n), ctx=ctx, mode=mode)
        x = synth.make_features(n, d, seed=n + 1)
        g = synth.hnsw_like_graph(n, m, seed=n + 2, locality=loc, window=300, iso_frac=0.02)
        bounds = rvdist.even_bounds(n, world)
        lo, hi = int(bounds[rank]), int(bounds[rank + 1])
        s, e = int(g.row_ptr[lo]), int(g.row_ptr[hi])
        rows = CsrGraph((g.row_ptr[lo:hi + 1] - g.row_ptr[lo]).astype(np.uint64), g.col_idx[s:e].copy(), g.edge_w[s:e].copy())
        d_.set_chunks(3)
        part = rvdist.Partition(d_, bounds, rows.row_ptr, rows.col_idx, rows.edge_w, plane=plane)
        # the device plan against the host model
        model = rvdist.HaloPlan.build(rows, rank, bounds)
        got = part.download()
        plan_ok = part.n_halo == model.n_halo and np.array_equal(got["recv_counts"], model.recv_counts.astype(np.uint64))
        if part.dense:
            plan_ok = plan_ok and np.array_equal(got["col_idx"], rvdist.dense_columns(model, rvdist.chunk_starts(hi - lo, 3)))
        else:
            plan_ok = plan_ok and np.array_equal(got["col_idx"], model.graph.col_idx) and np.array_equal(got["need_ids"], model.need_ids.astype(np.uint32))
        xd = torch.from_numpy(x[lo:hi]).cuda()
        od, od_exact = torch.empty(hi - lo, d, device="cuda"), torch.empty(hi - lo, d, device="cuda")
        part.forward(layer, xd, od, overlap=True)
        st = d_.last_stats()
        part.forward(layer, xd, od_exact)  # default: exact
        ctx.sync()
        # message-row halo (default; K, V rebuilt on arrival) vs K|V|M rows on the wire: the owner's bits either way
        od_m, od_m_exact = torch.empty_like(od), torch.empty_like(od)
        part.forward(layer, xd, od_m, overlap=True, msg_halo=False)
        part.forward(layer, xd, od_m_exact, msg_halo=False)  # K|V|M rows on the wire instead of message rows
        ctx.sync()
        msg_ok = bool(torch.equal(od_m_exact, od_exact)) and row_rel(od_m.cpu().numpy(), od_exact.cpu().numpy()) <= 2e-5
        # peer-direct plane (sparse plans): remote rows read by the gather over NVLink, single pass = the single-GPU bits
        od_pd = torch.empty_like(od)
        part.forward(layer, xd, od_pd, peer_direct=True)
        st_pd = d_.last_stats()
        ctx.sync()
        if st_pd["peer_direct"]:
            msg_ok = msg_ok and bool(torch.equal(od_pd, od_exact))
        else:
            msg_ok = msg_ok and bool(torch.equal(od_pd, od_exact))
        out_host, out_host_exact = np.empty((hi - lo, d), np.float32), np.empty((hi - lo, d), np.float32)
        part.forward_host(layer, np.ascontiguousarray(x[lo:hi]), out_host, overlap=True)
        part.forward_host(layer, np.ascontiguousarray(x[lo:hi]), out_host_exact)
        same_host = np.array_equal(out_host, od.cpu().numpy()) and np.array_equal(out_host_exact, od_exact.cpu().numpy())
        gathered = [None] * world
        dist.all_gather_object(gathered, (od.cpu().numpy(), od_exact.cpu().numpy(), plan_ok, same_host and msg_ok))
        if rank == 0:
            from tests.oracle_binding import Oracle

            ref = Oracle().forward_graph(layer, x, g)
            full = np.concatenate([t[0] for t in gathered])
            full_exact = np.concatenate([t[1] for t in gathered])
            single = layer.forward_graph(x, g.to_device(ctx))
            bit = np.array_equal(single, full_exact)
            rel_o, rel_s = row_rel(full, ref), row_rel(full, single)
            plans, hosts = all(t[2] for t in gathered), all(t[3] for t in gathered)
            print(f"n={n} d={d} heads={heads} mode={mode} locality={loc}: world={world} dense={part.dense} "
                  f"halo_rows={part.n_halo} interior_rows={part.interior_rows} passes={st['passes']} single_pass={st['exact']} peer_direct_ran={st_pd['peer_direct']} "
                  f"overlapped: vs oracle {rel_o:.2e}, vs single GPU {rel_s:.2e}; exact bit-identical={bit}; "
                  f"plan==model {plans}; host + message-halo variants same {hosts}", flush=True)
            ok = ok and rel_o <= 1e-4 and rel_s <= 2e-5 and bit and plans and hosts
        part.close()
    flag = [ok]
    dist.broadcast_object_list(flag, src=0)
    d_.close()
    dist.destroy_process_group()
    if rank == 0:
        print("MULTI_GPU_PARITY", "OK" if flag[0] else "FAILED")
    sys.exit(0 if flag[0] else 1)


if __name__ == "__main__":
    main()
