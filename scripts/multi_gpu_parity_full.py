#!/usr/bin/env python
"""Config-5 slice at full size (12.5M nodes x 128 x 4 heads per GPU; 100M nodes on 8 GPUs), one layer,
sampled rows against the oracle.  Launch with torchrun, one rank per GPU:
  python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 \
      --master-port 29533 scripts/multi_gpu_parity_full.py [--nodes 12500000] [--locality 0.0]
Every rank generates its rows on the device (synth_cuda), the library partitions and runs the layer
(overlapped multi-pass mode and exact mode); rank 0 regenerates the feature rows of its sampled nodes'
neighbours (whichever rank owns them), runs the oracle's per-node forward (the reference's own call
pattern, layer.rs:379-418) and compares."""
import argparse
import os
import sys
from pathlib import Path

import numpy as np
import torch
import torch.distributed as dist

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))

import ruvector_b200 as rv  # noqa: E402
from ruvector_b200 import dist as rvdist, synth, synth_cuda  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--nodes", type=int, default=12_500_000)
    ap.add_argument("--locality", type=float, default=0.0)
    ap.add_argument("--samples", type=int, default=48)
    a = ap.parse_args()
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("gloo")
    stream = torch.cuda.Stream()
    torch.cuda.set_stream(stream)
    ctx = rv.Context(local, stream.cuda_stream)
    d_ = rvdist.Dist.from_env(ctx)
    n, d, heads, m = a.nodes, 128, 4, 32
    slab = 1 << 21
    n_pad = (n + slab - 1) // slab * slab
    layer = rv.RuvectorLayer(d, d, heads, 0.1, weights=synth.make_weights(d, d), ctx=ctx)
    x = synth_cuda.make_features(n, d, seed=0x5EED, row_offset=rank * n_pad)
    rp, col, w = synth_cuda.hnsw_like_graph(n, m, seed=0xBEEF + rank, locality=a.locality, n_total=n * world, row_offset=rank * n)
    part = rvdist.Partition(d_, rvdist.even_bounds(n * world, world), rp, col, w, device_ptrs=True)
    out, out_exact = torch.empty(n, d, device="cuda"), torch.empty(n, d, device="cuda")
    part.forward(layer, x, out, overlap=True, asynchronous=False)
    st = d_.last_stats()
    part.forward(layer, x, out_exact, asynchronous=False)  # default: exact
    ok = True
    if rank == 0:
        from tests.oracle_binding import Oracle

        orc = Oracle()
        rng = np.random.default_rng(11)
        deg = (rp[1:] - rp[:-1])
        iso = torch.nonzero(deg == 0).squeeze(1)[:2].cpu().numpy()
        sample = np.unique(np.concatenate([rng.integers(0, n, size=a.samples), [0, n - 1], iso]))
        worst = worst_exact = 0.0
        remote = 0
        for i in sample:
            b, e = int(rp[i]), int(rp[i + 1])
            nb = col[b:e].cpu().numpy().astype(np.int64)  # GLOBAL ids
            ws = w[b:e].cpu().numpy()
            owner, loc = nb // n, nb % n
            remote += int((owner != 0).sum())
            xn = synth_cuda.feature_rows(owner * n_pad + loc, d, seed=0x5EED).cpu().numpy() if nb.size else np.zeros((0, d), np.float32)
            ref = orc.forward_node(layer, x[i].cpu().numpy(), xn, ws)
            den = max(np.linalg.norm(ref), 1e-6)
            worst = max(worst, float(np.linalg.norm(out[i].cpu().numpy() - ref) / den))
            worst_exact = max(worst_exact, float(np.linalg.norm(out_exact[i].cpu().numpy() - ref) / den))
        ok = worst <= 1e-4 and worst_exact <= 1e-4
        print(f"config-5 slice: {n * world} nodes over {world} GPUs ({n} per GPU), d={d}, {heads} heads, locality {a.locality}: "
              f"plane={'all-gather' if part.dense else 'all-to-all-v'} halo_rows={part.n_halo} passes={st['passes']}; "
              f"{sample.size} sampled rows of rank 0 ({remote} remote neighbours) vs the oracle: overlapped {worst:.2e}, "
              f"exact {worst_exact:.2e} (bound 1e-4)", flush=True)
    flag = [ok]
    dist.broadcast_object_list(flag, src=0)
    part.close()
    d_.close()
    dist.destroy_process_group()
    if rank == 0:
        print("FULL_SIZE_PARITY", "OK" if flag[0] else "FAILED")
    sys.exit(0 if flag[0] else 1)


if __name__ == "__main__":
    main()
