#!/usr/bin/env python
"""Partitioned forward vs the single-graph CPU oracle.  Launch with torchrun (one rank per GPU):
  python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 \
      --master-port 29533 scripts/multi_gpu_parity.py
Every rank builds the same global graph, keeps its row range, runs project -> halo exchange
(NCCL all-to-all-v) -> aggregate/update; rank 0 gathers the rows and compares with the oracle
and (bit-exactly) with its own single-GPU run of the whole graph."""
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


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    ok = True
    for (n, d, heads, m, mode, loc) in [(6000, 128, 4, 16, rv.MODE_FP32_SIMT, 0.0), (5000, 384, 8, 16, rv.MODE_FP32_TC, 0.5),
                                        (4001, 128, 4, 8, rv.MODE_FP32_TC, 0.9),
                                        (5004, 256, 8, 16, rv.MODE_FP32_SIMT, 0.0),  # dense plane, K|V|M halo rows rebuilt
                                        # pure id-window graph: most rows are interior and run under the exchange
                                        (12000, 384, 8, 16, rv.MODE_FP32_TC, 1.0)]:
        stream = torch.cuda.Stream()
        torch.cuda.set_stream(stream)
        ctx = rv.Context(local, stream.cuda_stream)
        layer = rv.RuvectorLayer(d, d, heads, 0.1, weights=synth.make_weights(d, d, seed=n), ctx=ctx, mode=mode)
        x = synth.make_features(n, d, seed=n + 1)
        g = synth.hnsw_like_graph(n, m, seed=n + 2, locality=loc, window=300, iso_frac=0.02)
        bounds = rvdist.even_bounds(n, world)
        lo, hi = int(bounds[rank]), int(bounds[rank + 1])
        s, e = int(g.row_ptr[lo]), int(g.row_ptr[hi])
        rows = CsrGraph((g.row_ptr[lo:hi + 1] - g.row_ptr[lo]).astype(np.uint64), g.col_idx[s:e].copy(), g.edge_w[s:e].copy())
        plan = rvdist.exchange_requests(rvdist.HaloPlan.build(rows, rank, bounds))
        dense = rvdist.choose_dense(plan)  # locality 0 -> all-gather plane, 0.9 -> all-to-all-v
        part = rvdist.Partition(plan, ctx, dense=dense)
        xd = torch.from_numpy(x[lo:hi]).cuda()
        od = torch.empty(hi - lo, d, device="cuda")
        part.forward(layer, xd, od)
        torch.cuda.synchronize()
        # same rows without the interior/boundary split (exchange not overlapped): identical bits
        od2 = torch.empty_like(od)
        rvdist.Partition(plan, ctx, dense=dense, overlap=False).forward(layer, xd, od2)
        torch.cuda.synchronize()
        same_split = bool(torch.equal(od, od2)) and (loc < 1.0 or part.interior_rows > 0)
        sizes = [int(bounds[p + 1] - bounds[p]) for p in range(world)]
        full = torch.empty(n, d, device="cuda")
        dist.all_gather(list(full.split(sizes)), od)
        # host-buffer variant of the same partitioned call
        out_host = np.empty((hi - lo, d), dtype=np.float32)
        part.forward_host(layer, np.ascontiguousarray(x[lo:hi]), out_host)
        same_host = np.array_equal(out_host, od.cpu().numpy())
        if rank == 0:
            from tests.oracle_binding import Oracle

            ref = Oracle().forward_graph(layer, x, g)
            got = full.cpu().numpy()
            rel = np.linalg.norm(got - ref, axis=1) / np.maximum(np.linalg.norm(ref, axis=1), 1e-6)
            single = layer.forward_graph(x, g.to_device(ctx))
            bit = np.array_equal(single, got)
            print(f"n={n} d={d} mode={mode} locality={loc}: world={world} dense={dense} halo_rows={part.plan.n_halo} "
                  f"interior_rows={part.interior_rows} max_row_rel={rel.max():.2e} bit_identical_to_single_gpu={bit} "
                  f"host_variant_same={same_host}")
            ok = ok and rel.max() <= 1e-4 and bit and same_host
        ok = ok and same_split
        torch.cuda.set_stream(torch.cuda.default_stream())
    flag = torch.tensor([1 if ok else 0], device="cuda")
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    dist.destroy_process_group()
    if rank == 0:
        print("MULTI_GPU_PARITY", "OK" if ok else "FAILED")
    sys.exit(0 if int(flag.item()) else 1)


if __name__ == "__main__":
    main()
