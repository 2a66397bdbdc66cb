#!/usr/bin/env python
"""Phase trace of the dense (message-only, chunked all-gather) halo plane.  torchrun, one rank per GPU:
  python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 \
      --master-port 29512 scripts/dense_plane_trace.py [--nodes 1000000] [--chunks 1,4,8] [--comm-sms 0,16,32]
Prints rank 0's per-phase device times (CUDA events on the launching stream) per variant."""
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
from ruvector_b200 import dist as rvdist, synth  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--nodes", type=int, default=1000000)
    ap.add_argument("--dim", type=int, default=384)
    ap.add_argument("--heads", type=int, default=8)
    ap.add_argument("--m", type=int, default=32)
    ap.add_argument("--chunks", default="1,4,8")
    ap.add_argument("--comm-sms", default="0,16,32")
    a = ap.parse_args()
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    stream = torch.cuda.Stream()
    torch.cuda.set_stream(stream)
    ctx = rv.Context(local, stream.cuda_stream)
    layer = rv.RuvectorLayer(a.dim, a.dim, a.heads, 0.1, weights=synth.make_weights(a.dim, a.dim), ctx=ctx,
                             mode=rv.MODE_FP32_TC)
    x = torch.from_numpy(synth.make_features(a.nodes, a.dim, seed=rank)).cuda()
    out = torch.empty(a.nodes, a.dim, device="cuda")
    g = synth.hnsw_like_graph(a.nodes, a.m, seed=0xBEEF + rank, n_total=a.nodes * world, row_offset=rank * a.nodes)
    plan = rvdist.exchange_requests(rvdist.HaloPlan.build(g, rank, rvdist.even_bounds(a.nodes * world, world)))
    for chunks in map(int, a.chunks.split(",")):
        for sms in map(int, a.comm_sms.split(",")):
            part = rvdist.Partition(plan, ctx, dense=True, chunks=chunks, comm_sms=sms)
            for _ in range(2):
                part.forward(layer, x, out)
            dist.barrier()
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(3):
                part.forward(layer, x, out)
            e1.record()
            torch.cuda.synchronize()
            dist.barrier()
            part.trace = []
            part.forward(layer, x, out)
            tr = part.trace_ms()
            if rank == 0:
                print(f"chunks={chunks} comm_sms={sms} NCCL_MAX_CTAS={os.environ.get('NCCL_MAX_CTAS')} "
                      f"ms/step={e0.elapsed_time(e1) / 3:.2f} trace={tr}", flush=True)
            del part
            torch.cuda.empty_cache()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
