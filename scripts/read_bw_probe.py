#!/usr/bin/env python
"""Read-only vs copy HBM bandwidth on this GPU (run under gpurun).  The roofline denominator of K-agg is the
driver-measured COPY bandwidth (MEASURED_PEAKS.json: read + write); K-agg is ~98% reads, so this probe puts a
read-only number next to it.  Library kernels only (torch reductions / copies),
CUDA events, inputs far larger than L2.
  python scripts/read_bw_probe.py [GiB]"""
import json
import sys

import torch


def timed(fn, iters=10):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    best = float("inf")
    for _ in range(iters):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    return best


def main():
    gib = float(sys.argv[1]) if len(sys.argv) > 1 else 8.0
    n = int(gib * (1 << 30)) // 4
    x = torch.empty(n, dtype=torch.float32, device="cuda").normal_()
    y = torch.empty_like(x)
    out = {"bytes": n * 4}
    ms = timed(lambda: torch.sum(x))
    out["read_only_sum_gbs"] = n * 4 / (ms * 1e-3) / 1e9
    ms = timed(lambda: torch.amax(x))
    out["read_only_amax_gbs"] = n * 4 / (ms * 1e-3) / 1e9
    ms = timed(lambda: y.copy_(x))
    out["copy_read_plus_write_gbs"] = 2 * n * 4 / (ms * 1e-3) / 1e9
    ms = timed(lambda: y.zero_())
    out["write_only_gbs"] = n * 4 / (ms * 1e-3) / 1e9
    print(json.dumps(out))


if __name__ == "__main__":
    main()
