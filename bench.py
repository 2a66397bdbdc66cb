#!/usr/bin/env python
"""bench.py -- one GNN layer (RuvectorLayer forward) over an HNSW-shaped graph on B200.

A "step" is one full layer pass over the whole graph: K-proj -> K-agg -> K-upd.
Workload at N=1 = BASELINE.json configs[1]: 1M nodes, d=384, HNSW M=32 (level-0 out-degree
U{32..64}), 8 heads, single layer, synthetic seeded inputs (ruvector_b200/synth.py).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

Prints ONE JSON line (rank 0).  `value` = node-updates/s with inputs resident in HBM;
`e2e` = the same through the host-buffer C-ABI call (H2D of X and D2H of the output inside
the timed region).  `roofline` describes the dominant kernel (K-agg, HBM-bound);
`cpu_baseline` is the CPU restatement of the reference's per-node loop timed on this
box's host cores on a bounded node sample.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))

METRIC = "gnn_layer_node_updates_per_sec"
UNIT = "node-updates/s"


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--nodes", type=int, default=1_000_000, help="nodes per GPU")
    ap.add_argument("--dim", type=int, default=384)
    ap.add_argument("--heads", type=int, default=8)
    ap.add_argument("--m", type=int, default=32)
    ap.add_argument("--locality", type=float, default=0.0)
    ap.add_argument("--window", type=int, default=4096, help="id window of the local neighbours (with --locality)")
    ap.add_argument("--no-overlap", action="store_true", help="N>1: do not run interior rows under the halo exchange")
    ap.add_argument("--mode", type=int, default=-1, help="-1: best parity mode available")
    ap.add_argument("--cpu-seconds", type=float, default=15.0)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    return ap.parse_args()


def workload_name(a, n_gpus):
    return (f"{a.nodes * n_gpus} nodes d={a.dim} HNSW M={a.m} (level-0 deg U{{{a.m}..{2 * a.m}}}), "
            f"{a.heads} heads, 1 layer")


def measured_peaks():
    p = ROOT / "MEASURED_PEAKS.json"
    if p.exists():
        d = json.loads(p.read_text())
        return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """Samples nvidia-smi clocks / throttle reasons during the timed region."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.rows, self.proc, self.gpu = [], None, gpu_index

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100",
                 "-i", str(self.gpu)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self) -> dict:
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except subprocess.TimeoutExpired:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        for r in self.rows:
            if len(r) < 9:
                continue
            try:
                sm.append(float(r[1]))
                mx.append(float(r[2]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def agg_bytes_per_node(H: int, deg_sum: int, n: int, R: int = 3, heads: int = 1) -> float:
    """Algorithmic bytes of K-agg (SURVEY.md 8d / DESIGN.md).
    R=3: per node 4H (Q) + 16 (row_ptr pair) + deg*(4 col + 4 w + 12H) + 8H (o, agg).
    R=1: per node 4hH (folded queries) + 16 + deg*(8 + 4H) + 4(h+1)H (S_0..S_{h-1}, agg)."""
    if R == 1:
        return (n * (4 * heads * H + 16 + 4 * (heads + 1) * H) + deg_sum * (8 + 4 * H)) / n
    return (n * (4 * H + 16 + 8 * H) + deg_sum * (8 + 12 * H)) / n


def cpu_baseline(a, layer, x, g, seconds, threads=None):
    """Times the CPU restatement of the reference's per-node loop (per-edge re-projection
    included) on a bounded node sample, all host threads."""
    from tests.oracle_binding import Oracle

    orc = Oracle()
    threads = threads or os.cpu_count() or 1
    n = x.shape[0]
    probe = min(n, 16 * threads)
    t0 = time.perf_counter()
    orc.forward_graph(layer, x, g, threads=threads, node_range=(0, probe))
    dt = max(time.perf_counter() - t0, 1e-6)
    sample = int(min(n, max(probe, probe * seconds / dt)))
    t0 = time.perf_counter()
    orc.forward_graph(layer, x, g, threads=threads, node_range=(0, sample))
    dt = time.perf_counter() - t0
    return {"value": sample / dt, "unit": UNIT, "cores": threads, "kind": "port",
            "sample": f"first {sample} nodes of the same graph, faithful per-node loop, {dt:.1f} s"}, sample, dt


def build_inputs(a, rank, world):
    from ruvector_b200 import synth

    x = synth.make_features(a.nodes, a.dim, seed=0x5EED + rank)
    # rank r owns rows [r*nodes, (r+1)*nodes) of a (world*nodes)-node graph; columns are global
    g = synth.hnsw_like_graph(a.nodes, a.m, seed=0xBEEF + rank, locality=a.locality, window=a.window,
                              n_total=a.nodes * world, row_offset=rank * a.nodes)
    w = synth.make_weights(a.dim, a.dim)
    return x, g, w


def run_reference(a):
    """--impl reference: the reference's own CPU algorithm (restated; the reference is Rust and
    cannot be built here) on this box's host cores, bounded sample per step."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import ruvector_b200 as rv
    from tests.oracle_binding import Oracle

    nodes = min(a.nodes, 200_000)  # the sample only touches a prefix; keep generation cheap
    aa = argparse.Namespace(**{**vars(a), "nodes": nodes})
    x, g, w = build_inputs(aa, 0, 1)
    layer = rv.RuvectorLayer(a.dim, a.dim, a.heads, 0.1, weights=w)
    orc = Oracle()
    threads = os.cpu_count() or 1
    probe = 8 * threads
    t0 = time.perf_counter()
    orc.forward_graph(layer, x, g, threads=threads, node_range=(0, probe))
    per_node = (time.perf_counter() - t0) / probe
    budget = 120.0 / max(a.steps + a.warmup, 1)
    sample = int(max(threads, min(nodes, budget / per_node)))
    for _ in range(a.warmup):
        orc.forward_graph(layer, x, g, threads=threads, node_range=(0, sample))
    t0 = time.perf_counter()
    for _ in range(a.steps):
        orc.forward_graph(layer, x, g, threads=threads, node_range=(0, sample))
    dt = time.perf_counter() - t0
    val = sample * a.steps / dt
    line = {
        "impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": a.gpus, "steps": a.steps,
        "warmup": a.warmup, "ms_per_step": 1e3 * dt / a.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": workload_name(a, a.gpus), "sampled_nodes_per_step": sample},
        "cpu_baseline": {"value": val, "unit": UNIT, "cores": threads, "kind": "port",
                         "sample": f"{sample} nodes per step of the same synthetic graph (faithful per-node loop)"},
        "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "edges_per_sec": val * float(g.row_ptr[sample]) / sample,
    }
    print(json.dumps(line))


def main():
    a = parse()
    if a.impl == "reference":
        return run_reference(a)

    import torch
    import torch.distributed as dist

    import ruvector_b200 as rv
    from ruvector_b200 import _lib

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (no CPU fallback)")
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    n_gpus = world

    x, g, w = build_inputs(a, rank, world)
    # a dedicated non-default stream, made current, so torch's events and the library's
    # kernels are on the same stream (handle 0 would make the library create its own)
    tstream = torch.cuda.Stream()
    torch.cuda.set_stream(tstream)
    ctx = rv.Context(local, tstream.cuda_stream)
    layer = rv.RuvectorLayer(a.dim, a.dim, a.heads, 0.1, weights=w, ctx=ctx)
    mode = a.mode
    if mode < 0:
        mode = _lib.MODE_FP32_SIMT
        try:
            layer.set_mode(_lib.MODE_FP32_TC)
            layer._device_layer()
            mode = _lib.MODE_FP32_TC
        except rv.GnnError:
            layer.set_mode(_lib.MODE_FP32_SIMT)
    else:
        layer.set_mode(mode)
    n, H = a.nodes, a.dim
    E = g.n_edges

    if world > 1:
        from ruvector_b200 import dist as rvdist

        part = rvdist.Partition.build(g, rank, world, a.nodes, ctx, overlap=not a.no_overlap)

        def step_fn(xd, od, timing=False):
            ctx.set_timing(timing)
            part.forward(layer, xd, od)
        dg = part.graph
    else:
        dg = g.to_device(ctx)
        step_fn = lambda xd, od, timing=False: layer.forward_graph_device(xd.data_ptr(), od.data_ptr(), dg,
                                                                          timing=timing)

    x_pin = torch.from_numpy(x).pin_memory()
    x_dev = x_pin.cuda(non_blocking=True)
    out_dev = torch.empty(n, H, device="cuda", dtype=torch.float32)
    torch.cuda.synchronize()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- device-resident leg: `value` --------------------------------------------
    for _ in range(max(a.warmup, 3)):
        step_fn(x_dev, out_dev)
    barrier()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    launches0 = ctx.launch_count
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    stage = np.zeros(3)
    ev0.record()
    for _ in range(a.steps):
        step_fn(x_dev, out_dev, timing=True)
        if world == 1:
            pass
    ev1.record()
    barrier()
    ms = ev0.elapsed_time(ev1)
    launches = ctx.launch_count - launches0
    # per-stage device times of the last timed step (events on the launching stream)
    exchange_ms = None
    try:
        t = ctx.last_timing()
        stage = np.array([t["project_ms"], t["aggregate_ms"], t["update_ms"]])
        exchange_ms = t["exchange_ms"]
    except rv.GnnError:
        pass
    if world > 1 and part.interior_rows > 0:
        # the overlapped step runs K-agg once per row range, so its stage events cover the last
        # range only: take the per-kernel times from one extra un-split pass over the same rows
        whole = rvdist.Partition(part.plan, ctx, dense=part.dense, overlap=False)
        for _ in range(2):
            ctx.set_timing(True)
            whole.forward(layer, x_dev, out_dev)
        barrier()
        t = ctx.last_timing()
        stage = np.array([t["project_ms"], t["aggregate_ms"], t["update_ms"]])
        exchange_ms = whole.exposed_wait_ms()
        del whole
    exposed_wait_ms = part.exposed_wait_ms() if world > 1 else None  # of the last device-resident step
    ctx.set_timing(False)
    clocks = sampler.stop() if rank == 0 else None
    if world > 1:
        tmax = torch.tensor([ms], device="cuda")
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
        ms = float(tmax.item())
        tot_e = torch.tensor([float(E)], device="cuda", dtype=torch.float64)
        dist.all_reduce(tot_e)
        E_total = float(tot_e.item())
    else:
        E_total = float(E)
    ms_per_step = ms / a.steps
    value = n * n_gpus / (ms_per_step * 1e-3)

    # ---- e2e leg: host buffers through the C ABI (H2D + D2H inside the timed region) ----
    e2e = None
    if not a.no_e2e:
        out_pin = torch.empty(n, H, dtype=torch.float32).pin_memory()
        x_np, out_np = x_pin.numpy(), out_pin.numpy()
        if world > 1:
            e2e_fn = lambda: part.forward_host(layer, x_np, out_np)
        else:
            e2e_fn = lambda: layer.forward_graph(x_np, dg, out=out_np)
        e2e_fn()
        barrier()
        k2 = max(2, min(a.steps, 5))
        t0 = time.perf_counter()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(k2):
            e2e_fn()
        e1.record()
        barrier()
        e_ms = max(e0.elapsed_time(e1), 0.0)
        wall_ms = (time.perf_counter() - t0) * 1e3
        e_ms = max(e_ms, wall_ms) if e_ms <= 0 else e_ms
        if world > 1:
            tm = torch.tensor([e_ms], device="cuda")
            dist.all_reduce(tm, op=dist.ReduceOp.MAX)
            e_ms = float(tm.item())
        e2e = {"value": n * n_gpus / (e_ms / k2 * 1e-3), "unit": UNIT,
               "h2d_bytes_per_step": int(x.nbytes), "d2h_bytes_per_step": int(n * H * 4),
               "ms_per_step": e_ms / k2, "steps": k2}

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    peak, peak_src = measured_peaks()
    agg_ms = float(stage[1]) if stage[1] > 0 else None
    R = int(_lib.load().ruvgnn_layer_halo_width(layer._device_layer())) // H
    bytes_launch = agg_bytes_per_node(H, E, n, R, a.heads) * n
    roof = {"kernel": f"{'gather_r1_batched_kernel' if R == 1 else 'gather_fast_kernel'} (K-agg, R={R})", "bound": "hbm",
            "achieved": (bytes_launch / (agg_ms * 1e-3) / 1e9) if agg_ms else None, "peak": peak, "unit": "GB/s",
            "frac": (bytes_launch / (agg_ms * 1e-3) / 1e9 / peak) if agg_ms else None,
            "traffic": None, "peak_source": peak_src,
            "frac_of_nominal_8TBs": (bytes_launch / (agg_ms * 1e-3) / 1e9 / 8000.0) if agg_ms else None, "algorithmic_bytes_per_launch": bytes_launch,
            "kernel_ms": agg_ms, "stage_ms": {"project": float(stage[0]), "aggregate": float(stage[1]),
                                               "update": float(stage[2])}}
    traffic_file = ROOT / "profiles" / "kagg_traffic.json"
    if traffic_file.exists():
        try:
            tr = json.loads(traffic_file.read_text())
            if tr.get("nodes") == n and tr.get("dim") == H:
                roof["traffic"] = tr.get("dram_bytes_per_launch")
        except (ValueError, OSError):
            pass

    # Second roofline: the six projection GEMMs against the tensor pipe.  USEFUL FLOPs per node
    # = 2 (D H + 11 H^2); the FP32-parity path issues three TF32 MMAs per product (3xTF32), so the pipe does
    # 3x that.  Peak = half the measured dense bf16 rate (TF32 runs at half the bf16 rate).  The time is the
    # K-proj + K-upd stage time of the last step; K-upd includes the LayerNorm kernel (~1% of the layer).
    gemm_ms = float(stage[0] + stage[2])
    D = a.dim
    flops_useful = 2.0 * (D * H + 11.0 * H * H) * n
    passes = 3 if mode == _lib.MODE_FP32_TC else 1
    tf32_peak, tf32_src = None, None
    try:
        mp = json.loads((ROOT / "MEASURED_PEAKS.json").read_text())
        tf32_peak, tf32_src = float(mp["bf16_tflops"]) / 2.0, "measured bf16_tflops (burst) / 2"
    except (OSError, ValueError, KeyError):
        tf32_peak, tf32_src = 1125.0, "nominal 2.25 PFLOP/s bf16 / 2"
    if gemm_ms > 0 and mode != _lib.MODE_FP32_SIMT:
        useful = flops_useful / (gemm_ms * 1e-3) / 1e12
        roof_gemm = {"kernels": "gemm_tc_kernel / gemm_tc_ts_kernel (K-proj + K-upd, 6 GEMMs)", "bound": "tensor",
                     "achieved": useful * passes, "achieved_useful": useful, "peak": tf32_peak, "unit": "TFLOP/s",
                     "frac": useful * passes / tf32_peak, "frac_useful": useful / tf32_peak,
                     "mma_passes_per_product": passes, "peak_source": tf32_src, "ms": gemm_ms}
    else:
        roof_gemm = None

    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": n_gpus, "steps": a.steps, "warmup": max(a.warmup, 3),
        "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": {"workload": workload_name(a, n_gpus), "nodes_per_gpu": n, "dim": H, "heads": a.heads, "m": a.m,
                   "edges": E_total, "locality": a.locality, "window": a.window, "l2": "inputs larger than L2 (X, projected rows >> 126 MB)",
                   "projection_mode": {0: "fp32_simt", 1: "fp32_3xtf32_tcgen05", 2: "tf32_tcgen05"}[mode],
                   "parallelism": f"range-partition x{n_gpus}" if n_gpus > 1 else "single"},
        "edges_per_sec": E_total / (ms_per_step * 1e-3),
        "clocks": clocks, "e2e": e2e, "gpu_launches": int(launches), "roofline": roof,
        "roofline_projections": roof_gemm,
    }
    if world > 1:
        line["halo"] = {"rows_received_rank0": part.plan.n_halo, "rows_sent_rank0": int(part.plan.send_idx.size),
                        "bytes_exchanged_rank0": int(part.bytes_exchanged), "exchange_ms_rank0": exchange_ms,
                        "plane": "all-gather" if part.dense else "all-to-all-v",
                        "interior_rows_rank0": int(part.interior_rows), "overlap": not a.no_overlap,
                        "exposed_wait_ms_rank0": round(exposed_wait_ms, 3)}
    if not a.no_cpu_baseline and world == 1:
        cb, _, _ = cpu_baseline(a, layer, x, g, a.cpu_seconds)
        line["cpu_baseline"] = cb
        # the reference itself is single-threaded (no rayon / SIMD use in crates/ruvector-gnn/src):
        # the same faithful loop on ONE host thread, short sample
        c1, _, _ = cpu_baseline(a, layer, x, g, 4.0, threads=1)
        line["cpu_baseline_single_thread"] = c1
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
