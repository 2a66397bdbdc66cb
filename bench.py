#!/usr/bin/env python
"""bench.py -- one GNN layer (RuvectorLayer forward) over an HNSW-shaped graph on B200.

A "step" is one full layer pass over the whole graph: K-proj -> K-agg -> K-upd.
Workload at N=1 = BASELINE.json configs[1]: 1M nodes, d=384, HNSW M=32 (level-0 out-degree
U{32..64}), 8 heads, single layer, synthetic seeded inputs (ruvector_b200/synth.py); at N>1 the same
shape per GPU (weak scaling), range-partitioned, halo exchange inside the library (NCCL).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

Prints ONE JSON line (rank 0).  `value` = node-updates/s with inputs resident in HBM;
`e2e` = the same through the host-buffer C-ABI call (H2D of X and D2H of the output inside
the timed region).  `roofline` describes the dominant kernel (K-agg, HBM-bound);
`roofline_projections` the six GEMMs against min(tensor roof, HBM roof x intensity);
`cpu_baseline` is the CPU restatement of the reference's per-node loop timed on this
box's host cores on a bounded node sample.

`secondary` (same run, same JSON line) carries the north-star d=128 configurations:
  N=1: config 3 -- 10M nodes x 128, 4 heads, three stacked layers over HNSW levels 2, 1, 0;
  N>1: the per-GPU slice of config 5 -- 12.5M nodes x 128 per GPU (100M nodes at N=8), at
       locality 0 (insertion-order worst case) and 0.9 (after a locality reorder).
"""
from __future__ import annotations

import argparse
import hashlib
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
    ap.add_argument("--exact", action="store_true", help="N>1: (default) single-pass gather after the halo landed, bit-identical to 1 GPU")
    ap.add_argument("--overlap", action="store_true", help="N>1: multi-pass gather overlapped with the exchange (f32 summation order differs)")
    ap.add_argument("--chunks", type=int, default=0, help="N>1: row chunks of the dense plane (0: the library's choice from the world size)")
    ap.add_argument("--plane", default=None, choices=["sparse", "dense"], help="N>1: force the halo plan's data plane")
    ap.add_argument("--peer-direct", action="store_true", help="N>1, sparse plans: the gather reads remote rows from the peers' tables over NVLink")
    ap.add_argument("--row-halo", action="store_true", help="N>1, K|V|M layers, dense plane: ship K|V|M rows instead of message rows + rebuild")
    ap.add_argument("--mode", type=int, default=-1, help="-1: the library default (tcgen05 3xTF32 where the shape allows)")
    ap.add_argument("--cpu-seconds", type=float, default=15.0)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-secondary", action="store_true")
    ap.add_argument("--secondary-nodes", type=int, default=0, help="0: 10M at N=1 (config 3), 12.5M per GPU at N>1 (config 5 slice)")
    ap.add_argument("--secondary-only", action="store_true", help="skip the config-2 legs (experiments)")
    return ap.parse_args()


def workload_name(a, n_gpus):
    return (f"{a.nodes * n_gpus} nodes d={a.dim} HNSW M={a.m} (level-0 deg U{{{a.m}..{2 * a.m}}}), "
            f"{a.heads} heads, 1 layer")


def peaks():
    """(hbm GB/s, source, tf32 TFLOP/s, source) from MEASURED_PEAKS.json, else the profiling guide's fallback."""
    p = ROOT / "MEASURED_PEAKS.json"
    try:
        d = json.loads(p.read_text())
        return (float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)", float(d["bf16_tflops"]) / 2.0,
                "measured bf16_tflops (burst) / 2")
    except (OSError, ValueError, KeyError):
        return 6650.0, "fallback (B200_PROFILING.md)", 1125.0, "nominal 2.25 PFLOP/s bf16 / 2"


def active_knobs() -> dict:
    """Every RUVGNN_* / NCCL_* environment knob that is set: they steer the hot path, so they are part of the config."""
    return {k: v for k, v in sorted(os.environ.items()) if k.startswith("RUVGNN_") or k.startswith("NCCL_")}


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


# ---- algorithmic work (SURVEY.md 8d / DESIGN.md) -------------------------------------------------
def agg_bytes(H: int, deg_sum: float, n: float, R: int, heads: int) -> float:
    """Algorithmic bytes of one K-agg pass over n nodes with deg_sum edges.
    R=3: per node 4H (Q) + 16 (row_ptr pair) + deg*(4 col + 4 w + 12H) + 8H (o, agg).
    R=1: per node 4hH (folded queries) + 16 + deg*(8 + 4H) + 4(h+1)H (S_0..S_{h-1}, agg)."""
    if R == 1:
        return n * (4 * heads * H + 16 + 4 * (heads + 1) * H) + deg_sum * (8 + 4 * H)
    return n * (4 * H + 16 + 8 * H) + deg_sum * (8 + 12 * H)


def proj_bytes_per_node(D: int, H: int, R: int, heads: int) -> float:
    """HBM bytes the projections must move per node (SURVEY 8d whole-layer formula minus K-agg):
    read X, write M and Q|K|V (or the folded queries), read o|agg (or S_g|agg) and m, write out.
    Weights are L2 / SMEM resident and excluded."""
    if R == 1:
        return 4 * D + 4 * H * (1 + heads) + 4 * H * (heads + 1) + 4 * H + 4 * H
    return 4 * D + 4 * H * 4 + 12 * H + 4 * H


def kagg_roofline(H, heads, R, E, n, agg_ms, peak, peak_src, stage):
    bytes_launch = agg_bytes(H, E, n, R, heads)
    ach = (bytes_launch / (agg_ms * 1e-3) / 1e9) if agg_ms else None
    return {"kernel": f"{'gather_r1_batched_kernel' if R == 1 else 'gather_fast_kernel'} (K-agg, R={R})", "bound": "hbm",
            "achieved": ach, "peak": peak, "unit": "GB/s", "frac": (ach / peak) if ach else None, "traffic": None,
            "peak_source": peak_src, "frac_of_nominal_8TBs": (ach / 8000.0) if ach else None,
            "algorithmic_bytes_per_launch": bytes_launch, "kernel_ms": agg_ms,
            "stage_ms": {"project": float(stage[0]), "aggregate": float(stage[1]), "update": float(stage[2])}}


def gemm_roofline(D, H, heads, R, n, gemm_ms, passes, hbm_peak, tf32_peak, tf32_src, msg_only_rows=0.0,
                  update_launches=5):
    """The six projection GEMMs against min(tensor roof, HBM roof x intensity).  USEFUL FLOPs per node
    = 2 (D H + 11 H^2) (the reference's 12 Linears); the FP32-parity path issues `passes` TF32 MMAs per product and,
    in the single-row formulation, folded matrices (issued_flops_per_node).  `msg_only_rows`: rows without
    neighbours on a sparse level, which only go through W_msg and the LayerNorm (layer.rs:385-389)."""
    if not gemm_ms or gemm_ms <= 0:
        return None
    useful_node = 2.0 * (D * H + 11.0 * H * H)
    units = (2 * heads + 7) if R == 1 else 11  # H x H units: W_qk, W_ov (h each), W_agg, gates (4), candidate (2)
    issued_node = 2.0 * (D * H + units * H * H) * passes
    bytes_node = proj_bytes_per_node(D, H, R, heads)
    useful_total = useful_node * n + 2.0 * D * H * msg_only_rows
    issued_total = issued_node * n + 2.0 * D * H * passes * msg_only_rows
    bytes_total = bytes_node * n + (4.0 * D + 12.0 * H) * msg_only_rows
    useful = useful_total / (gemm_ms * 1e-3) / 1e12
    issued = issued_total / (gemm_ms * 1e-3) / 1e12
    intensity = issued_total / bytes_total  # MMA FLOPs the pipe must issue per HBM byte
    hbm_bound_tflops = hbm_peak * 1e9 * intensity / 1e12
    bound = min(tf32_peak, hbm_bound_tflops)
    kernels = ("gemm_tc_ts_kernel (K-proj, 2 GEMMs) + upd_fused_kernel (K-upd: out-proj, w_agg, GRU, dropout scale and "
               "LayerNorm in one tcgen05 kernel)" if update_launches == 1 else
               "gemm_tc_kernel / gemm_tc_ts_kernel (K-proj + K-upd, 6 GEMMs + LayerNorm)")
    return {"kernels": kernels, "update_launches_per_block": update_launches,
            "bound": "tensor" if tf32_peak <= hbm_bound_tflops else "hbm", "unit": "TFLOP/s",
            "achieved": issued, "achieved_useful": useful, "peak": bound, "frac": issued / bound,
            "frac_useful": useful / bound, "tensor_peak": tf32_peak, "hbm_bound_tflops": hbm_bound_tflops,
            "intensity_flop_per_byte": intensity, "algorithmic_bytes_per_node": bytes_node,
            "achieved_GBs": bytes_total / (gemm_ms * 1e-3) / 1e9, "frac_of_hbm": bytes_total / (gemm_ms * 1e-3) / 1e9 / hbm_peak,
            "useful_flops_per_node": useful_node, "issued_flops_per_node": issued_node,
            "rows_full_chain": n, "rows_message_only": msg_only_rows,
            "mma_passes_per_product": passes, "peak_source": tf32_src, "ms": gemm_ms}


def traffic_from_profile(n, H, heads):
    """`dram__bytes_read.sum + dram__bytes_write.sum` of one K-agg launch from the committed ncu capture, only
    when it was taken on THIS gather.cu (source hash) and this shape; otherwise null."""
    f = ROOT / "profiles" / "kagg_traffic.json"
    try:
        entries = json.loads(f.read_text())
        entries = entries if isinstance(entries, list) else [entries]
        want = hashlib.sha256((ROOT / "ruvector_b200" / "csrc" / "gather.cu").read_bytes()).hexdigest()
        for t in entries:
            if t.get("nodes") == n and t.get("dim") == H and t.get("heads", heads) == heads:
                if t.get("gather_cu_sha256") == want:
                    return t.get("dram_bytes_per_launch"), t.get("capture")
                return None, f"stale capture {t.get('capture')}: gather.cu changed since"
    except (ValueError, OSError):
        pass
    return None, None


def cpu_baseline(layer, x, g, seconds, threads=None):
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
            "sample": f"first {sample} nodes of the same graph, faithful per-node loop, {dt:.1f} s"}


def build_inputs(a, rank, world):
    from ruvector_b200 import synth

    x = synth.make_features(a.nodes, a.dim, seed=0x5EED + rank)
    # rank r owns rows [r*nodes, (r+1)*nodes) of a (world*nodes)-node graph; columns are global
    g = synth.hnsw_like_graph(a.nodes, a.m, seed=0xBEEF + rank, locality=a.locality, window=a.window,
                              n_total=a.nodes * world, row_offset=rank * a.nodes)
    w = synth.make_weights(a.dim, a.dim)
    return x, g, w


def run_reference(a):
    """--impl reference: the reference's own CPU algorithm (restated in oracle/; the reference is Rust and
    cannot be built here) on this box's host cores, bounded sample per step.  NumPy inputs and the oracle
    only: libruvgnn.so is never loaded by this arm."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from tests.oracle_binding import HostLayer, Oracle

    nodes = min(a.nodes, 200_000)  # the sample only touches a prefix; keep generation cheap
    aa = argparse.Namespace(**{**vars(a), "nodes": nodes})
    x, g, w = build_inputs(aa, 0, 1)
    layer = HostLayer(a.dim, a.dim, a.heads, 0.1, w)
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
    loaded = [l.split()[-1] for l in open("/proc/self/maps") if "libruvgnn" in l or "liboracle" in l]
    line = {
        "impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": a.gpus, "steps": a.steps,
        "warmup": a.warmup, "ms_per_step": 1e3 * dt / a.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": workload_name(a, a.gpus), "sampled_nodes_per_step": sample,
                   "repo_libraries_loaded": sorted(set(os.path.basename(p) for p in loaded))},
        "cpu_baseline": {"value": val, "unit": UNIT, "cores": threads, "kind": "port",
                         "sample": f"{sample} nodes per step of the same synthetic graph (faithful per-node loop)"},
        "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "edges_per_sec": val * float(g.row_ptr[sample]) / sample,
    }
    print(json.dumps(line))


class Harness:
    """Process-wide plumbing: torch.distributed for barriers and max-over-ranks, CUDA events for timing."""

    def __init__(self):
        import torch
        import torch.distributed as dist

        self.torch, self.dist = torch, dist
        self.world = int(os.environ.get("WORLD_SIZE", "1"))
        self.rank = int(os.environ.get("RANK", "0"))
        self.local = int(os.environ.get("LOCAL_RANK", "0"))
        if not torch.cuda.is_available():
            raise SystemExit("bench.py needs a CUDA device (no CPU fallback)")
        torch.cuda.set_device(self.local)
        if self.world > 1:
            dist.init_process_group("nccl", device_id=torch.device("cuda", self.local))
        # a dedicated non-default stream, made current, so torch's events and the library's
        # kernels are on the same stream (handle 0 would make the library create its own)
        self.stream = torch.cuda.Stream()
        torch.cuda.set_stream(self.stream)

    def barrier(self):
        if self.world > 1:
            self.dist.barrier()
        self.torch.cuda.synchronize()

    def max_over_ranks(self, v: float) -> float:
        if self.world == 1:
            return v
        t = self.torch.tensor([v], device="cuda", dtype=self.torch.float64)
        self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        return float(t.item())

    def sum_over_ranks(self, v: float) -> float:
        if self.world == 1:
            return v
        t = self.torch.tensor([v], device="cuda", dtype=self.torch.float64)
        self.dist.all_reduce(t)
        return float(t.item())

    def timed(self, fn, steps, warmup):
        """W untimed steps, then exactly K steps between barrier + synchronize, device time, max over ranks."""
        for _ in range(warmup):
            fn()
        self.barrier()
        e0, e1 = self.torch.cuda.Event(enable_timing=True), self.torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            fn()
        e1.record()
        self.barrier()
        return self.max_over_ranks(float(e0.elapsed_time(e1))) / steps

    def mem_used_gb(self):
        free, total = self.torch.cuda.mem_get_info()
        return round((total - free) / 1e9, 2)


def halo_block(part, dist_, stats, overlap_mode):
    i = part.info
    return {"plane": "all-gather" if part.dense else ("peer-direct" if stats.get("peer_direct") else "all-to-all-v"),
            "chunks": i["chunks"],
            "rows_received_rank0": i["n_halo"] if not part.dense else i["n_total"] - i["n_local"],
            "rows_sent_rank0": i["n_send"] if not part.dense else i["n_local"] * (dist_.world - 1),
            "remote_edge_fraction_rank0": round(i["remote_edges"] / max(i["n_edges"], 1), 4),
            "bytes_received_rank0": stats["bytes_received"], "bytes_sent_rank0": stats["bytes_sent"],
            "exchange_ms_rank0": round(stats["exchange_ms"], 3), "exposed_wait_ms_rank0": round(stats["exposed_wait_ms"], 3),
            "gather_passes": stats["passes"],
            "mode": ("peer-direct (gather reads remote rows over NVLink)" if stats.get("peer_direct") else
                     "exact single-pass" if stats["exact"] else overlap_mode),
            "interior_rows_rank0": i["interior_rows"]}


def secondary_config3(h, a, hbm_peak, peak_src, tf32_peak, tf32_src):
    """Config 3 at size on one B200: 10M nodes x 128, 4 heads, three stacked layers with distinct weights, GNN
    layer l over HNSW level 2 - l (top-down, search.rs:116), graphs and features generated on the device."""
    import ruvector_b200 as rv
    from ruvector_b200 import _lib, synth, synth_cuda

    torch = h.torch
    n = a.secondary_nodes or 10_000_000
    d, heads, m = 128, 4, 32
    ctx = rv.Context(h.local, h.stream.cuda_stream)
    x = synth_cuda.make_features(n, d, seed=0x5EED)
    levels = [2, 1, 0]
    graphs, edges, active = [], [], []
    for lv in levels:
        rp, col, w = synth_cuda.hnsw_like_graph(n, m, seed=0xBEEF, level=lv)
        graphs.append(rv.DeviceGraph(ctx, rp, col, w))
        edges.append(int(col.numel()))
        active.append(int(((rp[1:] - rp[:-1]) > 0).sum().item()))
        del rp, col, w
    torch.cuda.empty_cache()
    layers = [rv.RuvectorLayer(d, d, heads, 0.1, weights=synth.make_weights(d, d, seed=0xC0FFEE + 17 * l), ctx=ctx)
              for l in range(3)]
    out = torch.empty(n, d, device="cuda", dtype=torch.float32)
    steps, warmup = max(2, min(a.steps, 5)), 3
    ms = h.timed(lambda: rv.forward_stack_device(layers, x.data_ptr(), out.data_ptr(), graphs), steps, warmup)
    for _ in range(2):
        rv.forward_stack_device(layers, x.data_ptr(), out.data_ptr(), graphs, timing=True)
    t = ctx.last_timing()  # averages over the 3 x 2 per-layer slots
    stage = np.array([t["project_ms"], t["aggregate_ms"], t["update_ms"]]) * 3.0
    mode = layers[0].mode
    passes = 3 if mode == _lib.MODE_FP32_TC else 1
    R = int(_lib.load().ruvgnn_layer_halo_width(layers[0]._device_layer())) // d
    E = float(sum(edges))
    # sparse levels (fewer than half of the rows active) run gather and update on the active rows only
    full = float(sum(a_ if 2 * a_ < n else n for a_ in active))
    msg_only = float(3 * n) - full
    roof = kagg_roofline(d, heads, R, E, full, float(stage[1]), hbm_peak, peak_src, stage)
    roof["note"] = "three K-agg launches per step (levels 2, 1, 0); bytes and time summed; sparse levels gather the active rows only"
    finite = bool(torch.isfinite(out[:: max(n // 4096, 1)]).all().item())
    blk = {"config": "3-layer stacked forward across HNSW levels, 10M nodes d=128, 1 B200 (BASELINE configs[2])",
           "workload": f"{n} nodes d={d} M={m}, {heads} heads, 3 stacked layers over levels {levels} (edges {edges})",
           "entry": "ruvgnn_stack_forward (device pointers)", "metric": METRIC, "unit": UNIT,
           "value": 3.0 * n / (ms * 1e-3), "ms_per_step": ms, "steps": steps, "warmup": warmup,
           "layer_node_updates_per_step": 3 * n, "edges_per_sec": E / (ms * 1e-3),
           "projection_mode": {0: "fp32_simt", 1: "fp32_3xtf32_tcgen05", 2: "tf32_tcgen05"}[mode],
           "roofline": roof,
           "rows_with_neighbours_per_level": active,
           "roofline_projections": gemm_roofline(d, d, heads, R, full, float(stage[0] + stage[2]), passes, hbm_peak,
                                                 tf32_peak, tf32_src, msg_only_rows=msg_only,
                                                 update_launches=layers[-1].update_launches),
           "hbm_used_gb": h.mem_used_gb(), "output_finite_sampled": finite}
    for o in layers:
        o.close()
    for gph in graphs:
        gph.close()
    del x, out
    ctx.close()
    torch.cuda.empty_cache()
    return blk


def secondary_partitioned(h, a, hbm_peak, peak_src, tf32_peak, tf32_src, n, d, heads, config, lams=(0.0, 0.9),
                          msg_halo=None, time_overlap=True):
    """A range-partitioned configuration generated on the device: n nodes x d x heads per GPU, M=32, one layer, halo
    exchange inside the library; one block per locality value."""
    import ruvector_b200 as rv
    from ruvector_b200 import _lib, dist as rvdist, synth, synth_cuda

    torch = h.torch
    m = 32
    world, rank = h.world, h.rank
    ctx = rv.Context(h.local, h.stream.cuda_stream)
    layer = rv.RuvectorLayer(d, d, heads, 0.1, weights=synth.make_weights(d, d), ctx=ctx)
    slab = 1 << 21
    n_pad = (n + slab - 1) // slab * slab  # feature slabs are addressed by global slab id
    x = synth_cuda.make_features(n, d, seed=0x5EED, row_offset=rank * n_pad)
    out = torch.empty(n, d, device="cuda", dtype=torch.float32)
    out2 = None
    mode = layer.mode
    passes = 3 if mode == _lib.MODE_FP32_TC else 1
    R = int(_lib.load().ruvgnn_layer_halo_width(layer._device_layer())) // d
    bounds = rvdist.even_bounds(n * world, world)
    blocks = {}
    for lam in lams:
        dist_ = rvdist.Dist.from_env(ctx)  # one communicator (and one set of exchange buffers) per graph
        dist_.set_chunks(a.chunks)
        rp, col, w = synth_cuda.hnsw_like_graph(n, m, seed=0xBEEF + rank, locality=lam, window=a.window,
                                                 n_total=n * world, row_offset=rank * n)
        E = int(col.numel())
        t0 = time.perf_counter()
        part = rvdist.Partition(dist_, bounds, rp, col, w, device_ptrs=True, plane=a.plane)
        torch.cuda.synchronize()
        plan_s = time.perf_counter() - t0
        del rp, col, w
        torch.cuda.empty_cache()
        steps, warmup = max(2, min(a.steps, 5)), 3
        ms = h.timed(lambda: part.forward(layer, x, out, overlap=a.overlap, peer_direct=a.peer_direct, msg_halo=msg_halo), steps, warmup)
        stats = dist_.last_stats()
        ctx.set_timing(True)
        for _ in range(2):
            part.forward(layer, x, out, overlap=a.overlap, peer_direct=a.peer_direct, msg_halo=msg_halo)
        t = ctx.last_timing()
        ctx.set_timing(False)
        stage = np.array([t["project_ms"], t["aggregate_ms"], t["update_ms"]])
        # the same partition in the overlapped multi-pass mode, timed, and checked against the default (exact) result
        ms_other, check = None, None
        if not a.overlap and time_overlap:
            ms_other = h.timed(lambda: part.forward(layer, x, out, overlap=True, msg_halo=msg_halo), 3, 1)
            stats_other = dist_.last_stats()
            if out2 is None:
                out2 = torch.empty_like(out)
            part.forward(layer, x, out2, overlap=True, asynchronous=False, msg_halo=msg_halo)
            part.forward(layer, x, out, asynchronous=False, msg_halo=msg_halo)
            sl = slice(0, n, max(n // 65536, 1))
            num = (out2[sl] - out[sl]).norm(dim=1)
            den = out[sl].norm(dim=1).clamp_min(1e-6)
            check = {"overlapped_vs_exact_max_row_rel": float((num / den).max().item()), "rows_compared": int(num.numel()),
                     "finite": bool(torch.isfinite(out[sl]).all().item())}
        E_total = h.sum_over_ranks(float(E))
        blk = {"workload": f"{n * world} nodes d={d} M={m}, {heads} heads, 1 layer, {n} nodes per GPU, locality {lam}",
               "metric": METRIC, "unit": UNIT, "value": n * world / (ms * 1e-3), "ms_per_step": ms, "steps": steps,
               "warmup": warmup, "edges_per_sec": E_total / (ms * 1e-3), "plan_build_s_rank0": round(plan_s, 3),
               "roofline": kagg_roofline(d, heads, R, float(E), float(n), float(stage[1]), hbm_peak, peak_src, stage),
               "roofline_projections": gemm_roofline(d, d, heads, R, float(n), float(stage[0] + stage[2]), passes,
                                                     hbm_peak, tf32_peak, tf32_src,
                                                     update_launches=layer.update_launches),
               "halo": halo_block(part, dist_, stats, "overlapped multi-pass"), "check": check,
               "overlap_mode": None if ms_other is None else {
                   "ms_per_step": ms_other, "value": n * world / (ms_other * 1e-3), "gather_passes": stats_other["passes"],
                   "exposed_wait_ms_rank0": round(stats_other["exposed_wait_ms"], 3),
                   "exchange_ms_rank0": round(stats_other["exchange_ms"], 3),
                   "note": "RUVGNN_DIST_OVERLAP: one gather pass per landed chunk; opt-in"},
               "hbm_used_gb_rank0": h.mem_used_gb()}
        if stats["passes"] > 1:
            blk["roofline"]["note"] = "K-agg time summed over the gather passes; bytes = the single-pass algorithmic bytes"
        blocks[f"locality_{lam}"] = blk
        part.close()
        dist_.close()
        torch.cuda.empty_cache()
    res = {"config": config, "entry": "ruvgnn_dist_forward (device pointers, NCCL inside the library)",
           "projection_mode": {0: "fp32_simt", 1: "fp32_3xtf32_tcgen05", 2: "tf32_tcgen05"}[mode], **blocks}
    layer.close()
    del x, out, out2
    ctx.close()
    torch.cuda.empty_cache()
    return res


def main():
    a = parse()
    if a.impl == "reference":
        return run_reference(a)

    h = Harness()
    torch, world, rank, local = h.torch, h.world, h.rank, h.local
    import ruvector_b200 as rv
    from ruvector_b200 import _lib

    n_gpus = world
    hbm_peak, peak_src, tf32_peak, tf32_src = peaks()
    line = None
    if not a.secondary_only:
        x, g, w = build_inputs(a, rank, world)
        ctx = rv.Context(local, h.stream.cuda_stream)
        layer = rv.RuvectorLayer(a.dim, a.dim, a.heads, 0.1, weights=w, ctx=ctx, mode=None if a.mode < 0 else a.mode)
        mode = layer.mode  # no fallback: an unsupported explicit mode has raised above
        n, H, E = a.nodes, a.dim, g.n_edges
        R = int(_lib.load().ruvgnn_layer_halo_width(layer._device_layer())) // H
        dist_ = part = None
        if world > 1:
            from ruvector_b200 import dist as rvdist

            dist_ = rvdist.Dist.from_env(ctx)
            dist_.set_chunks(a.chunks)
            part = rvdist.Partition.from_graph(dist_, g, a.nodes, plane=a.plane)
            step_fn = lambda xd, od: part.forward(layer, xd, od, overlap=a.overlap, msg_halo=False if a.row_halo else None, peer_direct=a.peer_direct)
            dg = None
        else:
            dg = g.to_device(ctx)
            step_fn = lambda xd, od: layer.forward_graph_device(xd.data_ptr(), od.data_ptr(), dg)

        x_pin = torch.from_numpy(x).pin_memory()
        x_dev = x_pin.cuda(non_blocking=True)
        out_dev = torch.empty(n, H, device="cuda", dtype=torch.float32)
        torch.cuda.synchronize()

        # ---- device-resident leg: `value` --------------------------------------------
        for _ in range(max(a.warmup, 3)):
            step_fn(x_dev, out_dev)
        h.barrier()
        sampler = ClockSampler(local)
        if rank == 0:
            sampler.start()
        launches0 = ctx.launch_count
        ms_per_step = h.timed(lambda: step_fn(x_dev, out_dev), a.steps, 0)
        launches = ctx.launch_count - launches0
        clocks = sampler.stop() if rank == 0 else None
        stats = dist_.last_stats() if world > 1 else None
        # per-stage device times (events on the launching stream), two extra timed steps
        if world > 1:
            ctx.set_timing(True)
            for _ in range(2):
                step_fn(x_dev, out_dev)
            t = ctx.last_timing()
            ctx.set_timing(False)
        else:
            for _ in range(2):
                layer.forward_graph_device(x_dev.data_ptr(), out_dev.data_ptr(), dg, timing=True)
            t = ctx.last_timing()
        stage = np.array([t["project_ms"], t["aggregate_ms"], t["update_ms"]])
        E_total = h.sum_over_ranks(float(E))
        value = n * n_gpus / (ms_per_step * 1e-3)

        # ---- e2e leg: host buffers through the C ABI (H2D + D2H inside the timed region) ----
        e2e = None
        if not a.no_e2e:
            out_pin = torch.empty(n, H, dtype=torch.float32).pin_memory()
            x_np, out_np = x_pin.numpy(), out_pin.numpy()
            if world > 1:
                e2e_fn = lambda: part.forward_host(layer, x_np, out_np, overlap=a.overlap, msg_halo=False if a.row_halo else None, peer_direct=a.peer_direct)
            else:
                e2e_fn = lambda: layer.forward_graph(x_np, dg, out=out_np)
            k2 = max(2, min(a.steps, 5))
            e_ms = h.timed(e2e_fn, k2, 1)
            e2e = {"value": n * n_gpus / (e_ms * 1e-3), "unit": UNIT, "h2d_bytes_per_step": int(x.nbytes),
                   "d2h_bytes_per_step": int(n * H * 4), "ms_per_step": e_ms, "steps": k2,
                   "entry": "ruvgnn_dist_forward (host pointers)" if world > 1 else "ruvgnn_layer_forward (host pointers)"}

        if rank == 0:
            passes = 3 if mode == _lib.MODE_FP32_TC else 1
            roof = kagg_roofline(H, a.heads, R, float(E), float(n), float(stage[1]) if stage[1] > 0 else None, hbm_peak,
                                 peak_src, stage)
            roof["traffic"], roof["traffic_capture"] = traffic_from_profile(n, H, a.heads)
            if world > 1 and stats and not stats["exact"]:
                roof["note"] = "K-agg time summed over the gather passes; bytes = the single-pass algorithmic bytes"
            line = {
                "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": n_gpus, "steps": a.steps, "warmup": max(a.warmup, 3),
                "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
                "dtype": "f32", "data": "synthetic",
                "config": {"workload": workload_name(a, n_gpus), "nodes_per_gpu": n, "dim": H, "heads": a.heads, "m": a.m,
                           "edges": E_total, "locality": a.locality, "window": a.window,
                           "l2": "inputs larger than L2 (X, projected rows >> 126 MB)",
                           "projection_mode": {0: "fp32_simt", 1: "fp32_3xtf32_tcgen05", 2: "tf32_tcgen05"}[mode],
                           "parallelism": f"range-partition x{n_gpus}, halo exchange in libruvgnn (NCCL)" if n_gpus > 1 else "single",
                           "env_knobs": active_knobs()},
                "edges_per_sec": E_total / (ms_per_step * 1e-3),
                "clocks": clocks, "e2e": e2e, "gpu_launches": int(launches), "roofline": roof,
                "roofline_projections": gemm_roofline(a.dim, H, a.heads, R, float(n), float(stage[0] + stage[2]), passes,
                                                      hbm_peak, tf32_peak, tf32_src,
                                                      update_launches=layer.update_launches),
            }
            if world > 1:
                line["halo"] = halo_block(part, dist_, stats, "overlapped multi-pass")
            if not a.no_cpu_baseline and world == 1:
                line["cpu_baseline"] = cpu_baseline(layer, x, g, a.cpu_seconds)
                # the reference itself is single-threaded (no rayon / SIMD use in crates/ruvector-gnn/src):
                # the same faithful loop on ONE host thread, short sample
                line["cpu_baseline_single_thread"] = cpu_baseline(layer, x, g, 4.0, threads=1)
        # release config 2 before the large secondary configurations
        if part is not None:
            part.close()
        if dist_ is not None:
            dist_.close()
        if dg is not None:
            dg.close()
        layer.close()
        del x_dev, out_dev, x_pin
        ctx.close()
        torch.cuda.empty_cache()

    if not a.no_secondary:
        # Safety net: the secondary configurations are large (up to ~150 GB of HBM per GPU).  If they hang
        # (a rank lost to an out-of-memory error inside a collective) the config-2 line must still come out.
        def _bail():
            if rank == 0 and line is not None:
                line["secondary"] = {"error": "secondary configurations timed out after 420 s"}
                print(json.dumps(line), flush=True)
            os._exit(0)

        watchdog = threading.Timer(420.0, _bail)
        watchdog.daemon = True
        watchdog.start()
        try:
            if world == 1:
                sec = secondary_config3(h, a, hbm_peak, peak_src, tf32_peak, tf32_src)
            else:
                # the per-GPU slice of config 5: 12.5M nodes x 128 x 4 heads per GPU (100M nodes at N=8)
                sec = {"config5_slice": secondary_partitioned(
                    h, a, hbm_peak, peak_src, tf32_peak, tf32_src, a.secondary_nodes or 12_500_000, 128, 4,
                    "100M nodes d=128, M=32 partitioned over 8xB200 (BASELINE configs[4]): per-GPU slice, weak scaling")}
                # config 4: 10M nodes x 768 x 8 heads over the N GPUs (strong: 10M / N per GPU).  At locality 0 every
                # GPU holds the K|V|M rows of all 10M nodes (92 GB) beside its own: that fits from 8 GPUs on.
                n4 = 10_000_000 // world
                if a.secondary_nodes == 0 and world >= 8:
                    sec["config4"] = secondary_partitioned(
                        h, a, hbm_peak, peak_src, tf32_peak, tf32_src, n4, 768, 8,
                        f"10M nodes d=768, 8 heads, M=32, sharded with halo exchange (BASELINE configs[3]): {n4} nodes per GPU; "
                        f"K|V|M rows on the wire (the message-row halo's staging buffer does not fit beside the 92 GB halo)",
                        msg_halo=False, time_overlap=False)
                elif a.secondary_nodes == 0:
                    sec["config4"] = {"skipped": f"10M x 768 at locality 0 needs the K|V|M rows of all 10M nodes per GPU "
                                                 f"(92 GB) + {n4} own rows x 12 KB: run at 8 GPUs"}
        except (rv.GnnError, RuntimeError, MemoryError) as e:  # reported, never silently dropped
            sec = {"error": f"{type(e).__name__}: {str(e)[:300]}"}
        watchdog.cancel()
        if rank == 0:
            if line is None:
                line = {"metric": METRIC, "n_gpus": n_gpus, "secondary_only": True}
            line["secondary"] = sec
    if rank == 0:
        print(json.dumps(line))
    if world > 1:
        h.dist.destroy_process_group()


if __name__ == "__main__":
    main()
