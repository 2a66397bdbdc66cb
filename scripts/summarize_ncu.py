#!/usr/bin/env python
"""Summarises ncu outputs brought back in gpurun_out/ into tracked files under profiles/.
  summarize_ncu.py launches <launches.csv> <out.md>       per-kernel time shares
  summarize_ncu.py full <report.ncu-rep> <out.csv> [json] key counters of a --set full capture
"""
import collections
import csv
import json
import re
import subprocess
import sys

KEYS = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__bytes_read.sum.per_second",
    "dram__throughput.avg.pct_of_peak_sustained_elapsed", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__grid_size",
    "launch__block_size", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
    "smsp__inst_executed.sum", "sm__cycles_elapsed.avg.per_second",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_tensor.sum", "sm__pipe_tensor_subpipe_tf32_cycles_active.avg.pct_of_peak_sustained_active",
    "smsp__average_warp_latency_issue_stalled_long_scoreboard.pct", "launch__shared_mem_per_block_dynamic",
]


def launches(src, dst):
    rows = [r for r in csv.reader(open(src)) if len(r) > 5]
    h = next(i for i, r in enumerate(rows) if r[0] == "ID")
    H, data = rows[h], rows[h + 1:]
    ki, vi = H.index("Kernel Name"), H.index("Metric Value")
    agg = collections.OrderedDict()
    for r in data:
        name = re.sub(r"\(.*", "", r[ki])
        name = re.sub(r"^void ", "", name).replace("ruvgnn::(anonymous namespace)::", "").replace("unnamed>::", "")
        agg.setdefault(name, []).append(float(r[vi].replace(",", "")))
    tot = sum(sum(v) for v in agg.values())
    with open(dst, "w") as f:
        f.write("| kernel | launches | avg ms | total ms | share |\n|---|---:|---:|---:|---:|\n")
        for k, v in sorted(agg.items(), key=lambda kv: -sum(kv[1])):
            f.write(f"| `{k}` | {len(v)} | {sum(v) / len(v) / 1e6:.3f} | {sum(v) / 1e6:.3f} | {100 * sum(v) / tot:.1f}% |\n")
        f.write(f"\nncu --metrics gpu__time_duration.sum --clock-control none; {len(data)} launches, "
                f"{tot / 1e6:.1f} ms total (cold-cache, serialised: compare shares, not absolutes)\n")


def full(rep, dst, js=None):
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    H, U = rows[0], rows[1]
    with open(dst, "w") as f:
        w = csv.writer(f)
        w.writerow(["kernel", "metric", "unit", "value"])
        for V in rows[2:]:
            name = re.sub(r"\(.*", "", V[H.index("Kernel Name")])
            vals = {}
            for h, u, v in zip(H, U, V):
                if h in KEYS:
                    w.writerow([name, h, u, v])
                    vals[h] = (v, u)
            if js and json.loads(js).get("kernel", "").split("<")[0] in name:
                rd, wr = vals.get("dram__bytes_read.sum"), vals.get("dram__bytes_write.sum")
                scale = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1, "Tbyte": 1e12}
                tot = float(rd[0].replace(",", "")) * scale[rd[1]] + float(wr[0].replace(",", "")) * scale[wr[1]]
                base = json.loads(js)
                base["dram_bytes_per_launch"] = tot
                print(json.dumps(base))
                js = None


if __name__ == "__main__":
    if sys.argv[1] == "launches":
        launches(sys.argv[2], sys.argv[3])
    else:
        full(sys.argv[2], sys.argv[3], sys.argv[4] if len(sys.argv) > 4 else None)
