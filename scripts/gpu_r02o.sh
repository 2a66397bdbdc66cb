#!/bin/bash
# Round 2, session O: fused K-upd on by default -- the whole GPU test tier, smoke, the phase timeline and the
# d=128 A/B again.
set -u
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/o_pytest_gpu.log 2>&1
echo "pytest exit $?"; tail -8 gpurun_out/o_pytest_gpu.log | cut -c1-300
timeout 300 python __graft_entry__.py --smoke > gpurun_out/o_smoke.log 2>&1; echo "smoke exit $?"; tail -4 gpurun_out/o_smoke.log
B="python bench.py --nodes 1000000 --dim 128 --heads 4 --steps 5 --warmup 3 --no-cpu-baseline --no-e2e --no-secondary"
RUVGNN_FUSED_DEBUG=16 timeout 300 $B > gpurun_out/o_trace.json 2> gpurun_out/o_trace.err; grep "upd_fused trace" gpurun_out/o_trace.err | tail -1
RUVGNN_FUSED_UPD=0 timeout 600 $B > gpurun_out/o_bench_h128_1M_unfused.json 2> gpurun_out/o_unfused.err
echo "unfused exit $?"; python scripts/print_bench.py gpurun_out/o_bench_h128_1M_unfused.json 2>/dev/null | head -20
timeout 600 $B > gpurun_out/o_bench_h128_1M_fused.json 2> gpurun_out/o_fused.err
echo "fused exit $?"; python scripts/print_bench.py gpurun_out/o_bench_h128_1M_fused.json 2>/dev/null | head -20
