#!/bin/bash
# Round 2, session L: the fused K-upd kernel (gemm_upd_fused.cu) -- parity tests, then an A/B of the
# d=128 layer shape (1M nodes) with the knob off / on.
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_fused.py -x -q > gpurun_out/l_pytest_fused.log 2>&1
rc=$?
echo "pytest exit $rc"; tail -25 gpurun_out/l_pytest_fused.log | cut -c1-400
if [ $rc -ne 0 ]; then exit 0; fi
B="python bench.py --nodes 1000000 --dim 128 --heads 4 --steps 5 --warmup 3 --no-cpu-baseline --no-e2e --no-secondary"
RUVGNN_FUSED_UPD=0 timeout 600 $B > gpurun_out/l_bench_h128_1M_unfused.json 2> gpurun_out/l_unfused.err
echo "unfused exit $?"; python scripts/print_bench.py gpurun_out/l_bench_h128_1M_unfused.json 2>/dev/null | head -20
RUVGNN_FUSED_UPD=1 timeout 600 $B > gpurun_out/l_bench_h128_1M_fused.json 2> gpurun_out/l_fused.err
echo "fused exit $?"; python scripts/print_bench.py gpurun_out/l_bench_h128_1M_fused.json 2>/dev/null | head -20
