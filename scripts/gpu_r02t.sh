#!/bin/bash
# Round 2, session T: last GPU seconds -- the d=128 line with the final fused K-upd, then the parity files that run it.
set -u
mkdir -p gpurun_out
B="python bench.py --nodes 1000000 --dim 128 --heads 4 --steps 5 --warmup 3 --no-cpu-baseline --no-e2e --no-secondary"
timeout 60 $B > gpurun_out/t_bench_h128_1M.json 2> gpurun_out/t_bench.err; python scripts/print_bench.py gpurun_out/t_bench_h128_1M.json 2>/dev/null | cut -c1-300
timeout 60 python -m pytest tests/test_gpu_tc.py tests/test_gpu_parity.py -m gpu -x -q > gpurun_out/t_pytest.log 2>&1
echo "pytest exit $?"; tail -3 gpurun_out/t_pytest.log | cut -c1-200
