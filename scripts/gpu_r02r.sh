#!/bin/bash
# Round 2, session R: validation of the round's final code on one B200 -- GPU test tier, smoke, the default bench
# line, the reference arm, the d=128 line, the ncu launch list of the d=128 command and one --set full capture of
# the fused K-upd kernel.
set -u
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r_pytest_gpu.log 2>&1
echo "pytest exit $?"; tail -5 gpurun_out/r_pytest_gpu.log | cut -c1-300
timeout 300 python __graft_entry__.py --smoke > gpurun_out/r_smoke.log 2>&1; echo "smoke exit $?"; tail -4 gpurun_out/r_smoke.log
timeout 900 python bench.py > gpurun_out/r_bench.json 2> gpurun_out/r_bench.err
echo "bench exit $?"; tail -3 gpurun_out/r_bench.err; python scripts/print_bench.py gpurun_out/r_bench.json | cut -c1-400
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r_bench_reference.json 2> gpurun_out/r_bench_reference.err
echo "reference arm exit $?"; cut -c1-300 gpurun_out/r_bench_reference.json
B="python bench.py --nodes 2000000 --dim 128 --heads 4 --steps 5 --warmup 3 --no-cpu-baseline --no-secondary"
timeout 600 $B > gpurun_out/r_bench_h128_2M.json 2> gpurun_out/r_h128.err
echo "d=128 exit $?"; python scripts/print_bench.py gpurun_out/r_bench_h128_2M.json | cut -c1-400
CMD="python bench.py --nodes 2000000 --dim 128 --heads 4 --steps 2 --warmup 3 --no-cpu-baseline --no-e2e --no-secondary"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/r_launches_h128_2M.csv $CMD > gpurun_out/r_ncu_launches.log 2>&1
echo "ncu launches exit $?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:upd_fused -s 2 -c 1 -o gpurun_out/r_upd_fused_h128 -f $CMD > gpurun_out/r_ncu_full.log 2>&1
echo "ncu full exit $?"; tail -2 gpurun_out/r_ncu_full.log
