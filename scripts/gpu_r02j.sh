#!/bin/bash
# Round 2, final 1-GPU validation on the frozen code: GPU test tier, smoke, the default bench line, ncu launch list
# of the same command and the K-agg traffic capture that bench.py's roofline.traffic reads.
set -u
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -m gpu -x -q > gpurun_out/j_pytest_gpu.log 2>&1
echo "pytest exit $?" | tee -a gpurun_out/j_pytest_gpu.log
tail -6 gpurun_out/j_pytest_gpu.log | cut -c1-300
timeout 300 python __graft_entry__.py --smoke > gpurun_out/j_smoke.log 2>&1; echo "smoke exit $?"; tail -4 gpurun_out/j_smoke.log
timeout 900 python bench.py > gpurun_out/j_bench.json 2> gpurun_out/j_bench.err
echo "bench exit $?"; tail -3 gpurun_out/j_bench.err; cat gpurun_out/j_bench.json | cut -c1-1500
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/j_bench_reference.json 2> gpurun_out/j_bench_reference.err
echo "reference arm exit $?"; cat gpurun_out/j_bench_reference.json | cut -c1-600
CMD="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e --no-secondary"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/j_launches.csv $CMD > gpurun_out/j_ncu_launches.log 2>&1
echo "ncu launches exit $?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:gather_fast -s 3 -c 1 -o gpurun_out/j_kagg_c2 -f $CMD > gpurun_out/j_ncu_full.log 2>&1
echo "ncu full exit $?"; tail -2 gpurun_out/j_ncu_full.log
