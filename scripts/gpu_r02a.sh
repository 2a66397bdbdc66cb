#!/bin/bash
# Round 2, first GPU session (1 GPU): tests, smoke, bench with the config-3 secondary, ncu passes.
set -u
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total,clocks.max.sm --format=csv > gpurun_out/gpu.txt 2>&1
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1
echo "pytest exit $?" | tee -a gpurun_out/pytest_gpu.log
tail -25 gpurun_out/pytest_gpu.log
timeout 300 python __graft_entry__.py --smoke > gpurun_out/smoke.log 2>&1; echo "smoke exit $?"; tail -6 gpurun_out/smoke.log
timeout 900 python bench.py --steps 5 --warmup 3 --cpu-seconds 5 > gpurun_out/bench.json 2> gpurun_out/bench.err
echo "bench exit $?"; cat gpurun_out/bench.json; tail -5 gpurun_out/bench.err
# d=128 shape (north-star layer shape), 2M nodes: launch list + full counters of one step's kernels
CMD="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e --no-secondary --nodes 2000000 --dim 128 --heads 4"
timeout 600 $CMD > gpurun_out/bench_h128_2M.json 2> gpurun_out/plain.log && cat gpurun_out/bench_h128_2M.json &&
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv \
    --log-file gpurun_out/launches_h128.csv $CMD > gpurun_out/ncu_launches.log 2>&1
echo "ncu launches exit $?"
timeout 1500 ncu --set full --clock-control none --import-source on -s 24 -c 8 \
    -o gpurun_out/h128_step -f $CMD > gpurun_out/ncu_full_h128.log 2>&1
echo "ncu full h128 exit $?"; tail -3 gpurun_out/ncu_full_h128.log
# config 2: K-agg traffic (one launch) for roofline.traffic
CMD2="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e --no-secondary"
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:gather_fast -s 3 -c 1 \
    -o gpurun_out/kagg_c2 -f $CMD2 > gpurun_out/ncu_full_c2.log 2>&1
echo "ncu full c2 exit $?"; tail -3 gpurun_out/ncu_full_c2.log
ls -la gpurun_out
