#!/bin/bash
# Run on the GPU box (via gpurun): parity tests, a short bench, then the ncu passes.
# Usage: scripts/gpu_check.sh [tests|bench|ncu|all]
set -u
what=${1:-all}
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total,clocks.max.sm --format=csv > gpurun_out/gpu.txt 2>&1
if [[ $what == tests || $what == all ]]; then
  timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1
  echo "pytest exit $?" | tee -a gpurun_out/pytest_gpu.log
  tail -15 gpurun_out/pytest_gpu.log
fi
if [[ $what == bench || $what == all ]]; then
  timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err
  echo "bench exit $?"; cat gpurun_out/bench.json; tail -5 gpurun_out/bench.err
fi
if [[ $what == ncu || $what == all ]]; then
  CMD="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e"
  timeout 600 $CMD > gpurun_out/plain.log 2>&1 &&
  timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv \
      --log-file gpurun_out/launches.csv $CMD > gpurun_out/ncu_launches.log 2>&1
  echo "ncu launches exit $?"
  timeout 600 $CMD > gpurun_out/plain2.log 2>&1 &&
  timeout 1200 ncu --set full --clock-control none --import-source on -k regex:gather_fast -s 3 -c 1 \
      -o gpurun_out/kagg -f $CMD > gpurun_out/ncu_full.log 2>&1
  echo "ncu full exit $?"; tail -3 gpurun_out/ncu_full.log
fi
