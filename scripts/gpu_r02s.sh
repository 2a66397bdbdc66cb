#!/bin/bash
# Round 2, session S (2 GPUs): the partitioned forward with the fused K-upd -- multi-GPU test tier (bit-identity with
# one GPU at d=128 included) and a short config-5-slice line.
set -u
N=${1:-2}
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_multi.py -m gpu -x -q > gpurun_out/s_pytest_gpu_n$N.log 2>&1
echo "pytest exit $?"; tail -4 gpurun_out/s_pytest_gpu_n$N.log | cut -c1-300
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511"
timeout 500 $TR bench.py --gpus $N --steps 3 --warmup 3 --secondary-only --secondary-nodes 2000000 --no-cpu-baseline --no-e2e > gpurun_out/s_n${N}_config5_slice_2M.json 2> gpurun_out/s_n${N}.err
echo "bench exit $?"; grep -iE "error|Traceback" gpurun_out/s_n${N}.err | head -3; cut -c1-600 gpurun_out/s_n${N}_config5_slice_2M.json
