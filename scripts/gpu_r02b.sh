#!/bin/bash
# Round 2, 2-GPU session: in-library dist (NCCL) parity + C caller + bench at N=2 with the config-5 slice.
set -u
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total --format=csv > gpurun_out/gpu2.txt 2>&1
timeout 300 python __graft_entry__.py --smoke > gpurun_out/smoke2.log 2>&1; echo "smoke exit $?"; tail -5 gpurun_out/smoke2.log
timeout 900 python -m pytest tests/test_gpu_multi.py tests/test_cpp_host.py tests/test_gpu_multipass.py -m gpu -x -q -s > gpurun_out/pytest_gpu_n2.log 2>&1
echo "pytest exit $?" | tee -a gpurun_out/pytest_gpu_n2.log
grep -E "MULTI_GPU|plane |n=|passed|failed|Error|error" gpurun_out/pytest_gpu_n2.log | tail -30
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511"
timeout 900 $TR bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/bench_n2.json 2> gpurun_out/bench_n2.err
echo "bench n2 exit $?"; cat gpurun_out/bench_n2.json; tail -5 gpurun_out/bench_n2.err
timeout 600 $TR bench.py --gpus 2 --steps 5 --warmup 3 --exact --no-secondary --no-e2e > gpurun_out/bench_n2_exact.json 2> gpurun_out/bench_n2_exact.err
echo "bench n2 exact exit $?"; cat gpurun_out/bench_n2_exact.json; tail -3 gpurun_out/bench_n2_exact.err
timeout 600 $TR bench.py --gpus 2 --steps 5 --warmup 3 --locality 0.9 --no-secondary --no-e2e > gpurun_out/bench_n2_l09.json 2> gpurun_out/bench_n2_l09.err
echo "bench n2 l0.9 exit $?"; cat gpurun_out/bench_n2_l09.json; tail -3 gpurun_out/bench_n2_l09.err
