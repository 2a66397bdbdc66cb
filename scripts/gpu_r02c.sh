#!/bin/bash
# Round 2, 8-GPU session: bench at N=8 (config 2 per GPU + the 100M-node config-5 slice), exact-mode line.
set -u
mkdir -p gpurun_out
N=${1:-8}
nvidia-smi --query-gpu=name,memory.total --format=csv > gpurun_out/gpu$N.txt 2>&1
nvidia-smi topo -m > gpurun_out/topo$N.txt 2>&1
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511"
timeout 1200 $TR bench.py --gpus $N --steps 5 --warmup 3 > gpurun_out/bench_n$N.json 2> gpurun_out/bench_n$N.err
echo "bench n$N exit $?"; cat gpurun_out/bench_n$N.json; grep -v "OMP_NUM\|\*\*\*\*" gpurun_out/bench_n$N.err | tail -8
timeout 600 $TR bench.py --gpus $N --steps 5 --warmup 3 --exact --no-secondary --no-e2e > gpurun_out/bench_n${N}_exact.json 2> gpurun_out/bench_n${N}_exact.err
echo "bench n$N exact exit $?"; cat gpurun_out/bench_n${N}_exact.json
timeout 600 $TR bench.py --gpus $N --steps 5 --warmup 3 --locality 0.9 --no-secondary --no-e2e > gpurun_out/bench_n${N}_l09.json 2> gpurun_out/bench_n${N}_l09.err
echo "bench n$N l0.9 exit $?"; cat gpurun_out/bench_n${N}_l09.json
timeout 600 $TR bench.py --gpus $N --steps 5 --warmup 3 --chunks 2 --no-secondary --no-e2e > gpurun_out/bench_n${N}_c2.json 2> gpurun_out/bench_n${N}_c2.err
echo "bench n$N chunks2 exit $?"; cat gpurun_out/bench_n${N}_c2.json
