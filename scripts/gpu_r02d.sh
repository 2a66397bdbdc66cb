#!/bin/bash
# Round 2, 2-GPU check of the pass-ordered multi-pass gather + message-row halo, then 1-GPU items:
# sparse levels (config 3), store streaming, L2 prefetch A/B at d=128.
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_multi.py tests/test_cpp_host.py tests/test_gpu_multipass.py tests/test_store.py tests/test_gpu_fullsize.py -m gpu -x -q -s > gpurun_out/pytest_gpu_d.log 2>&1
echo "pytest exit $?" | tee -a gpurun_out/pytest_gpu_d.log
grep -E "MULTI_GPU|plane |n=|passed|failed|Error|error|assert" gpurun_out/pytest_gpu_d.log | tail -30
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511"
for tag in "" "--msg-halo" "--exact" "--chunks 2"; do
  timeout 600 $TR bench.py --gpus 2 --steps 5 --warmup 3 --no-secondary --no-e2e $tag > gpurun_out/bench_n2_d.json 2> gpurun_out/bench_n2_d.err
  echo "bench n2 [$tag] exit $?"; cat gpurun_out/bench_n2_d.json
done
timeout 900 $TR bench.py --gpus 2 --steps 5 --warmup 3 --secondary-only > gpurun_out/bench_n2_sec.json 2> gpurun_out/bench_n2_sec.err
echo "bench n2 secondary exit $?"; cat gpurun_out/bench_n2_sec.json; tail -3 gpurun_out/bench_n2_sec.err
# single GPU: config 3 with sparse levels; prefetch A/B at d=128
timeout 600 python bench.py --steps 5 --warmup 3 --secondary-only > gpurun_out/bench_c3.json 2> gpurun_out/bench_c3.err
echo "bench c3 exit $?"; cat gpurun_out/bench_c3.json; tail -3 gpurun_out/bench_c3.err
for pf in 0 2 4 8; do
  RUVGNN_TC_PF=$pf timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-e2e --no-secondary --nodes 2000000 --dim 128 --heads 4 > gpurun_out/bench_h128_pf$pf.json 2>/dev/null
  echo "pf=$pf"; cat gpurun_out/bench_h128_pf$pf.json
done
for pf in 0 4; do
  RUVGNN_TC_PF=$pf timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-e2e --no-secondary > gpurun_out/bench_c2_pf$pf.json 2>/dev/null
  echo "c2 pf=$pf"; cat gpurun_out/bench_c2_pf$pf.json
done
