#!/bin/bash
# Round 2, N-GPU check of the final dist defaults (exact single-pass, message-row halo rebuilt chunk by chunk).
set -u
N=${1:-2}
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_multi.py tests/test_cpp_host.py -m gpu -x -q -s > gpurun_out/i_pytest_gpu.log 2>&1
echo "pytest exit $?" | tee -a gpurun_out/i_pytest_gpu.log
grep -E "MULTI_GPU|plane |^n=|passed|failed|Error|error|assert" gpurun_out/i_pytest_gpu.log | tail -14 | cut -c1-420
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511"
run() { tag=$1; shift; timeout 600 $TR bench.py --gpus $N --steps 5 --warmup 3 --no-secondary "$@" > gpurun_out/i_n${N}_$tag.json 2> gpurun_out/i_n${N}_$tag.err; echo "[$tag] exit $?"; grep -iE "error|Traceback" gpurun_out/i_n${N}_$tag.err | head -3; }
run default
run overlap --overlap --no-e2e
run rowhalo --row-halo --no-e2e
run l09 --locality 0.9 --no-e2e
run chunks4 --chunks 4 --no-e2e
run chunks1 --chunks 1 --no-e2e
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/i_n*_*.json')):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1]); h=d['halo']
        print(f.split('/')[-1], "ms=%.2f"%d['ms_per_step'], 'e2e', (d.get('e2e') or {}).get('ms_per_step'), {k:round(v,1) for k,v in d['roofline']['stage_ms'].items()}, h['plane'], h['mode'][:22], 'chunks',h['chunks'],'passes',h['gather_passes'],'exch',h['exchange_ms_rank0'],'exposed',h['exposed_wait_ms_rank0'],'GB',round(h['bytes_received_rank0']/1e9,2))
    except Exception as e:
        print(f,'ERR',e)
PY
