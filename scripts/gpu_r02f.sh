#!/bin/bash
# Round 2: peer-direct plane on N GPUs (parity + bench at locality 0.9, forced sparse at small N).
set -u
N=${1:-2}
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_multi.py tests/test_cpp_host.py -m gpu -x -q -s > gpurun_out/pytest_gpu_f.log 2>&1
echo "pytest exit $?" | tee -a gpurun_out/pytest_gpu_f.log
grep -E "MULTI_GPU|plane |^n=|passed|failed|Error|error|assert" gpurun_out/pytest_gpu_f.log | tail -20 | cut -c1-400
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511"
run() { tag=$1; shift; timeout 600 $TR bench.py --gpus $N --steps 5 --warmup 3 --no-secondary "$@" > gpurun_out/f_n${N}_$tag.json 2> gpurun_out/f_n${N}_$tag.err; echo "[$tag] exit $?"; tail -2 gpurun_out/f_n${N}_$tag.err | grep -v OMP; }
run l09_direct --locality 0.9 --plane sparse
run l09_nccl_sparse --locality 0.9 --plane sparse --no-peer-direct --no-e2e
run l09_nccl_sparse_exact --locality 0.9 --plane sparse --no-peer-direct --exact --no-e2e
run l09_dense_exact --locality 0.9 --plane dense --exact --no-e2e
run l0_direct --plane sparse --no-e2e
timeout 900 $TR bench.py --gpus $N --steps 5 --warmup 3 --secondary-only --plane sparse > gpurun_out/f_n${N}_sec_sparse.json 2> gpurun_out/f_n${N}_sec_sparse.err
echo "secondary sparse exit $?"
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/f_n*_*.json')):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        if 'halo' in d:
            h=d['halo']
            print(f.split('/')[-1], "ms=%.2f"%d['ms_per_step'], 'e2e', (d.get('e2e') or {}).get('ms_per_step'), {k:round(v,1) for k,v in d['roofline']['stage_ms'].items()}, h['plane'], h['mode'][:24], 'passes',h['gather_passes'],'exch',h['exchange_ms_rank0'],'exposed',h['exposed_wait_ms_rank0'],'GB',round(h['bytes_received_rank0']/1e9,2), 'remote', h['remote_edge_fraction_rank0'])
        for k,v in (d.get('secondary') or {}).items():
            if isinstance(v,dict):
                for kk,b in v.items():
                    if isinstance(b,dict) and 'ms_per_step' in b:
                        print('  SEC',k,kk,"ms=%.2f"%b['ms_per_step'], {a:round(c,1) for a,c in b['roofline']['stage_ms'].items()}, b['halo']['plane'], b['halo']['mode'][:20], 'exact', (b.get('exact_mode') or {}).get('ms_per_step'), 'check', b.get('check'))
    except Exception as e:
        print(f, 'ERR', e)
PY
