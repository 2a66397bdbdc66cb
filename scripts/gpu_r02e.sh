#!/bin/bash
# Round 2, N-GPU contention experiments: the exchange slows down 2-3x while the gather passes saturate HBM.
set -u
N=${1:-2}
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511"
run() {  # tag, env..., -- args
  tag=$1; shift
  envs=()
  while [[ $1 != "--" ]]; do envs+=("$1"); shift; done
  shift
  env "${envs[@]}" timeout 600 $TR bench.py --gpus $N --steps 5 --warmup 3 --no-secondary --no-e2e "$@" > gpurun_out/e_n${N}_$tag.json 2> gpurun_out/e_n${N}_$tag.err
  echo "[$tag] exit $?"
}
run default X=1 --
run exact X=1 -- --exact
run msghalo X=1 -- --msg-halo
run chunks2 X=1 -- --chunks 2
run chunks2_nch32 NCCL_MIN_NCHANNELS=32 -- --chunks 2
run chunks2_nch64 NCCL_MIN_NCHANNELS=64 NCCL_MAX_NCHANNELS=64 -- --chunks 2
run chunks2_sms32 RUVGNN_COMM_SMS=32 -- --chunks 2
run chunks2_sms64 RUVGNN_COMM_SMS=64 -- --chunks 2
run chunks4_msghalo X=1 -- --chunks 4 --msg-halo
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/e_n*_*.json')):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1]); h=d['halo']
        print(f.split('/')[-1], "ms=%.2f"%d['ms_per_step'], {k:round(v,1) for k,v in d['roofline']['stage_ms'].items()}, 'chunks',h['chunks'],'passes',h['gather_passes'],'exch',h['exchange_ms_rank0'],'exposed',h['exposed_wait_ms_rank0'],'GB',round(h['bytes_received_rank0']/1e9,2))
    except Exception as e:
        print(f, 'ERR', e)
PY
