#!/bin/bash
# Round 2: the default bench line at N GPUs on the final code (config 2 per GPU + secondary), one command.
set -u
N=${1:-4}
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511"
timeout 900 $TR bench.py --gpus $N --steps 5 --warmup 3 > gpurun_out/k_n${N}_bench.json 2> gpurun_out/k_n${N}_bench.err
echo "bench exit $?"; grep -v "OMP_NUM\|\*\*\*\*" gpurun_out/k_n${N}_bench.err | tail -5
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/k_n*_bench.json')):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1]); h=d['halo']
        print(f.split('/')[-1], "ms=%.2f value=%.3e"%(d['ms_per_step'],d['value']), 'e2e', (d.get('e2e') or {}).get('ms_per_step'), {k:round(v,1) for k,v in d['roofline']['stage_ms'].items()}, h['plane'], h['mode'][:22], 'chunks',h['chunks'],'exch',h['exchange_ms_rank0'],'exposed',h['exposed_wait_ms_rank0'],'GB',round(h['bytes_received_rank0']/1e9,2))
        for k,v in (d.get('secondary') or {}).items():
            if isinstance(v,dict):
                for kk,b in v.items():
                    if isinstance(b,dict) and 'ms_per_step' in b:
                        print('  SEC',k,kk,"ms=%.2f value=%.3e"%(b['ms_per_step'],b['value']), {a:round(c,1) for a,c in b['roofline']['stage_ms'].items()}, 'kagg',round(b['roofline']['frac'],3), b['halo']['plane'], 'exposed', b['halo']['exposed_wait_ms_rank0'], 'overlap', (b.get('overlap_mode') or {}).get('ms_per_step'), 'mem', b.get('hbm_used_gb_rank0'), 'check', (b.get('check') or {}).get('overlapped_vs_exact_max_row_rel'))
                    elif not isinstance(b,dict): print('  SEC',k,kk,str(b)[:200])
            else: print('  SEC',k,str(v)[:300])
    except Exception as e:
        print(f,'ERR',e)
PY
