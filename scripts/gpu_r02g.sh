#!/bin/bash
# Round 2, 8-GPU session G: the driver-like bench line (config 2 + config-5 slice + config 4), full-size sampled
# parity at 100M nodes, and the plane / mode variants that decide the defaults.
set -u
N=${1:-8}
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511"
timeout 1200 $TR bench.py --gpus $N --steps 5 --warmup 3 > gpurun_out/g_n${N}_bench.json 2> gpurun_out/g_n${N}_bench.err
echo "bench exit $?"; grep -v "OMP_NUM\|\*\*\*\*" gpurun_out/g_n${N}_bench.err | tail -5
timeout 900 $TR scripts/multi_gpu_parity_full.py > gpurun_out/g_n${N}_parity_full.log 2>&1
echo "parity full exit $?"; grep -E "config-5|FULL_SIZE|Error|error" gpurun_out/g_n${N}_parity_full.log | tail -5
run() { tag=$1; shift; timeout 600 $TR bench.py --gpus $N --steps 5 --warmup 3 --no-secondary --no-e2e "$@" > gpurun_out/g_n${N}_$tag.json 2> gpurun_out/g_n${N}_$tag.err; echo "[$tag] exit $?"; }
run msghalo --msg-halo
run msghalo_exact --msg-halo --exact
run l09_default --locality 0.9
run l09_direct --locality 0.9 --peer-direct
run l09_sparse_exact --locality 0.9 --exact
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/g_n*_*.json')):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        if 'halo' in d:
            h=d['halo']
            print(f.split('/')[-1], "ms=%.2f"%d['ms_per_step'], 'e2e', (d.get('e2e') or {}).get('ms_per_step'), {k:round(v,1) for k,v in d['roofline']['stage_ms'].items()}, h['plane'], h['mode'][:24], 'passes',h['gather_passes'],'exch',h['exchange_ms_rank0'],'exposed',h['exposed_wait_ms_rank0'],'GB',round(h['bytes_received_rank0']/1e9,2), 'remote', h['remote_edge_fraction_rank0'])
        for k,v in (d.get('secondary') or {}).items():
            if isinstance(v,dict):
                for kk,b in v.items():
                    if isinstance(b,dict) and 'ms_per_step' in b:
                        print('  SEC',k,kk,"ms=%.2f value=%.3e"%(b['ms_per_step'],b['value']), {a:round(c,1) for a,c in b['roofline']['stage_ms'].items()}, b['halo']['plane'], b['halo']['mode'][:20], 'exact', (b.get('exact_mode') or {}).get('ms_per_step'), 'mem', b.get('hbm_used_gb_rank0'), 'check', (b.get('check') or {}).get('overlapped_vs_exact_max_row_rel'))
                    elif not isinstance(b,dict): print('  SEC',k,kk,str(b)[:160])
            else: print('  SEC',k,str(v)[:200])
    except Exception as e:
        print(f, 'ERR', e)
PY
