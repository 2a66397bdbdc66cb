#!/bin/bash
# Round 2, 1-GPU validation: full GPU test tier, smoke, bench (default = row-block pipeline on) and the pipeline A/B.
set -u
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -m gpu -x -q > gpurun_out/h_pytest_gpu.log 2>&1
echo "pytest exit $?" | tee -a gpurun_out/h_pytest_gpu.log
tail -12 gpurun_out/h_pytest_gpu.log | cut -c1-300
timeout 300 python __graft_entry__.py --smoke > gpurun_out/h_smoke.log 2>&1; echo "smoke exit $?"; tail -4 gpurun_out/h_smoke.log
for sms in 48 0 32 64; do
  RUVGNN_UPD_SMS=$sms timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-e2e --no-secondary > gpurun_out/h_c2_sms$sms.json 2>/dev/null
  echo "c2 sms=$sms exit $?"
  RUVGNN_UPD_SMS=$sms timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-e2e --no-secondary --nodes 4000000 --dim 128 --heads 4 > gpurun_out/h_h128_sms$sms.json 2>/dev/null
  echo "h128 sms=$sms exit $?"
done
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/h_bench.json 2> gpurun_out/h_bench.err
echo "bench exit $?"; tail -3 gpurun_out/h_bench.err
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/h_*.json')):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f.split('/')[-1], "ms=%.2f value=%.3e"%(d['ms_per_step'],d['value']), 'e2e', (d.get('e2e') or {}).get('ms_per_step'), {k:round(v,2) for k,v in d['roofline']['stage_ms'].items()}, 'kagg', round(d['roofline']['frac'],3), 'gemm', round(d['roofline_projections']['frac'],3), (d['roofline_projections'].get('pipeline') or {}).get('blocks'))
        s=d.get('secondary')
        if s: print('  SEC ms=%.2f value=%.3e'%(s['ms_per_step'],s['value']), s['roofline']['stage_ms'], s.get('hbm_used_gb'))
    except Exception as e:
        print(f,'ERR',e)
PY
