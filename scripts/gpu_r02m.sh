#!/bin/bash
# Round 2, session M: where the fused K-upd's time goes -- timing experiments (RUVGNN_FUSED_DEBUG masks) and an
# ncu --set full capture of one launch (1M nodes x 128, 4 heads).
set -u
mkdir -p gpurun_out
B="python bench.py --nodes 1000000 --dim 128 --heads 4 --steps 5 --warmup 3 --no-cpu-baseline --no-e2e --no-secondary"
for dbg in 0 1 2 4 8 3 12 15; do
  RUVGNN_FUSED_UPD=1 RUVGNN_FUSED_DEBUG=$dbg timeout 300 $B > gpurun_out/m_dbg$dbg.json 2> gpurun_out/m_dbg$dbg.err
  echo "debug $dbg exit $?"; python scripts/print_bench.py gpurun_out/m_dbg$dbg.json 2>/dev/null | cut -c1-300
done
RUVGNN_FUSED_UPD=1 timeout 600 ncu --set full --clock-control none --import-source on -k regex:upd_fused -s 2 -c 1 -o gpurun_out/m_upd_fused -f $B > gpurun_out/m_ncu_full.log 2>&1
echo "ncu full exit $?"; tail -2 gpurun_out/m_ncu_full.log
