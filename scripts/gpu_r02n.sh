#!/bin/bash
# Round 2, session N: phase timeline of the fused K-upd (RUVGNN_FUSED_DEBUG=16 and combinations).
set -u
mkdir -p gpurun_out
B="python bench.py --nodes 300000 --dim 128 --heads 4 --steps 2 --warmup 3 --no-cpu-baseline --no-e2e --no-secondary"
for dbg in ${DBGS:-24 27 17 18}; do
  RUVGNN_FUSED_DEBUG=$dbg timeout 300 $B > gpurun_out/n_dbg$dbg.json 2> gpurun_out/n_dbg$dbg.err
  echo "debug $dbg exit $?"; grep "upd_fused trace" gpurun_out/n_dbg$dbg.err | tail -1
done
