B="python bench.py --nodes 1000000 --dim 128 --heads 4 --steps 5 --warmup 3 --no-cpu-baseline --no-e2e --no-secondary"
RUVGNN_FUSED_DEBUG=16 timeout 300 $B > gpurun_out/q_trace.json 2> gpurun_out/q_trace.err; grep "upd_fused trace" gpurun_out/q_trace.err | tail -1 | cut -c1-400
timeout 300 $B > gpurun_out/q_bench.json 2>/dev/null; python scripts/print_bench.py gpurun_out/q_bench.json | cut -c1-300
for d in 8320 16512 32896; do RUVGNN_FUSED_DEBUG=$d python scripts/dev/fused_dbg_run.py tf32 2>&1 | tail -1; done
