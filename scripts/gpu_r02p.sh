#!/bin/bash
# Round 2, session P/T: fused K-upd iteration -- its parity tests, the phase timeline, the d=128 line.
set -u
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_fused.py tests/test_gpu_multipass.py tests/test_store.py -m gpu -x -q > gpurun_out/p_pytest.log 2>&1
rc=$?
echo "pytest exit $rc"; tail -5 gpurun_out/p_pytest.log | cut -c1-300
if [ $rc -ne 0 ]; then exit 0; fi
B="python bench.py --nodes 1000000 --dim 128 --heads 4 --steps 5 --warmup 3 --no-cpu-baseline --no-e2e --no-secondary"
RUVGNN_FUSED_DEBUG=16 timeout 300 $B > gpurun_out/p_trace.json 2> gpurun_out/p_trace.err; grep "upd_fused trace" gpurun_out/p_trace.err | tail -1
python scripts/print_bench.py gpurun_out/p_trace.json 2>/dev/null | cut -c1-300
