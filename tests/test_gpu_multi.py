"""Partitioned (multi-GPU) forward parity; needs >= 2 GPUs, otherwise skipped."""
import subprocess
import sys
from pathlib import Path

import pytest

pytestmark = pytest.mark.gpu
ROOT = Path(__file__).resolve().parents[1]


def test_partitioned_forward_matches_oracle_and_single_gpu(built):
    import torch

    ngpu = torch.cuda.device_count()
    if ngpu < 2:
        pytest.skip("needs >= 2 GPUs")
    world = 4 if ngpu >= 4 else 2
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}",
           "--master-addr", "127.0.0.1", "--master-port", "29533", str(ROOT / "scripts" / "multi_gpu_parity.py")]
    res = subprocess.run(cmd, capture_output=True, text=True, timeout=600, cwd=ROOT)
    print(res.stdout[-3000:], res.stderr[-3000:])
    assert res.returncode == 0 and "MULTI_GPU_PARITY OK" in res.stdout
