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


@pytest.mark.parametrize("d,heads,mode", [(384, 8, "tc"), (256, 8, "simt"), (128, 4, "tc"), (768, 8, "tc"), (96, 3, "simt")])
def test_rebuilt_halo_rows_carry_the_owner_bits(built, d, heads, mode):
    """Message-only halo: K|V rebuilt from a received message row (ruvgnn_layer_expand_halo) must be
    bit-identical to the K|V|M row the owner projected (what ruvgnn_halo_pack would have sent)."""
    import ctypes as C

    import numpy as np
    import torch

    import ruvector_b200 as rv
    from ruvector_b200 import _lib, synth

    lib = _lib.load()
    # the torch tensors below are produced on torch's current stream: the ctx must enqueue on the
    # same one (a ctx created without a stream owns a non-blocking stream that does not order
    # against it)
    stream = torch.cuda.Stream()
    torch.cuda.set_stream(stream)
    ctx = rv.Context(0, stream.cuda_stream)
    n = 3000
    layer = rv.RuvectorLayer(d, d, heads, 0.1, weights=synth.make_weights(d, d, seed=d), ctx=ctx,
                             mode=rv.MODE_FP32_TC if mode == "tc" else rv.MODE_FP32_SIMT)
    h = layer._device_layer()
    x = torch.from_numpy(synth.make_features(n, d, seed=3)).cuda()
    _lib.check(lib.ruvgnn_layer_project(ctx.handle, h, n, C.c_void_p(x.data_ptr())))
    W = int(lib.ruvgnn_layer_halo_width(h))
    idx = torch.arange(n, dtype=torch.int32, device="cuda")
    sent = torch.empty(n, W, device="cuda")
    _lib.check(lib.ruvgnn_halo_pack(ctx.handle, h, C.c_void_p(idx.data_ptr()), n, C.c_void_p(sent.data_ptr())))
    msgs = torch.empty(n, d, device="cuda")
    _lib.check(lib.ruvgnn_layer_messages(ctx.handle, h, 0, n, C.c_void_p(msgs.data_ptr())))
    rebuilt = torch.full((n, W), float("nan"), device="cuda")
    # an odd row offset / count: the rebuild must not depend on where a row sits in a tile
    a, b = 37, n - 5
    _lib.check(lib.ruvgnn_layer_expand_halo(ctx.handle, h, C.c_void_p(msgs[a:b].data_ptr()), b - a,
                                            C.c_void_p(rebuilt[a:b].data_ptr())))
    ctx.sync()
    assert np.array_equal(sent[a:b].cpu().numpy(), rebuilt[a:b].cpu().numpy())
    assert torch.isnan(rebuilt[:a]).all() and torch.isnan(rebuilt[b:]).all()
    # fewer SMs for the grids changes nothing
    _lib.check(lib.ruvgnn_ctx_set_sm_limit(ctx.handle, -16))
    again = torch.empty(n, W, device="cuda")
    _lib.check(lib.ruvgnn_layer_expand_halo(ctx.handle, h, C.c_void_p(msgs.data_ptr()), n, C.c_void_p(again.data_ptr())))
    _lib.check(lib.ruvgnn_ctx_set_sm_limit(ctx.handle, 0))
    ctx.sync()
    assert np.array_equal(sent.cpu().numpy(), again.cpu().numpy())
    torch.cuda.set_stream(torch.cuda.default_stream())
