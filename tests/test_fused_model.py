"""NumPy models of the arithmetic the fused K-upd kernel (`ruvector_b200/csrc/gemm_upd_fused.cu`) relies on
(no GPU needed): the fold of `w_agg` into the out-projection (layer.rs:159,407,408), the 3xTF32 split of an
operand the epilogue writes back into TMEM, the thread-private LayerNorm sums that must reproduce the
xor-butterfly of `layernorm_fast_kernel<1>` bit for bit, and the k-block bookkeeping of the kernel's rings."""
import numpy as np
import pytest

F32 = np.float32


def tf32_hi(x):
    """Round-to-nearest of the low 13 mantissa bits, as the kernels do it ((bits + 0x1000) & 0xFFFFE000)."""
    b = np.asarray(x, F32).view(np.uint32)
    return ((b + np.uint32(0x1000)) & np.uint32(0xFFFFE000)).view(F32)


def gemm_3xtf32(a, w):
    """a_lo w_hi + a_hi w_lo + a_hi w_hi with FP32-grade accumulation (modelled in float64), dropping a_lo w_lo."""
    ah, wh = tf32_hi(a), tf32_hi(w)
    al, wl = (a - ah).astype(F32), (w - wh).astype(F32)
    f = np.float64
    return (al.astype(f) @ wh.astype(f).T + ah.astype(f) @ wl.astype(f).T + ah.astype(f) @ wh.astype(f).T)


@pytest.mark.parametrize("heads", [1, 2, 4, 8])
def test_folded_first_stage_equals_out_projection_then_w_agg(heads):
    """a = W_agg (W_ov S + b_ov + agg) + b_agg  ==  [W_agg W_ov | W_agg] [S | agg] + (W_agg b_ov + b_agg):
    exactly in real arithmetic, to FP32 rounding with the products formed in double and stored as f32
    (upload_layer), and still inside 1e-6 after the 3xTF32 split of both operands."""
    H, rows = 128, 64
    rng = np.random.default_rng(heads)
    scale = 1.0 / np.sqrt(H)
    Wov = (rng.standard_normal((H, heads * H)) * scale / np.sqrt(heads)).astype(F32)
    bov = (rng.standard_normal(H) * 0.1).astype(F32)
    Wagg = (rng.standard_normal((H, H)) * scale).astype(F32)
    bagg = (rng.standard_normal(H) * 0.1).astype(F32)
    S = rng.standard_normal((rows, heads * H)).astype(F32)
    agg = rng.standard_normal((rows, H)).astype(F32)
    f = np.float64
    exact = (S.astype(f) @ Wov.astype(f).T + bov + agg) @ Wagg.astype(f).T + bagg
    Wf = np.concatenate([(Wagg.astype(f) @ Wov.astype(f)).astype(F32), Wagg], axis=1)  # [H][(heads + 1) H]
    bf = (Wagg.astype(f) @ bov.astype(f) + bagg).astype(F32)
    A = np.concatenate([S, agg], axis=1)
    assert Wf.shape == (H, (heads + 1) * H) and Wf.shape[1] % 64 == 0  # an even number of 32-float k-blocks
    folded = A.astype(f) @ Wf.astype(f).T + bf
    ref = np.linalg.norm(exact, axis=1)
    assert (np.linalg.norm(folded - exact, axis=1) / ref).max() < 2e-7
    split = gemm_3xtf32(A, Wf) + bf
    assert (np.linalg.norm(split - exact, axis=1) / ref).max() < 1e-6
    single = tf32_hi(A).astype(f) @ tf32_hi(Wf).astype(f).T + bf
    assert (np.linalg.norm(single - exact, axis=1) / ref).max() > 1e-5  # why one TF32 pass is not a parity mode


def test_split_written_back_by_an_epilogue_is_exact():
    """hi + lo == x exactly, hi has 13 zero low bits, |lo| <= 2^-11 |x|: what an epilogue thread stores into
    P | Q is the value itself, so the next GEMM sees the FP32 operand up to the dropped lo x lo term."""
    rng = np.random.default_rng(3)
    x = (rng.standard_normal(100000) * np.exp(rng.uniform(-20, 20, 100000))).astype(F32)
    hi = tf32_hi(x)
    lo = (x - hi).astype(F32)
    assert np.array_equal((hi.astype(np.float64) + lo.astype(np.float64)).astype(F32), x)
    assert not (hi.view(np.uint32) & np.uint32(0x1FFF)).any()
    assert (np.abs(lo) <= np.abs(x) * 2.0 ** -11).all()


def butterfly_sum(p):
    """warp_sum of rowops.cu: v += shfl_xor(v, o) for o = 16, 8, 4, 2, 1 on 32 lane partials (float32)."""
    v = p.astype(F32).copy()
    lanes = np.arange(32)
    for o in (16, 8, 4, 2, 1):
        v = (v + v[lanes ^ o]).astype(F32)
    return v


def tree_sum(p):
    """The fused kernel's thread-private form: p[j] += p[j + w] for w = 16, 8, 4, 2, 1."""
    p = p.astype(F32).copy()
    w = 16
    while w > 0:
        p[:w] = (p[:w] + p[w:2 * w]).astype(F32)
        w //= 2
    return p[0]


def test_thread_private_tree_reproduces_the_warp_butterfly_bit_for_bit():
    """Every lane of the butterfly ends with the same bits (float addition commutes), and those are the bits
    of the sequential tree: the LayerNorm rows written by the fused kernel and by layernorm_fast_kernel<1> agree
    exactly, which the sparse-level form (a mix of both) depends on."""
    rng = np.random.default_rng(5)
    for trial in range(200):
        x = (rng.standard_normal(128) * (1.0 + 10.0 * rng.random())).astype(F32)
        x4 = x.reshape(32, 4)
        part = ((x4[:, 0] + x4[:, 1]).astype(F32) + (x4[:, 2] + x4[:, 3]).astype(F32)).astype(F32)
        b = butterfly_sum(part)
        assert (b.view(np.uint32) == b.view(np.uint32)[0]).all()
        assert tree_sum(part).view(np.uint32) == b.view(np.uint32)[0]
        mean = F32(b[0] * F32(1.0 / 128))
        d = (x4 - mean).astype(F32)
        sq = np.zeros(32, F32)
        for k in (3, 2, 1, 0):  # fmaf(x, x, fmaf(y, y, fmaf(z, z, fmaf(w, w, 0)))): innermost first
            sq = (d[:, k].astype(np.float64) * d[:, k].astype(np.float64) + sq.astype(np.float64)).astype(F32)
        assert tree_sum(sq).view(np.uint32) == butterfly_sum(sq).view(np.uint32)[0]


@pytest.mark.parametrize("k1", [256, 384, 640, 1152])
def test_ring_bookkeeping_of_one_tile(k1):
    """Order in which the weight k-blocks are streamed (TMA warp) and consumed (MMA warp), and the slot / parity
    sequence of the two TMEM rings, for the stage-1 widths the layer shapes produce ((heads + 1) * 128 or 256)."""
    nk1 = k1 // 32
    # weight stream per tile: Wf k-blocks, then per gate GEMM a0 m0 a1 m1 a2 m2 a3 m3, then Wh 0..7
    stream = [("F", j, 0) for j in range(nk1)]
    for half in range(2):
        for t in range(8):
            stream.append(("ZR", (t >> 1) + (t & 1) * 4, half * 128))
    stream += [("H", j, 0) for j in range(8)]
    assert len(stream) == nk1 + 24
    # what the MMA warp consumes, phase by phase
    consume = [("F", kb, 0) for kb in range(nk1)]
    for half in range(2):
        for kb in range(4):
            consume += [("ZR", kb, half * 128), ("ZR", 4 + kb, half * 128)]  # resident a k-block, streamed m k-block
    consume += [("H", kb, 0) for kb in range(4)] + [("H", 4 + kb, 0) for kb in range(4)]
    assert consume == stream
    # A-operand stream of the producer warps: gather state, then the message row twice (Z and R)
    a_stream = [("S", i) for i in range(nk1)] + [("m", j & 3) for j in range(8)]
    assert [kb for name, kb in a_stream if name == "m"] == [0, 1, 2, 3, 0, 1, 2, 3]
    # stage-1 ring: 4 slots in P | Q, wait parity (use & 1) ^ 1 for the producer, use & 1 for the consumer
    uses = {}
    for tile in range(3):
        for i in range(nk1):
            slot = (tile * nk1 + i) % 4
            u = uses.get(slot, 0)
            assert u == (tile * nk1 + i) // 4
            uses[slot] = u + 1
    # TMEM column plan: X | Y | P | Q, 128 columns each, 64 columns (hi | lo) per streamed k-block
    cols = {"X": 0, "Y": 128, "P": 256, "Q": 384}
    s1_slots = [cols["P"] + 64 * s for s in range(4)]
    m_slots = [cols["Y"] + 64 * s for s in range(2)]
    assert s1_slots[-1] + 64 == 512 and m_slots[-1] + 64 == cols["P"]


def test_cycle_model_reproduces_the_recorded_timeline():
    """scripts/fused_upd_model.py (per-MMA issue cost by commit density + exposed epilogues) against the clock64
    timeline recorded on a B200 (profiles/r02t_upd_fused_timeline.txt, slot 9 = end of the tile's last GEMM)."""
    import importlib.util
    import re
    from pathlib import Path

    root = Path(__file__).resolve().parent.parent
    spec = importlib.util.spec_from_file_location("fused_upd_model", root / "scripts" / "fused_upd_model.py")
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    trace = dict((int(k), int(v)) for k, v in re.findall(r"\[(\d+)\]=(\d+)", (root / "profiles" / "r02t_upd_fused_timeline.txt").read_text()))
    measured = trace[9]
    assert abs(mod.tile_cycles(640) - measured) / measured < 0.05
    # the phases the model is built from, as measured: S1, Z, R, C1, C2 issue spans
    assert abs((trace[1] - trace[0]) - 20 * 12 * 138.0) / (trace[1] - trace[0]) < 0.05
    assert abs((trace[3] - trace[2]) - 8 * 12 * 103.0) / (trace[3] - trace[2]) < 0.05
    assert abs((trace[9] - trace[8]) - 4 * 12 * 90.0) / (trace[9] - trace[8]) < 0.05
