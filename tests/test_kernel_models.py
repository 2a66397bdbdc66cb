"""NumPy models of the warp-level index algebra the gather kernels rely on (no GPU needed).

They pin, on the CPU, the layouts that `ruvector_b200/csrc/gather.cu` hard-codes: the select-free
halving butterfly (`reduce_scatter_xor`), the lane-permuted slots of `gather_r1_batched_kernel`
and its xor-shuffle broadcasts, and the "phys" column order of the fast layouts.  A 32-lane warp is
an array axis of length 32; `shfl_xor(v, m)[L] = v[L ^ m]`.
"""
import numpy as np
import pytest

LANES = np.arange(32)


def shfl_xor(v, mask):
    return v[LANES ^ mask]


def reduce_scatter_xor(v):
    """v: [32 lanes][N positions] -> per-lane scalar (gather.cu: reduce_scatter_xor<N>)."""
    v = v.copy()
    n, off = v.shape[1], 16
    while n > 1:
        for i in range(n // 2):
            v[:, i] = v[:, i] + shfl_xor(v[:, i + n // 2], off)
        n //= 2
        off //= 2
    while off > 0:
        v[:, 0] = v[:, 0] + shfl_xor(v[:, 0], off)
        off //= 2
    return v[:, 0]


@pytest.mark.parametrize("heads", [1, 2, 4])
def test_lane_permuted_butterfly_leaves_one_edge_head_total_per_lane(heads):
    eb = 8
    n, rep = eb * heads, 32 // (eb * heads)
    rng = np.random.default_rng(heads)
    partial = rng.integers(-50, 50, size=(32, eb, heads)).astype(np.float64)  # [lane][edge slot][head] partial dots
    lane_id = LANES // rep                     # the value index a lane ends up owning
    my_slot, my_head = LANES >> 2, lane_id % heads
    assert np.array_equal(my_slot, lane_id // heads)
    # what the kernel stores: position u*heads+g of lane L = partial of edge u ^ my_slot, head g ^ my_head
    v = np.empty((32, n))
    for u in range(eb):
        for g in range(heads):
            v[:, u * heads + g] = partial[LANES, u ^ my_slot, g ^ my_head]
    # i.e. position p holds value index p ^ lane_id
    for p in range(n):
        j = p ^ lane_id
        assert np.array_equal(v[:, p], partial[LANES, j // heads, j % heads])
    got = reduce_scatter_xor(v)
    want = partial.sum(axis=0)                 # [edge slot][head] totals over the 32 lanes
    assert np.array_equal(got, want[my_slot, my_head])


@pytest.mark.parametrize("heads", [1, 2, 4])
def test_xor_broadcasts_deliver_the_weight_each_slot_needs(heads):
    eb = 8
    rep = 32 // (eb * heads)
    my_slot, my_head = LANES >> 2, (LANES // rep) % heads
    # after the butterfly lane L holds p[(edge my_slot, head my_head)]; encode the pair as a number
    p = (my_slot * 10 + my_head).astype(np.float64)
    for u in range(eb):
        for g in range(heads):
            got = shfl_xor(p, (u * heads + g) * rep)          # kernel: __shfl_xor_sync(p, (u*HEADS+g)*REP)
            want = (u ^ my_slot) * 10 + (g ^ my_head)         # the pair stored in slot (u, g) of this lane
            assert np.array_equal(got, want)
    for g in range(heads):                                    # per-head scalars (corr, normaliser): mask g*REP
        got = shfl_xor(my_head.astype(np.float64), g * rep)
        assert np.array_equal(got, g ^ my_head)
    # step maximum / normaliser per head: lanes of one head differ in lane bits 2..4 only
    for m in (4, 8, 16):
        assert np.array_equal(shfl_xor(my_head, m), my_head)
    # the single prefetch instruction covers the 8 rows x 4 lines of the next step exactly once
    assert len({(int(s), int(l)) for s, l in zip(my_slot, LANES & 3)}) == 32


@pytest.mark.parametrize("h,heads", [(128, 4), (256, 8), (384, 8), (768, 8), (128, 1)])
def test_phys_column_order_gives_every_lane_whole_heads(h, heads):
    """phys float4 index i*32 + lane holds logical float4 lane*P + i (P = H/128): a warp-wide 128-bit load
    reads 512 contiguous bytes and lane L owns logical floats [4P*L, 4P*(L+1))."""
    P = h // 128
    phys_of_logical = np.empty(h, dtype=np.int64)
    for lane in range(32):
        for i in range(P):
            for c in range(4):
                phys_of_logical[(lane * P + i) * 4 + c] = (i * 32 + lane) * 4 + c
    assert np.array_equal(np.sort(phys_of_logical), np.arange(h))
    owner = np.empty(h, dtype=np.int64)       # lane that holds each logical float
    owner[np.arange(h)] = (phys_of_logical // 4) % 32
    assert np.array_equal(owner, np.arange(h) // (4 * P))
    d_h = h // heads
    lanes_per_head = 32 // heads
    for g in range(heads):                    # a head's floats live in a contiguous group of 32/heads lanes
        lanes = np.unique(owner[g * d_h:(g + 1) * d_h])
        assert np.array_equal(lanes, np.arange(g * lanes_per_head, (g + 1) * lanes_per_head))
