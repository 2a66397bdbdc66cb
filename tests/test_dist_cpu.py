"""Multi-rank host logic on CPU (gloo, world_size 2 and 3): partitioning, halo plan,
request exchange, per-layer row exchange.  The data plane is emulated with index_select
(no GPU here); the plan under test is the one the GPU path uses unchanged."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from ruvector_b200 import dist as rvdist
from ruvector_b200 import synth
from ruvector_b200.graph import CsrGraph


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _slice_rows(g: CsrGraph, lo: int, hi: int) -> CsrGraph:
    s, e = int(g.row_ptr[lo]), int(g.row_ptr[hi])
    return CsrGraph((g.row_ptr[lo:hi + 1] - g.row_ptr[lo]).astype(np.uint64), g.col_idx[s:e].copy(), g.edge_w[s:e].copy())


def test_plan_remaps_without_reordering():
    n, world = 1000, 4
    g = synth.hnsw_like_graph(n, 8, seed=3, locality=0.5, window=50)
    bounds = rvdist.even_bounds(n, world)
    for rank in range(world):
        lo, hi = int(bounds[rank]), int(bounds[rank + 1])
        plan = rvdist.HaloPlan.build(_slice_rows(g, lo, hi), rank, bounds)
        assert plan.n_local == hi - lo and plan.recv_counts.sum() == plan.n_halo
        assert plan.recv_counts[rank] == 0  # never asks itself
        # de-duplicated, sorted, all remote
        assert (np.diff(plan.need_ids) > 0).all()
        assert ((plan.need_ids < lo) | (plan.need_ids >= hi)).all()
        # mapping back through [own rows | halo slots] reproduces the global columns in order
        table = np.concatenate([np.arange(lo, hi), plan.need_ids])
        s, e = int(g.row_ptr[lo]), int(g.row_ptr[hi])
        np.testing.assert_array_equal(table[plan.graph.col_idx], g.col_idx[s:e])
        assert plan.graph.edge_w.tobytes() == g.edge_w[s:e].tobytes()
        np.testing.assert_array_equal(plan.graph.row_ptr, g.row_ptr[lo:hi + 1] - g.row_ptr[lo])
        # requests are peer-local indices
        for p in range(world):
            assert ((plan.req_local[p] >= 0) & (plan.req_local[p] < bounds[p + 1] - bounds[p])).all()


def test_plan_edge_cases():
    bounds = rvdist.even_bounds(10, 2)
    empty = CsrGraph(np.zeros(6, np.uint64), np.zeros(0, np.uint32), np.zeros(0, np.float32))
    plan = rvdist.HaloPlan.build(empty, 0, bounds)
    assert plan.n_halo == 0 and plan.graph.n_edges == 0
    plan.set_requests([np.zeros(0, np.int64), np.zeros(0, np.int64)])
    assert plan.send_idx.size == 0
    with pytest.raises(ValueError):
        rvdist.HaloPlan.build(CsrGraph(np.array([0, 1, 1, 1, 1, 1], np.uint64), np.array([10], np.uint32),
                                       np.ones(1, np.float32)), 0, bounds)
    with pytest.raises(ValueError):
        plan.set_requests([np.array([7]), np.zeros(0, np.int64)])
    assert rvdist.even_bounds(10, 3).tolist() == [0, 3, 6, 10]


def _worker(rank, world, port, n, width, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        g = synth.hnsw_like_graph(n, 6, seed=11, locality=0.3, window=40)  # same global graph on every rank
        table = (np.arange(n * width, dtype=np.float32).reshape(n, width) * 0.5)  # "projected rows"
        bounds = rvdist.even_bounds(n, world)
        lo, hi = int(bounds[rank]), int(bounds[rank + 1])
        plan = rvdist.exchange_requests(rvdist.HaloPlan.build(_slice_rows(g, lo, hi), rank, bounds))
        local = torch.from_numpy(table[lo:hi])
        send = local.index_select(0, torch.from_numpy(plan.send_idx.astype(np.int64)))  # = ruvgnn_halo_pack
        halo = rvdist.exchange_rows(plan, send)
        ok = np.array_equal(halo.numpy(), table[plan.need_ids])
        full = np.concatenate([table[lo:hi], halo.numpy()])
        s, e = int(g.row_ptr[lo]), int(g.row_ptr[hi])
        ok = ok and np.array_equal(full[plan.graph.col_idx], table[g.col_idx[s:e]])
        ok = ok and int(plan.send_counts.sum()) == plan.send_idx.size
        # dense (all-gather) data plane: remote rows addressed by global id
        dense = rvdist.choose_dense(plan, threshold=0.0)
        gathered = [torch.empty_like(local) for _ in range(world)] if (hi - lo) * world == n else None
        if dense and gathered is not None:
            dist.all_gather(gathered, local)
            block = np.concatenate([t.numpy() for t in gathered])
            full_d = np.concatenate([table[lo:hi], block])
            ok = ok and np.array_equal(full_d[rvdist.dense_columns(plan)], table[g.col_idx[s:e]])
        ok = ok and (dense == ((hi - lo) * world == n))  # unequal partitions must fall back
        q.put((rank, bool(ok), plan.n_halo))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_halo_exchange_over_gloo(world):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, 600, 12, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert sorted(r[0] for r in res) == list(range(world))
    assert all(r[1] for r in res) and all(r[2] > 0 for r in res)


def test_partition_rows_generator_matches_global_semantics():
    part = synth.hnsw_like_graph(200, 6, seed=5, n_total=800, row_offset=400, iso_frac=0.0)
    assert part.n_nodes == 200 and part.col_idx.max() < 800 and part.col_idx.min() >= 0
    for i in (0, 57, 199):
        s, e = int(part.row_ptr[i]), int(part.row_ptr[i + 1])
        row = part.col_idx[s:e]
        assert (400 + i) not in row and len(set(row.tolist())) == len(row)
    assert (part.col_idx >= 600).any() and (part.col_idx < 400).any()  # references other partitions


def test_interior_runs_tile_the_local_rows_and_touch_no_halo():
    n, world = 6000, 3
    g = synth.hnsw_like_graph(n, 8, seed=5, locality=1.0, window=40, iso_frac=0.02)
    bounds = rvdist.even_bounds(n, world)
    for rank in range(world):
        lo, hi = int(bounds[rank]), int(bounds[rank + 1])
        plan = rvdist.HaloPlan.build(_slice_rows(g, lo, hi), rank, bounds)
        interior, boundary = rvdist.interior_runs(plan, min_rows=64, max_runs=3)
        both = sorted(interior + boundary)
        assert both[0][0] == 0 and both[-1][1] == plan.n_local
        assert all(a[1] == b[0] for a, b in zip(both, both[1:])) and all(b > a for a, b in both)
        rp = plan.graph.row_ptr.astype(np.int64)
        for a, b in interior:
            assert b - a >= 64
            assert (plan.graph.col_idx[rp[a]:rp[b]] < plan.n_local).all()
        # a +-40 id window leaves at most ~40 rows at each end of the range on the boundary
        assert sum(b - a for a, b in interior) >= plan.n_local - 2 * 40 - 2
        # every run is maximal: the row just outside an interior run has a remote neighbour
        for a, b in interior:
            if a > 0:
                assert (plan.graph.col_idx[rp[a - 1]:rp[a]] >= plan.n_local).any()
    # insertion-order worst case: nothing is interior; tiny min_rows keeps at most max_runs runs
    g0 = synth.hnsw_like_graph(n, 8, seed=6, locality=0.0)
    plan0 = rvdist.HaloPlan.build(_slice_rows(g0, 0, int(bounds[1])), 0, bounds)
    interior0, boundary0 = rvdist.interior_runs(plan0, min_rows=64)
    assert interior0 == [] and boundary0 == [(0, plan0.n_local)]
    few, rest = rvdist.interior_runs(plan0, min_rows=1, max_runs=2)
    assert len(few) <= 2 and sum(b - a for a, b in few + rest) == plan0.n_local


@pytest.mark.parametrize("chunks", [1, 3, 4])
def test_dense_columns_address_the_chunked_all_gather_layout(chunks):
    n_per, world = 1001, 4
    n = n_per * world
    g = synth.hnsw_like_graph(n, 8, seed=11, locality=0.0)
    bounds = rvdist.even_bounds(n, world)
    starts = rvdist.chunk_starts(n_per, chunks)
    assert starts[0] == 0 and starts[-1] == n_per and (np.diff(starts) > 0).all()
    # what the chunked all-gather leaves in the buffer: piece c of every rank, ordered by owner
    gathered = np.concatenate([np.arange(bounds[r] + starts[c], bounds[r] + starts[c + 1])
                               for c in range(len(starts) - 1) for r in range(world)])
    assert gathered.size == n
    for rank in (0, 2, 3):
        lo, hi = int(bounds[rank]), int(bounds[rank + 1])
        rows = _slice_rows(g, lo, hi)
        plan = rvdist.HaloPlan.build(rows, rank, bounds)
        col = rvdist.dense_columns(plan, starts).astype(np.int64)
        orig = rows.col_idx.astype(np.int64)
        local = (orig >= lo) & (orig < hi)
        assert np.array_equal(col[local], orig[local] - lo)
        assert (col[~local] >= n_per).all()
        assert np.array_equal(gathered[col[~local] - n_per], orig[~local])
