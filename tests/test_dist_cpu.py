"""CPU, world_size 2 over gloo: the multi-GPU host logic (row/unit sharding and the packed (distance, row) key whose
MAX all-reduce gives np.argmax semantics across ranks).  The per-shard arithmetic here is test-only numpy."""
import os
import socket

import numpy as np
import pytest
import torch.distributed as dist
import torch.multiprocessing as mp

from ssdr_al_b200 import dist as D


def test_shard_range_and_items_cover_everything():
    for n in (0, 1, 7, 8, 500_000):
        for world in (1, 2, 3, 8):
            spans = [D.shard_range(n, world, r) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            assert max(e - b for b, e in spans) - min(e - b for b, e in spans) <= 1
            items = sorted(i for r in range(world) for i in D.shard_items(n if n < 100 else 37, world, r))
            assert items == list(range(n if n < 100 else 37))


def test_packed_key_orders_like_argmax():
    rng = np.random.default_rng(0)
    d = rng.random(1000).astype(np.float32)
    d[[10, 500, 900]] = d.max() + 1  # ties: lowest row must win
    keys = D.pack_candidate(d, np.arange(1000))
    dist_, row = D.unpack_candidate(keys.max())
    assert row == 10 == int(np.argmax(d)) and dist_ == d[10]
    assert (D.unpack_candidate(keys)[1] == np.arange(1000)).all()


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, F, n_samples, first, out_q):
    import torch
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    N = len(F)
    b, e = D.shard_range(N, world, rank)
    mind = np.full(e - b, 1e10, np.float32)
    picks = [first]
    for _ in range(n_samples - 1):
        dcur = np.sum((F[b:e] - F[picks[-1]]) ** 2, axis=-1)
        mind = np.minimum(mind, dcur)
        j = int(np.argmax(mind))
        key = D.pack_candidate(mind[j], b + j)
        t = torch.tensor([int(key) - (1 << 63)], dtype=torch.int64)  # order-preserving shift into int64 for gloo
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        _, row = D.unpack_candidate(np.uint64(int(t.item()) + (1 << 63)))
        picks.append(int(row))
    out_q.put((rank, picks))
    dist.barrier()
    dist.destroy_process_group()


def test_row_sharded_fps_matches_single_process_oracle(oracle):
    rng = np.random.default_rng(5)
    base = rng.integers(0, 3, (60, 8)).astype(np.float32)
    F = np.concatenate([rng.standard_normal((700, 8)).astype(np.float32), base[rng.integers(0, 60, 500)]])  # with ties
    n_samples, first = 60, 3
    want = oracle.fps(F, n_samples, first)
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, F, n_samples, first, q)) for r in range(2)]
    for p in procs:
        p.start()
    got = dict(q.get(timeout=120) for _ in procs)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert got[0] == got[1] == want.tolist()


def test_balanced_slabs_are_contiguous_and_balanced():
    rng = np.random.default_rng(1)
    for world in (1, 2, 3, 8):
        for cnt in (rng.integers(0, 1000, 500), np.array([100, 1, 1, 1]), np.array([0, 0, 0, 9]), np.array([3]),
                    np.zeros(5, np.int64), np.array([], np.int64)):
            b = D.balanced_slabs(cnt, world)
            assert b[0] == 0 and b[-1] == len(cnt) and (np.diff(b) >= 0).all() and len(b) == world + 1
            per_rank = [int(np.sum(cnt[b[r]:b[r + 1]])) for r in range(world)]
            assert sum(per_rank) == int(np.sum(cnt))
            if len(cnt):  # no slab exceeds the ideal share by more than one layer
                assert max(per_rank) <= -(-int(np.sum(cnt)) // world) + int(np.max(cnt))
    dest, send = D.route_plan([0, 1, 2, 3, 4, 5], D.balanced_slabs([1, 2, 3, 4, 5, 6], 3))
    assert dest.tolist() == [0, 0, 0, 0, 1, 2] and send.tolist() == [4, 1, 1]


def _route_worker(rank, world, port, layers, rows, out_q):
    import torch
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    b, e = [(0, 1000), (1000, len(layers))][rank]  # uneven chunks, rank order == row order
    # slab bounds from the summed layer histogram, then this rank's rows grouped by owner in input order (on a GPU the
    # library does this partition, ssdr_grid_route_dev; here numpy's stable sort stands in) and one all-to-all
    hist = torch.from_numpy(np.bincount(layers[b:e], minlength=int(layers.max()) + 1))
    dist.all_reduce(hist)
    bounds = D.balanced_slabs(hist.numpy(), world)
    dest, send = D.route_plan(layers[b:e], bounds)
    order = np.argsort(dest, kind="stable")
    r, none = D.exchange_groups((torch.from_numpy(rows[b:e][order]), None), send.tolist())
    assert none is None
    out_q.put((rank, bounds.tolist(), r.numpy()))
    dist.barrier()
    dist.destroy_process_group()


def test_slab_route_keeps_global_row_order_per_owner():
    """Every row reaches the rank owning its layer, and each rank sees its rows in global input order (what makes the
    sharded barycentres bit-identical to a single-GPU run)."""
    rng = np.random.default_rng(9)
    n = 3000
    layers = rng.integers(0, 40, n).astype(np.int32)
    rows = np.stack([np.arange(n), layers], 1).astype(np.float32)
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_route_worker, args=(r, 2, port, layers, rows, q)) for r in range(2)]
    for p in procs:
        p.start()
    got = {r: (b, a) for r, b, a in (q.get(timeout=120) for _ in procs)}
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    bounds = got[0][0]
    assert bounds == got[1][0] == D.balanced_slabs(np.bincount(layers, minlength=40), 2).tolist()
    for r in range(2):
        mask = (layers >= bounds[r]) & (layers < bounds[r + 1])
        assert np.array_equal(got[r][1], rows[mask])


def _gather_worker(rank, world, port, sizes, out_q):
    import torch
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    got = []
    for case in sizes:  # rows of rank r: r * 1000 + 0, 1, ... in column 0
        n = case[rank]
        rows = torch.stack([torch.arange(n, dtype=torch.float32) + 1000 * rank, torch.full((n,), float(rank))], 1)
        full, span = D.gather_rows(rows)
        got.append((full.numpy(), span))
    out_q.put((rank, got))
    dist.barrier()
    dist.destroy_process_group()


def test_gather_rows_of_uneven_and_empty_slabs():
    """The slabs of a sharded subsampling differ in length and may be empty: every rank gets all rows in rank order
    and the span of its own block (gloo takes the padded all-gather; NCCL moves each block once, unpadded)."""
    cases = [(5, 3), (0, 4), (7, 0), (1, 1)]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_gather_worker, args=(r, 2, port, cases, q)) for r in range(2)]
    for p in procs:
        p.start()
    got = dict(q.get(timeout=120) for _ in procs)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for i, (n0, n1) in enumerate(cases):
        want = np.concatenate([np.stack([np.arange(n0, dtype=np.float32), np.zeros(n0, np.float32)], 1),
                               np.stack([np.arange(n1, dtype=np.float32) + 1000, np.ones(n1, np.float32)], 1)])
        for r in range(2):
            full, span = got[r][i]
            assert np.array_equal(full, want)
            assert tuple(span) == ((0, n0) if r == 0 else (n0, n0 + n1))
