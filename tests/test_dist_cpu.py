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
