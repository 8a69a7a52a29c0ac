"""Multi-GPU host logic on CPU: LPT partition + the final gather over gloo with world_size 2."""
import os
import socket

import numpy as np
import pytest

from magpurify2_b200 import sharding


def test_lpt_partition_balances():
    rng = np.random.default_rng(51)
    w = rng.lognormal(15, 0.3, 1000)
    parts = sharding.lpt_partition(w, 8)
    assert sorted(np.concatenate(parts).tolist()) == list(range(1000))
    loads = np.array([w[p].sum() for p in parts])
    assert loads.max() / loads.mean() < 1.01
    # long-contig stress shape: a few huge items
    w2 = np.array([100.0, 90, 80, 1, 1, 1, 1, 1])
    parts2 = sharding.lpt_partition(w2, 3)
    assert max(w2[p].sum() for p in parts2) == 100.0
    assert sharding.lpt_partition([], 4) and all(len(p) == 0 for p in sharding.lpt_partition([], 4))


def _worker(rank, world, port, q):
    import torch
    import torch.distributed as dist

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    rng = np.random.default_rng(52)
    sizes = rng.integers(1, 40, 37)            # contigs per MAG
    bp = rng.lognormal(15, 0.5, 37)
    mine = sharding.mags_of_rank(bp, rank, world)
    # "score" of contig j of MAG m (stands in for the per-contig GPU results)
    rows = np.concatenate([np.stack([np.full(sizes[m], m, np.float64), np.arange(sizes[m], dtype=np.float64)], 1)
                           for m in mine.tolist()]) if len(mine) else np.zeros((0, 2))
    full = sharding.gather_rows(torch.from_numpy(rows), mine, sizes)
    want = np.concatenate([np.stack([np.full(sizes[m], m, np.float64), np.arange(sizes[m], dtype=np.float64)], 1)
                           for m in range(37)])
    q.put((rank, bool(np.array_equal(full.numpy(), want))))
    dist.destroy_process_group()


def test_gather_rows_gloo_world2():
    import torch.multiprocessing as mp

    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in procs)
    for p in procs:
        p.join(60)
    assert res == [(0, True), (1, True)]
