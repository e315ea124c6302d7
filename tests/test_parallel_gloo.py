"""CPU: the N>1 host logic (view sharding, packed gradient all-reduce, image gather) with world_size-2 gloo."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from volprim_balance_b200 import parallel


def test_shards_partition_views_and_rows():
    for n, ws in ((64, 8), (8, 8), (7, 2), (3, 8), (1, 2)):
        got = sum((parallel.shard_views(n, r, ws) for r in range(ws)), [])
        assert got == list(range(n))
        sizes = [len(parallel.shard_views(n, r, ws)) for r in range(ws)]
        assert max(sizes) - min(sizes) <= 1
    for h, ws in ((1080, 8), (1080, 3), (6, 4)):
        bands = [parallel.shard_rows(h, r, ws) for r in range(ws)]
        assert bands[0][0] == 0 and bands[-1][1] == h
        assert all(bands[i][1] == bands[i + 1][0] for i in range(ws - 1))
        assert all(b[0] % 4 == 0 for b in bands if b[1] > b[0])


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, ws, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=ws)
    try:
        n, C = 11, 12
        g = {"data": torch.full((n * 10,), float(rank + 1)), "opacities": torch.arange(n, dtype=torch.float32) * (rank + 1),
             "sh_coeffs": torch.ones(n * C) * (10 ** rank)}
        red = parallel.allreduce_gradients(g)
        ok = bool((red["data"] == 3).all() and torch.equal(red["opacities"], torch.arange(n, dtype=torch.float32) * 3)
                  and (red["sh_coeffs"] == 11).all())
        n_views = 5
        local = {v: torch.full((4, 6, 3), float(v)) for v in parallel.shard_views(n_views)}
        imgs = parallel.gather_images(local, n_views, dst=0)
        if rank == 0:
            ok = ok and len(imgs) == n_views and all(float(imgs[v].mean()) == v for v in range(n_views))
        else:
            ok = ok and imgs is None
        q.put((rank, ok))
    finally:
        dist.destroy_process_group()


@pytest.mark.timeout(120)
def test_allreduce_and_gather_world_size_2():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=100) for _ in range(2))
    for p in procs:
        p.join(timeout=30)
    assert res == [(0, True), (1, True)]
