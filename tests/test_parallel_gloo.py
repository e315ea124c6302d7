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
    for h, ws, k in ((1080, 8, 4), (1080, 3, 5), (6, 4, 2), (50, 2, 1)):     # interleaved strips tile the rows exactly once
        strips = sorted(b for r in range(ws) for b in parallel.shard_row_strips(h, r, ws, k))
        assert strips[0][0] == 0 and strips[-1][1] == h
        assert all(a[1] == b[0] for a, b in zip(strips, strips[1:]))
        rows = [sum(b - a for a, b in parallel.shard_row_strips(h, r, ws, k)) for r in range(ws)]
        assert max(rows) - min(rows) <= 4 * k + 4


def test_tapered_ranges_tile_the_primitives_and_shrink():
    """RefineStep's all-reduce ranges: contiguous, 4-aligned boundaries, sizes falling towards the end (the last range's
    all-reduce is the exposed one), degenerate counts handled."""
    for n, k in ((1_000_000, 4), (30_000, 3), (1001, 6), (7, 4), (4, 1), (0, 3)):
        r = parallel.chunk_ranges(n, k, taper=True)
        if n == 0:
            assert r == []
            continue
        assert r[0][0] == 0 and r[-1][1] == n and all(a[1] == b[0] for a, b in zip(r, r[1:]))
        assert all(p0 % 4 == 0 for p0, _ in r) and all(p1 > p0 for p0, p1 in r) and len(r) <= k
        sizes = [p1 - p0 for p0, p1 in r]
        assert all(a >= b - 4 for a, b in zip(sizes, sizes[1:]))
        if n >= 1000 and len(r) == k and k > 1:
            assert sizes[-1] < n / k          # smaller than an even split
        b = parallel.GradientBucket(n, 12, "cpu", ranges=r)
        assert b.flat.numel() >= n * (10 + 1 + 12)


def test_packed_gradient_segments_are_aligned_and_ranges_tile():
    for n in (11, 12, 1001):
        g = {"data": torch.arange(n * 10.0), "opacities": torch.arange(n * 1.0), "sh_coeffs": torch.arange(n * 27.0)}
        keys = sorted(g)
        flat = parallel.pack_gradients(g, keys)
        back = parallel.unpack_gradients(flat, g, keys)
        assert all(torch.equal(back[k], g[k]) for k in g)
        assert all((back[k].data_ptr() - flat.data_ptr()) % 16 == 0 for k in g)      # any N, aligned slices
        for chunks in (1, 3, 4, 64):
            r = parallel.chunk_ranges(n, chunks)
            assert r[0][0] == 0 and r[-1][1] == n and all(a[1] == c[0] for a, c in zip(r, r[1:]))
            assert all(p0 % 4 == 0 for p0, _ in r)
            b = parallel.GradientBucket(n, 27, "cpu", ranges=r)
            base = b.flat.data_ptr()
            for c, (p0, p1) in enumerate(r):
                views = b.chunk_views(c)
                assert [v.numel() for v in views] == [(p1 - p0) * 10, p1 - p0, (p1 - p0) * 27]
                assert all((v.data_ptr() - base) % 16 == 0 for v in views)            # aligned segments for any N
                # the shifted base pointers put global row p0 at the start of the range's own segments
                gd, ga, gs = b.pointers(c)
                assert (gd + 40 * p0, ga + 4 * p0, gs + 108 * p0) == tuple(v.data_ptr() for v in views)
                assert gd % 8 == 0 and ga % 4 == 0
                views[0].fill_(c + 1)
            data, attr, sh = b.gather()
            assert data.numel() == n * 10 and attr.numel() == n and sh.numel() == n * 27
            assert all(bool((data[10 * p0:10 * p1] == c + 1).all()) for c, (p0, p1) in enumerate(r))
            assert sum(b.chunk_flat(c).numel() for c in range(len(r))) == b.flat.numel() or n < 4


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
        ranges = parallel.chunk_ranges(n, 3)
        b = parallel.GradientBucket(n, C, "cpu", ranges=ranges)
        for c in range(len(ranges)):
            d_, a_, s_ = b.chunk_views(c)
            d_ += rank + 1
            s_ += 10 ** rank
        for c in range(len(ranges)):
            for w in b.all_reduce_chunk(c):
                w.wait()
        data, attr, sh = b.gather()
        ok = ok and bool((data == 3).all() and (sh == 11).all() and (attr == 0).all())

        class _S:   # image-tile sharding: every rank renders its row band, rank 0 assembles
            width, height = 6, 10
        band_of = lambda scene, sensor, rows: torch.arange(rows[0], rows[1], dtype=torch.float32)[:, None, None].expand(-1, 6, 3)
        tiled = parallel.render_tiles(None, _S, band_of)
        if rank == 0:
            ok = ok and tiled.shape == (10, 6, 3) and torch.equal(tiled[:, 0, 0], torch.arange(10.0))
        else:
            ok = ok and tiled is None

        class _S2:
            width, height = 6, 50
        tiled = parallel.render_tiles(None, _S2, band_of, strips_per_rank=3)        # interleaved strips
        if rank == 0:
            ok = ok and tiled.shape == (50, 6, 3) and torch.equal(tiled[:, 0, 0], torch.arange(50.0))
        else:
            ok = ok and tiled is None
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
