"""Multi-GPU plumbing: one process per GPU (torch.distributed, NCCL over NVLink / NVSwitch).

The path shards by RAYS: every rank holds the full primitive set and its own LBVH (replicated), and renders its
own camera views (`render_views`) -- or, when there are fewer views than ranks, its own band of 4-row tile strips of
a view (`render_tiles`: `render(..., rows=(y0, y1))` per rank + gather).  The forward pass needs no data-path
collective; images are gathered to rank 0 off the critical path.  Optimisation
needs exactly one exchange per step: the SUM of the per-primitive gradients, done as ONE all-reduce over a packed
[N * (10 + 1 + C)] buffer; every rank then applies the identical optimiser step and rebuilds its own LBVH, which is
deterministic, so the replicas stay bit-identical.
"""
from __future__ import annotations

from typing import Dict, List, Sequence

import torch
import torch.distributed as dist


def world():
    if dist.is_available() and dist.is_initialized():
        return dist.get_rank(), dist.get_world_size()
    return 0, 1


def shard_views(n_views: int, rank: int = None, world_size: int = None) -> List[int]:
    """Contiguous, balanced block of view indices for `rank` (64 views on 8 GPUs -> 8 each)."""
    if rank is None:
        rank, world_size = world()
    base, extra = divmod(n_views, world_size)
    start = rank * base + min(rank, extra)
    return list(range(start, start + base + (1 if rank < extra else 0)))


def shard_rows(height: int, rank: int = None, world_size: int = None, granule: int = 4):
    """Row band [y0, y1) of one image for `rank`, in multiples of the 4-row warp tile (views < GPUs)."""
    if rank is None:
        rank, world_size = world()
    strips = (height + granule - 1) // granule
    base, extra = divmod(strips, world_size)
    s0 = rank * base + min(rank, extra)
    s1 = s0 + base + (1 if rank < extra else 0)
    return min(s0 * granule, height), min(s1 * granule, height)


def shard_row_strips(height: int, rank: int = None, world_size: int = None, strips_per_rank: int = 1, granule: int = 4):
    """Row bands [(y0, y1), ...] of one image for `rank` when the image is cut into `strips_per_rank * world_size` strips
    dealt round-robin (strip j belongs to rank j % world_size): interleaving evens out the load when the primitives
    cover only part of the film.  `strips_per_rank=1` is `shard_rows`."""
    if rank is None:
        rank, world_size = world()
    n = max(1, strips_per_rank) * world_size
    out = []
    for j in range(rank, n, world_size):
        y0, y1 = shard_rows(height, j, n, granule)
        if y1 > y0:
            out.append((y0, y1))
    return out


def _padded(n: int, granule: int = 4) -> int:
    return (n + granule - 1) // granule * granule


def pack_gradients(grads: Dict[str, torch.Tensor], keys: Sequence[str]) -> torch.Tensor:
    """One flat fp32 buffer; every segment starts on a 16-byte boundary (the fused optimiser kernel and the vector
    reductions of the adjoint need aligned rows, and primitive counts are arbitrary)."""
    some = grads[keys[0]]
    flat = torch.zeros(sum(_padded(grads[k].numel()) for k in keys), dtype=some.dtype, device=some.device)
    off = 0
    for k in keys:
        n = grads[k].numel()
        flat[off:off + n] = grads[k].reshape(-1)
        off += _padded(n)
    return flat


def unpack_gradients(flat: torch.Tensor, like: Dict[str, torch.Tensor], keys: Sequence[str]) -> Dict[str, torch.Tensor]:
    out, off = {}, 0
    for k in keys:
        n = like[k].numel()
        out[k] = flat[off:off + n].reshape(like[k].shape)
        off += _padded(n)
    return out


class GradientBucket:
    """The primitive gradients of one rank in ONE flat buffer that the adjoint kernels accumulate into directly and the
    all-reduce runs over -- no packing copy.

    Layout: the primitives are cut into `ranges` ([p0, p1), boundaries on multiples of 4); range c owns one contiguous
    slice [data rows p0..p1 | attr p0..p1 | sh rows p0..p1], every segment 16-byte aligned.  So the gradient of a range
    is reduced with ONE collective call as soon as the primitive-major pass has finished that range, while the next range
    is still being accumulated.  The kernels index their gradient arguments by the global primitive number; `pointers(c)`
    returns base tensors shifted such that rows p0..p1 land in range c's slice (rows outside it are never touched by a
    call restricted to that range).  With a single range the layout is the plain [data | attr | sh]."""

    def __init__(self, n: int, sh_floats: int, device, ranges=None):
        self.n, self.sh_floats = n, sh_floats
        self.ranges = list(ranges) if ranges else [(0, n)]
        per = 10 + 1 + sh_floats
        self.flat = torch.zeros(max(sum(_padded((p1 - p0) * 10) + _padded(p1 - p0) + _padded((p1 - p0) * sh_floats)
                                        for p0, p1 in self.ranges), 4), dtype=torch.float32, device=device)
        self._seg, off = [], 0
        for p0, p1 in self.ranges:
            m = p1 - p0
            d0 = off; off += _padded(m * 10)
            a0 = off; off += _padded(m)
            s0 = off; off += _padded(m * sh_floats)
            self._seg.append((d0, a0, s0, off))
        assert per >= 11

    def chunk_flat(self, c: int) -> torch.Tensor:
        d0, _, _, end = self._seg[c]
        return self.flat[d0:end]

    def chunk_views(self, c: int):
        """(data [m*10], attr [m], sh [m*C] | None) views of range c."""
        (p0, p1), (d0, a0, s0, _) = self.ranges[c], self._seg[c]
        m = p1 - p0
        return (self.flat[d0:d0 + m * 10], self.flat[a0:a0 + m],
                self.flat[s0:s0 + m * self.sh_floats] if self.sh_floats else None)

    def pointers(self, c: int):
        """Raw device addresses (g_data10, g_attr, g_sh) for a kernel call restricted to range c, shifted so that the
        kernels' global primitive index p lands at row p - p0 of the range's slice."""
        (p0, _), (d0, a0, s0, _) = self.ranges[c], self._seg[c]
        base = self.flat.data_ptr()
        return (base + 4 * (d0 - 10 * p0), base + 4 * (a0 - p0), base + 4 * (s0 - self.sh_floats * p0) if self.sh_floats else 0)

    def gather(self):
        """(data [N*10], attr [N], sh [N*C] | None) in the reference layouts (copies unless there is a single range)."""
        parts = [self.chunk_views(c) for c in range(len(self.ranges))]
        if len(parts) == 1:
            return parts[0]
        return (torch.cat([p[0] for p in parts]), torch.cat([p[1] for p in parts]),
                torch.cat([p[2] for p in parts]) if self.sh_floats else None)

    def zero_(self):
        self.flat.zero_()

    def all_reduce(self, group=None):
        if world()[1] > 1:
            dist.all_reduce(self.flat, op=dist.ReduceOp.SUM, group=group)

    def all_reduce_chunk(self, c: int, group=None):
        """Asynchronous SUM over the slice of range c (one collective); returns the work handle in a list."""
        if world()[1] == 1:
            return []
        return [dist.all_reduce(self.chunk_flat(c), op=dist.ReduceOp.SUM, group=group, async_op=True)]


def chunk_ranges(n: int, n_chunks: int, granule: int = 4, taper: bool = False):
    """[p0, p1) primitive ranges with boundaries on multiples of `granule` (aligned slices): about equal sizes, or with
    `taper` sizes falling like n_chunks : ... : 2 : 1 -- the all-reduce of the LAST range cannot overlap anything, so it
    should be the smallest."""
    n_chunks = max(1, min(n_chunks, max(1, n // granule)))
    if not taper:
        step = _padded((n + n_chunks - 1) // n_chunks, granule)
        out, p = [], 0
        while p < n:
            out.append((p, min(n, p + step)))
            p += step
        return out
    total = n_chunks * (n_chunks + 1) // 2
    out, p, acc = [], 0, 0
    for c in range(n_chunks):
        acc += n_chunks - c
        q = n if c + 1 == n_chunks else min(n, _padded(n * acc // total, granule))
        if q > p:
            out.append((p, q))
        p = q
    return out


def allreduce_gradients(grads: Dict[str, torch.Tensor], group=None, async_op: bool = False):
    """SUM the primitive gradients of all ranks with a single all-reduce over one packed fp32 buffer
    (236 MB at 1M primitives / SH3).  Returns the reduced dict (and the work handle when async_op)."""
    keys = sorted(grads)
    rank, ws = world()
    if ws == 1:
        return (grads, None) if async_op else grads
    flat = pack_gradients(grads, keys)
    work = dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=group, async_op=async_op)
    out = unpack_gradients(flat, grads, keys)
    return (out, work) if async_op else out


def gather_images(local: Dict[int, torch.Tensor], n_views: int, dst: int = 0, group=None):
    """Collect {view index: image} from all ranks on `dst`; returns the ordered list there, None elsewhere."""
    rank, ws = world()
    if ws == 1:
        return [local[i] for i in range(n_views)]
    some = next(iter(local.values())) if local else None
    shape = list(some.shape) if some is not None else None
    shapes = [None] * ws
    dist.all_gather_object(shapes, shape, group=group)
    shape = next(s for s in shapes if s is not None)
    device = some.device if some is not None else (torch.device('cuda', torch.cuda.current_device())
                                                   if dist.get_backend(group) == 'nccl' else torch.device('cpu'))
    counts = [len(shard_views(n_views, r, ws)) for r in range(ws)]
    maxc = max(counts)
    mine = shard_views(n_views, rank, ws)
    buf = torch.zeros([maxc] + shape, dtype=torch.float32, device=device)
    for k, v in enumerate(mine):
        buf[k] = local[v]
    bufs = [torch.empty_like(buf) for _ in range(ws)] if rank == dst else None
    dist.gather(buf, bufs, dst=dst, group=group)
    if rank != dst:
        return None
    out = []
    for r in range(ws):
        for k in range(counts[r]):
            out.append(bufs[r][k])
    return out


def render_tiles(scene, sensor, render_fn, dst: int = 0, group=None, strips_per_rank: int = 1, **kw):
    """ONE view over all ranks (views < GPUs): every rank renders its 4-row-granular strips (`shard_row_strips`;
    `strips_per_rank` > 1 interleaves them for load balance) with `render_fn(scene, sensor=sensor, rows=(y0, y1), **kw)`;
    the strips are gathered on rank `dst` (ONE NCCL gather over NVLink when the process group is NCCL), which returns
    the assembled [H, W, 3] image; None elsewhere."""
    rank, ws = world()
    H, W = sensor.height, sensor.width
    mine = shard_row_strips(H, rank, ws, strips_per_rank)
    parts = [render_fn(scene, sensor=sensor, rows=b, **kw) for b in mine]
    if ws == 1:
        return parts[0] if len(parts) == 1 else torch.cat(parts, dim=0)
    layout = [shard_row_strips(H, r, ws, strips_per_rank) for r in range(ws)]
    max_rows = max(sum(b - a for a, b in bands) for bands in layout)
    device = parts[0].device if parts else (torch.device('cuda', torch.cuda.current_device())
                                            if dist.get_backend(group) == 'nccl' else torch.device('cpu'))
    buf = torch.zeros((max_rows, W, 3), dtype=torch.float32, device=device)
    at = 0
    for (a, b), part in zip(mine, parts):
        buf[at:at + b - a] = part
        at += b - a
    bufs = [torch.empty_like(buf) for _ in range(ws)] if rank == dst else None
    dist.gather(buf, bufs, dst=dst, group=group)
    if rank != dst:
        return None
    image = torch.empty((H, W, 3), dtype=torch.float32, device=device)
    for r, bands in enumerate(layout):
        at = 0
        for a, b in bands:
            image[a:b] = bufs[r][at:at + b - a]
            at += b - a
    return image


def render_views(scene, sensors: Sequence, render_fn, dst: int = 0, **kw):
    """Render `sensors` sharded by view over the ranks with `render_fn(scene, sensor=..., **kw)`; the images end
    up on rank `dst` in sensor order."""
    rank, ws = world()
    mine = shard_views(len(sensors), rank, ws)
    local = {i: render_fn(scene, sensor=sensors[i], **kw) for i in mine}
    return gather_images(local, len(sensors), dst)
