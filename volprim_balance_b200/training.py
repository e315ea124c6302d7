"""One optimisation step of the reference's multi-view refinement loop, scheduled for one process per GPU.

Reference: examples/refine_3dg_dataset.py:170-189 --
    image = mi.render(scene, params, sensor=batch_sensor, spp, seed=it); loss = l1(ref_image, image)
    dr.backward(loss); opt.step(); update_params(opt)        (update_params -> Ellipsoid.ravel + params.update())
Here the views of the batch sensor are sharded over the ranks (primitives and LBVH replicated).  Per view: primal pass
recording compressed hit lists -> L1 gradient of the view -> gather adjoint into ONE flat gradient buffer.  The only
exchange of the step is the SUM of that buffer; it is cut into primitive ranges, and the all-reduce of range c runs on
NCCL's stream while the primitive-major accumulation of range c + 1 (last view of the rank) is still running.  Then
every rank applies the identical BoundedAdam step and refits / rebuilds its own LBVH.

`loss.backward()` through `volprim_balance_b200.render` gives the same gradients (tested); this class exists because
autograd cannot split the backward pass of the last view into ranges.
"""
from __future__ import annotations

import torch
import torch.distributed as dist

from . import _cabi, parallel
from .accel import RaySource
from .optimizers import l1_loss_grad
from .integrators.common import Ellipsoid


class RefineStep:
    def __init__(self, scene, sensors, targets, opt, n_chunks: int = 4, group=None, rebuild: str = 'rebuild',
                 rebuild_every: int = 8):
        """sensors: all views of the batch (PerspectiveSensor objects, identical film sizes); targets: {view index:
        [H, W, 3] CUDA tensor} for at least the views of this rank; opt: BoundedAdam holding 'centers', 'scales',
        'quats', 'opacities', 'sh_coeffs' (the keys of refine_3dg_dataset.py:131-153).  rebuild: 'rebuild' = a full LBVH
        build after every step (what params.update() does in the reference), 'refit' = keep order and topology and
        refresh the boxes, with a full build every `rebuild_every` steps (learning rates of 1e-4 barely move a
        primitive between two steps)."""
        self.scene, self.integrator, self.shape = scene, scene.integrator, scene.ellipsoids()
        if self.integrator.integrator_id != _cabi.INTEGRATOR_RF:
            raise Exception("RefineStep drives the volprim_rf integrator")
        self.sensors, self.targets, self.opt, self.group = list(sensors), targets, opt, group
        self.rank, self.world = parallel.world()
        self.views = parallel.shard_views(len(self.sensors), self.rank, self.world)
        self.n_chunks = n_chunks
        self.shape.rebuild_policy = rebuild
        self.rebuild_every, self._steps = max(1, int(rebuild_every)), 0
        n = self.shape.count
        shf = self.shape.attributes['sh_coeffs'].numel() // n if n else 0
        # several ranks: tapered ranges, the last (whose all-reduce is exposed) the smallest; one rank: nothing to overlap
        self.ranges = parallel.chunk_ranges(n, n_chunks if self.world > 1 else 1, taper=True)
        self.bucket = parallel.GradientBucket(n, shf, self.shape.device, ranges=self.ranges)
        self._sums = torch.zeros(2, dtype=torch.float32, device=self.shape.device)
        self.record = None
        self.timing = {}
        s0 = self.sensors[0]
        self.n_pix = len(self.sensors) * s0.width * s0.height * 3
        self._pinned_totals = torch.zeros((max(len(self.sensors), 1), 2), dtype=torch.int64).pin_memory()
        # records over capacity / rays cut at the per-ray cap, over all views and ranks: known once the LAST primal pass
        # is done, i.e. long before the adjoint of that view is -- the host takes the redo decision while the GPU works
        self._flags = torch.zeros(2, dtype=torch.float32, device=self.shape.device)
        self._pinned_flags = torch.zeros(2, dtype=torch.float32).pin_memory()
        self._flags_ready = torch.cuda.Event()
        self._side = torch.cuda.Stream(device=self.shape.device)
        self.update_params()

    # refine_3dg_dataset.py:155-159
    def update_params(self):
        o = self.opt
        self.shape.data = Ellipsoid.ravel(o['centers'].detach(), o['scales'].detach(), o['quats'].detach()).contiguous()
        self.shape.attributes['opacities'] = o['opacities'].detach().reshape(-1).contiguous()
        self.shape.attributes['sh_coeffs'] = o['sh_coeffs'].detach().reshape(-1).contiguous()
        self.shape.parameters_changed()
        if self.shape.rebuild_policy == 'refit' and self._steps % self.rebuild_every == 0:
            self.shape._topology_n = -1          # periodic full build
        self.shape.bind('opacities', with_sh=True)

    def _ensure_record(self, acc, n_rays, id_cap):
        cap = int(n_rays * min(float(id_cap), acc.hits_per_ray_estimate * 1.3 + 4.0)) + 4096
        if self.record is None or self.record.capacity < cap or self.record.n_rays != n_rays or self.record.id_cap != id_cap:
            self.record = acc.new_record(n_rays, id_cap, cap)
        return self.record

    def step(self, want_images: bool = False):
        """One optimisation step.  Returns (loss, sum of squared errors / n_pix[, images]) as 0-d CUDA tensors (global
        over all ranks)."""
        for _ in range(3):
            out = self._step_once(want_images)
            if out is not None:
                return out
        raise _cabi.VolprimCudaError("RefineStep: hit records kept overflowing their capacity")

    def accumulate_gradients(self, want_images: bool = False):
        """Primal + adjoint of this rank's views into the flat gradient buffer; with several ranks the all-reduce of
        every primitive range is launched (asynchronously, on NCCL's stream) as soon as the last view has accumulated
        that range.  Returns (bucket, NCCL work handles, (loss, squared error[, images], flags)) -- local, unreduced
        statistics; the caller waits on the handles."""
        acc, integ, shape = self.shape.accel(), self.integrator, self.shape
        params = integ._vp_params(self.scene, None)
        id_cap = integ._cap()
        self.bucket.zero_()
        self._sums.zero_()
        flags = self._flags
        torch.cuda.current_stream().wait_event(self._flags_ready)     # the previous step's copy of the flags
        flags.zero_()
        works, images = [], {}
        n_ranges = len(self.ranges)
        for k, vi in enumerate(self.views):
            s = self.sensors[vi]
            rays = RaySource(camera=s.vp_camera(), spp=1)
            rec = self._ensure_record(acc, rays.n_rays, id_cap)
            res = acc.render_forward(params, rays, record=rec, id_cap=id_cap, want_beta=False, want_nhits=False)
            self._pinned_totals[k % self._pinned_totals.shape[0]].copy_(rec.total, non_blocking=True)
            flags += torch.stack([(rec.total[0] > rec.capacity).float(), (rec.total[1] > 0).float()])
            last = k + 1 == len(self.views)
            if last:
                self._post_flags()
            # l1 over the batch film, its gradient and the squared error in one pass (optimizers.py:170-186)
            dL, _ = l1_loss_grad(self.targets[vi], res.rgb.reshape(s.height, s.width, 3), n_total=self.n_pix, sums=self._sums)
            if want_images:
                images[vi] = res.rgb.reshape(s.height, s.width, 3)
            acc.adjoint_begin(params, rays, dL.reshape(-1, 3), res.rgb, rec, self.bucket.pointers(0))
            for c, (p0, p1) in enumerate(self.ranges):
                acc.adjoint_finish(params, rays, rec, p0, p1, self.bucket.pointers(c))
                if last and self.world > 1:
                    # the all-reduce of range c runs on NCCL's stream while range c + 1 is still being accumulated
                    works += self.bucket.all_reduce_chunk(c, self.group)
        if not self.views:
            self._post_flags()
            if self.world > 1:
                for c in range(n_ranges):
                    works += self.bucket.all_reduce_chunk(c, self.group)
        loss, sq = self._sums[0], self._sums[1]
        return self.bucket, works, ((loss, sq, images, flags) if want_images else (loss, sq, flags))

    def _post_flags(self):
        """Sum the overflow flags over the ranks and start their copy to pinned memory, on a side stream: the main stream
        goes on with the adjoint, the host reads the decision as soon as this rank's primal passes are done."""
        main = torch.cuda.current_stream()
        self._side.wait_stream(main)
        with torch.cuda.stream(self._side):
            if self.world > 1:
                dist.all_reduce(self._flags, group=self.group)
            self._pinned_flags.copy_(self._flags, non_blocking=True)
            self._flags_ready.record(self._side)

    def _step_once(self, want_images):
        acc = self.shape.accel()
        ev = {k: torch.cuda.Event(enable_timing=True) for k in ('start', 'compute_done', 'comm_done', 'end')}
        ev['start'].record()
        _, works, stats_local = self.accumulate_gradients(want_images)
        loss, sq, flags = stats_local[0], stats_local[1], stats_local[-1]
        images = stats_local[2] if want_images else None
        ev['compute_done'].record()
        # Every record must have fitted its capacity (the kernels skip an unusable record), and ALL ranks must take
        # the same decision: the flags were summed over the ranks right after the last primal pass, so the host
        # has them while the GPU is still busy with the adjoint and can queue the optimiser behind it without a gap.
        self._flags_ready.synchronize()
        overflowed, cut = self._pinned_flags.tolist()
        if len(self.views):
            entries = int(self._pinned_totals[:len(self.views), 0].max())
            acc.hits_per_ray_estimate = max(4.0, entries / max(self.record.n_rays, 1))
        for w in works:
            w.wait()
        ev['comm_done'].record()
        if cut:
            raise _cabi.VolprimCudaError("RefineStep: a ray recorded more hits than the integrator's record cap")
        if overflowed:
            return None                # redo the step with records sized from the new estimate
        stats = torch.stack([loss, sq])
        if self.world > 1:
            dist.all_reduce(stats, group=self.group)
        loss, sq = stats[0], stats[1]
        # identical optimiser step on every rank (refine_3dg_dataset.py:178-189)
        g_data, g_attr, g_sh = self.bucket.gather()
        g = g_data.view(-1, 10)
        o = self.opt
        o['centers'].grad = g[:, 0:3].contiguous()
        o['scales'].grad = g[:, 3:6].contiguous()
        o['quats'].grad = g[:, 6:10].contiguous()
        o['opacities'].grad = g_attr.reshape(o['opacities'].shape)      # views of the bucket: valid until the next step
        o['sh_coeffs'].grad = g_sh.reshape(o['sh_coeffs'].shape)
        o.step()
        self._steps += 1
        self.update_params()
        ev['end'].record()
        torch.cuda.current_stream().synchronize()
        self.timing = {'step_ms': ev['start'].elapsed_time(ev['end']),
                       'exposed_allreduce_ms': ev['compute_done'].elapsed_time(ev['comm_done']),
                       'optimizer_and_rebuild_ms': ev['comm_done'].elapsed_time(ev['end']),
                       'allreduce_bytes': self.bucket.flat.numel() * 4 if self.world > 1 else 0}
        return (loss, sq, images) if want_images else (loss, sq)
