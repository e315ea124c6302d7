#!/usr/bin/env python
"""bench.py -- headline benchmark: volprim_rf forward render, Mrays/s (BASELINE.json configs[1]).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload cfg2|cfg3|cfg5|small]

One "step" = one batch of VIEWS_PER_STEP volprim_rf forward passes, each over one 1920x1080 view (1 spp, pixel-centre
rays) of the synthetic 1M-primitive Gaussian cloud, SH degree 3 (cfg 2) -- 24 views, so that the
driver's 20 timed steps last about two seconds (a sustained figure).  With N > 1 (torchrun, one rank per GPU) every
rank renders its own views of the replicated cloud -- the forward path shards by view with no data-path collective, so
its scaling is "weak" and `value` is the sum over ranks.

Printed JSON (rank 0, one line): the driver contract plus
  roofline     -- k_trace_forward: algorithmic bytes per launch / CUDA-event duration vs measured HBM peak
  cpu_baseline -- the CPU oracle (restatement of the reference loop, C + OpenMP) on one full view (or a contiguous crop)
  e2e          -- the same metric through volprim_balance_b200.render_to_host(): camera host->device, trace (rays
                  generated in-kernel), image device->host into pinned memory, all inside the timed region
  train_step   -- the optimisation step of examples/refine_3dg_dataset.py on a FIXED batch of 8 views (cfg 4), sharded
                  over the N ranks: forward (recording) + gather adjoint per view, gradient all-reduce (NCCL) cut into
                  primitive ranges and overlapped with the last view's accumulation, BoundedAdam, LBVH rebuild.  Strong
                  scaling: the same batch at every N.
  north_star   -- forward + adjoint per view on the 3M Epanechnikov cloud (cfg 3, the north-star target), N = 1 only
  build        -- LBVH build time for 1M / 3M / 10M primitives, N = 1 only
`--impl reference` times the CPU restatement itself on the SAME rays as this arm's views (one full view per step, or a
contiguous crop of it when the host is slow); Mitsuba/Dr.Jit cannot be installed here.
"""
from __future__ import annotations

import argparse
import hashlib
import json
import os
import statistics
import subprocess
import sys
import tempfile
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WORKLOADS = {
    # name: (n_prims, crossings (calibrated so that mean processed hits/ray ~ 32), W, H, n_views)
    "cfg2": dict(n=1_000_000, crossings=60.0, W=1920, H=1080, views=8, seed=1,
                 desc="volprim_rf forward, 1M Gaussian ellipsoids SH3, 1920x1080, 1 spp (BASELINE configs[1])"),
    "small": dict(n=100_000, crossings=40.0, W=640, H=360, views=8, seed=1,
                  desc="volprim_rf forward, 100k Gaussian ellipsoids SH3, 640x360 (smoke-size)"),
    # sigma0 calibrated on the GPU path (scripts/calibrate.py).  cfg3: with Epanechnikov kernels the processed hits
    # saturate at ~32.7/ray for mu_o = -1 whatever sigma0 is (the 0.01 cut-off ends the ray), so mu_o = -1.5 is used
    # to reach the 48 hits/ray BASELINE.md asks for.
    "cfg3": dict(n=3_000_000, sigma0=0.0053, W=1920, H=1080, views=8, seed=2, centers="truck", kernel="epanechnikov",
                 mu_opacity=-1.5, desc="volprim_rf forward, 3M Epanechnikov ellipsoids SH3 (truck-scale), 1920x1080 "
                                       "(BASELINE configs[2], one view per step per GPU)"),
    "cfg5": dict(n=10_000_000, sigma0=0.00197, W=3840, H=2160, views=8, seed=4, centers="dense", kernel="gaussian",
                 mu_opacity=-4.0, max_depth=-1, desc="stress: 10M overlapping Gaussian ellipsoids SH3, 3840x2160, "
                                                     "~200 hits/ray, max_depth=-1 (BASELINE configs[4])"),
}
VIEWS_PER_STEP = 32        # one step = this many views (a >= 2 s timed region at the driver's 20 steps)
TRAIN_VIEWS = 8            # fixed global batch of the training step (BASELINE configs[3]: 8 views / step)
RAY_IO_BYTES = 44          # 28 B read (o, d, maxt) + 16 B written (rgb, T)        BASELINE.md section 5
EVAL_BYTES_SH3 = 236       # 40 geometry + 4 opacity + 192 SH per primitive evaluation
# gather-formulation adjoint: per recorded hit 4 B id read + 4 B rank written (counting pass), 20 B record
# (id + colour/transmittance state) + 4 B rank read + 32 B bucket entry written (ray pass), 32 B entry read (primitive pass);
# per primitive 236 B parameters read + 59 gradient floats read-modify-written; per ray 12 B image gradient + 4 B hit count
ADJ_HIT_BYTES = 96
ADJ_PRIM_BYTES_SH3 = 236 + 2 * 59 * 4
ADJ_RAY_BYTES = 16
REC_HIT_BYTES = 20         # what the recording forward adds per hit (id + state)


def adjoint_bytes(rays, hits, prims):
    return rays * ADJ_RAY_BYTES + hits * ADJ_HIT_BYTES + prims * ADJ_PRIM_BYTES_SH3


def measured_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """SM clock and throttle reasons sampled DURING the timed region: an NVML polling thread (2 ms period; the timed
    region of the default run lasts < 100 ms, too short for an `nvidia-smi -lms` child to start), with `nvidia-smi`
    as the fallback when NVML cannot be loaded."""
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index=0):
        self.index, self.proc, self.path = index, None, None
        self.rows, self.thread, self.stop, self.max_mhz, self.source = [], None, False, None, None
        try:
            import pynvml
            pynvml.nvmlInit()
            # under torchrun every rank sees all GPUs: LOCAL_RANK is the NVML index unless CUDA_VISIBLE_DEVICES remaps
            vis = os.environ.get("CUDA_VISIBLE_DEVICES")
            phys = int(vis.split(",")[index]) if vis and all(x.strip().isdigit() for x in vis.split(",")) else index
            self.nv, self.h = pynvml, pynvml.nvmlDeviceGetHandleByIndex(phys)
            self.max_mhz = float(pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM))
            self.source = "nvml"
        except Exception:
            self.nv = None

    def _poll(self):
        nv = self.nv
        bits = (("hw_slowdown", nv.nvmlClocksThrottleReasonHwSlowdown),
                ("hw_thermal_slowdown", nv.nvmlClocksThrottleReasonHwThermalSlowdown),
                ("sw_thermal_slowdown", nv.nvmlClocksThrottleReasonSwThermalSlowdown),
                ("sw_power_cap", nv.nvmlClocksThrottleReasonSwPowerCap))
        while not self.stop:
            try:
                mhz = float(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                mask = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                self.rows.append((mhz, [n for n, b in bits if mask & b]))
            except Exception:
                pass
            time.sleep(0.002)

    def __enter__(self):
        if self.nv is not None:
            import threading
            self.thread = threading.Thread(target=self._poll, daemon=True)
            self.thread.start()
            return self
        try:
            fd, self.path = tempfile.mkstemp(suffix=".csv")
            os.close(fd)
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=open(self.path, "w"), stderr=subprocess.DEVNULL)
            self.source = "nvidia-smi"
        except Exception:
            self.proc = None
        return self

    def __exit__(self, *exc):
        if self.thread is not None:
            self.stop = True
            self.thread.join(timeout=2)
        if self.proc is not None:
            time.sleep(0.15)
            self.proc.terminate()
            try:
                self.proc.wait(timeout=5)
            except Exception:
                self.proc.kill()

    def summary(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": []}
        if self.thread is not None and self.rows:
            out["sm_mhz"] = statistics.median(r[0] for r in self.rows)
            out["sm_max_mhz"] = self.max_mhz
            out["reasons"] = sorted({n for r in self.rows for n in r[1]})
            out["samples"] = len(self.rows)
            out["source"] = self.source
            return out
        try:
            rows = [r.strip().split(",") for r in open(self.path) if r.strip()]
            sm = [float(r[0]) for r in rows]
            out["sm_mhz"] = statistics.median(sm)
            out["sm_max_mhz"] = float(rows[0][1])
            names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
            for i, nme in enumerate(names):
                if any("Active" == r[2 + i].strip() for r in rows):
                    out["reasons"].append(nme)
            out["samples"] = len(rows)
            out["source"] = self.source
        except Exception:
            pass
        finally:
            try:
                if self.path:
                    os.unlink(self.path)
            except Exception:
                pass
        return out


def build_cloud(wl):
    from volprim_balance_b200 import synthetic
    sigma0 = wl.get("sigma0") or synthetic.sigma0_for_hits(wl["n"], wl["crossings"])
    return synthetic.make_cloud(wl["n"], sigma0, seed=wl["seed"], sh_degree=3, centers=wl.get("centers", "uniform"),
                                mu_opacity=wl.get("mu_opacity", -1.0))


def source_stamp():
    """Hash of the kernel sources: ties a measured ncu figure (profiles/traffic.json) to the build that is benched.
    (.git does not travel to the GPU box, so the commit id is not available there.)"""
    h = hashlib.sha256()
    base = os.path.join(ROOT, "volprim_balance_b200", "csrc")
    for f in sorted(os.listdir(base)):
        if f.endswith((".cu", ".cuh")) or f == "Makefile":
            h.update(open(os.path.join(base, f), "rb").read())
    h.update(open(os.path.join(ROOT, "include", "volprim_cuda.h"), "rb").read())
    return h.hexdigest()[:16]


def view_rays(wl, view, rows=None):
    """The rays of one view of the workload (numpy, the oracle's input); rows = (y0, y1) selects a contiguous crop."""
    from volprim_balance_b200 import synthetic
    cam = synthetic.ring_camera(view % wl["views"], wl["views"], wl["W"], wl["H"])
    o, d, mt = synthetic.camera_rays(cam)
    if rows is not None:
        sl = slice(rows[0] * wl["W"], rows[1] * wl["W"])
        o, d, mt = o[sl], d[sl], mt[sl]
    return o, d, mt


class CpuReference:
    """CPU restatement of the reference loop (oracle/, C + OpenMP, one closest-hit BVH query per hit) on this
    workload: the ONLY use of oracle/ outside the tests -- as the thing timed beside the GPU path, never on it."""

    def __init__(self, wl, cloud, threads=None):
        from oracle import oracle as O
        # torchrun exports OMP_NUM_THREADS=1: ask for every host core this process may use
        O.set_num_threads(threads or len(os.sched_getaffinity(0)))
        self.O, self.wl, self.cores = O, wl, O.num_threads()
        self.scene = O.Scene(cloud.data, cloud.opacities, cloud.sh_coeffs, cloud.extent)
        self.params = O.Params(integrator=O.RF, kernel=O.EPAN if wl.get("kernel") == "epanechnikov" else O.GAUSS,
                               max_depth=wl.get("max_depth", 128), srgb_primitives=True)

    def time_view(self, view, rows=None):
        o, d, mt = view_rays(self.wl, view, rows)
        t0 = time.perf_counter()
        res = self.scene.forward(self.params, o, d, mt)
        dt = time.perf_counter() - t0
        return o.shape[0], dt, float(res.nhits.mean())

    def crop_for(self, seconds):
        """Rows of the contiguous, vertically centred crop of a view that takes about `seconds` (the full view if it
        fits), from a probe on 32 centre rows."""
        H = self.wl["H"]
        n, dt, _ = self.time_view(0, (H // 2 - 16, H // 2 + 16))
        rate = n / dt
        rows = int(seconds * rate / self.wl["W"]) & ~3
        if rows >= H:
            return None, rate
        rows = max(rows, 32)
        y0 = ((H - rows) // 2) & ~3
        return (y0, y0 + rows), rate


def run_reference(args, wl, rank, world):
    """Reference arm: the CPU restatement on the same rays as our arm's views -- step s traces view s % 8 in full (a
    contiguous crop of it when the host needs more than ~6 s per view, so that the run ends within a few minutes)."""
    if rank != 0:
        return
    cloud = build_cloud(wl)
    ref = CpuReference(wl, cloud)
    rows, _ = ref.crop_for(6.0)
    times, total_rays, hits = [], 0, []
    for step in range(args.warmup + args.steps):
        n, dt, h = ref.time_view(step, rows)
        if step >= args.warmup:
            times.append(dt)
            total_rays += n
            hits.append(h)
    value = total_rays / sum(times) / 1e6
    what = "the full view" if rows is None else f"rows {rows[0]}..{rows[1]} of the view (contiguous crop)"
    sample = (f"CPU restatement of the reference loop (oracle/volprim_oracle.c, OpenMP, one BVH closest-hit query per hit); "
              f"Mitsuba llvm_ad_rgb is not installable here.  Each step traces {what} -- the same pixel-centre rays our arm "
              f"renders ({total_rays // len(times)} rays), mean hits/ray {sum(hits) / len(hits):.1f}")
    line = {
        "impl": "reference", "metric": "volprim_rf forward Mrays/s", "value": value, "unit": "Mrays/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": 1e3 * sum(times) / len(times), "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": workload_config(wl),
        "step_definition": f"one step = {what} of one view on the host CPU",
        "cpu_baseline": {"value": value, "unit": "Mrays/s", "cores": ref.cores, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": "Mrays/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


def workload_config(wl):
    """Identical for both arms: what is rendered."""
    return {"workload": wl["desc"], "primitives": wl["n"], "film": f'{wl["W"]}x{wl["H"]}', "spp": 1,
            "views": "ring of 8 cameras, radius 4, pixel-centre rays",
            "l2": "resident inputs (geometry %d MB + SH %d MB + ordering records %d MB + BVH %d MB) exceed the 126 MB L2 "
                  "and the view changes with every launch" % (wl["n"] * 48 // 2**20, wl["n"] * 192 // 2**20,
                                                              wl["n"] * 48 // 2**20, wl["n"] * 64 // 2**20)}


def make_scene(vp, wl, cloud, dev, data=None, opacities=None, sh=None):
    from volprim_balance_b200 import synthetic
    W, H, V = wl["W"], wl["H"], wl["views"]
    data = cloud.data if data is None else data
    scene_dict = {
        "type": "scene",
        "integrator": {"type": "volprim_rf", "max_depth": wl.get("max_depth", 128), "rr_depth": -1,
                       "kernel_type": wl.get("kernel", "gaussian")},
        "primitives": {"type": "ellipsoidsmesh", "centers": data[:, 0:3], "scales": data[:, 3:6],
                       "quaternions": data[:, 6:10], "opacities": (cloud.opacities if opacities is None else opacities)[:, None],
                       "sh_coeffs": cloud.sh_coeffs if sh is None else sh, "extent": 3.0},
    }
    for i in range(V):
        c = synthetic.ring_camera(i, V, W, H)
        scene_dict[f"cam_{i:04d}"] = {"type": "perspective", "fov": c.fov_x_deg, "fov_axis": "x",
                                      "to_world": vp.Transform4f(c.to_world), "near_clip": c.near_clip,
                                      "far_clip": c.far_clip,
                                      "film": {"type": "hdrfilm", "width": W, "height": H, "rfilter": {"type": "box"}}}
    return vp.load_dict(scene_dict, device=dev)


def time_fwd_adjoint(torch, acc, params, sensors, views, id_cap, R, n, shf, reps):
    """Forward (recording compressed hit lists) + gather adjoint, ms per view, kernels timed separately with events."""
    from volprim_balance_b200.accel import RaySource
    dev = acc.device
    dL = torch.randn((R, 3), device=dev) * (1.0 / R)
    gbuf = (torch.zeros(n * 10, device=dev), torch.zeros(n, device=dev), torch.zeros(n * shf, device=dev))
    rec = None
    ev = []
    hits = {}

    def one(v, timed):
        nonlocal rec
        rays = RaySource(camera=sensors[v].vp_camera(), spp=1)
        if rec is None:
            rec = acc.new_record(R, id_cap)
        e = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
        e[0].record()
        r = acc.render_forward(params, rays, record=rec, id_cap=id_cap, want_beta=False, want_nhits=False)
        e[1].record()
        acc.adjoint_begin(params, rays, dL, r.rgb, rec, gbuf)
        e[2].record()
        acc.adjoint_finish(params, rays, rec, 0, n, gbuf)
        e[3].record()
        if timed:
            ev.append((e, v))

    for v in views[:2]:
        one(v, False)
        torch.cuda.synchronize()
        entries, cut = rec.totals()
        if entries > rec.capacity or cut:
            acc.hits_per_ray_estimate = max(4.0, entries / R)
            rec = None
            one(v, False)
            torch.cuda.synchronize()
        hits[v] = rec.totals()[0]
    t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0.record()
    for k in range(reps):
        one(views[k % len(views)], True)
    t1.record()
    torch.cuda.synchronize()
    assert rec.usable(), "hit record overflowed during the timed fwd+adjoint region"
    per = lambda i, j: sum(e[i].elapsed_time(e[j]) for e, _ in ev) / len(ev)
    total_hits = rec.totals()[0]
    return {"fwd_adjoint_ms_per_view": t0.elapsed_time(t1) / reps, "forward_record_ms": per(0, 1), "adjoint_ray_pass_ms": per(1, 2),
            "adjoint_primitive_pass_ms": per(2, 3), "record_bytes_per_view": total_hits * 4 + (R + 1) * 8,
            "hits_last_view": total_hits}


def measure_build(torch, vp, dev):
    """LBVH build (Morton codes, radix sort, Karras hierarchy, box fit) for 1M / 3M / 10M primitives: device time of
    vp_build, clouds generated on the device (the build only looks at the 10-float records; SH rows are re-ordered)."""
    out = {}
    for n in (1_000_000, 3_000_000, 10_000_000):
        g = torch.Generator(device=dev)
        g.manual_seed(n)
        data = torch.empty((n, 10), device=dev)
        data[:, 0:3] = torch.rand((n, 3), generator=g, device=dev) * 2 - 1
        data[:, 3:6] = torch.exp(torch.randn((n, 3), generator=g, device=dev) * 0.5 - 5.5)
        q = torch.randn((n, 4), generator=g, device=dev)
        data[:, 6:10] = q / q.norm(dim=1, keepdim=True)
        sh = torch.randn((n, 48), generator=g, device=dev) * 0.1
        acc = vp.accel.EllipsoidAccel(dev)
        acc.set_primitives(data, torch.rand(n, generator=g, device=dev), sh, 3.0)
        acc.build()
        torch.cuda.synchronize()
        best, best_refit = None, None
        for _ in range(3):
            a, b, c = (torch.cuda.Event(enable_timing=True) for _ in range(3))
            a.record()
            acc.build()
            b.record()
            acc.refit()
            c.record()
            torch.cuda.synchronize()
            best = min(best or 1e9, a.elapsed_time(b))
            best_refit = min(best_refit or 1e9, b.elapsed_time(c))
        # the build streams ~ (8 passes x 2 x 12 B sort traffic + 40 B record + 4 C B colour rows in and out + 64 B node + 48 B
        # ordering record + 48 B SoA + 32 B leaf boxes) per primitive
        bytes_per_prim = 8 * 2 * 12 + 40 + 2 * 192 + 64 + 48 + 48 + 32
        out[f"{n // 1_000_000}M"] = {"build_ms": best, "refit_ms": best_refit,
                                     "build_GBps": n * bytes_per_prim / (best * 1e-3) / 1e9}
        acc.close()
        del data, sh, acc
        torch.cuda.empty_cache()
    return out


def run_ours(args, wl, rank, world, local_rank):
    import torch
    import torch.distributed as dist
    import volprim_balance_b200 as vp
    from volprim_balance_b200 import parallel, synthetic, training
    from volprim_balance_b200.accel import RaySource
    from volprim_balance_b200.integrators.common import Ellipsoid

    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    cloud = build_cloud(wl)
    W, H, V = wl["W"], wl["H"], wl["views"]
    R = W * H
    VPS = args.views_per_step
    scene = make_scene(vp, wl, cloud, dev)
    integ = scene.integrator
    shape = scene.ellipsoids()
    shape.bind("opacities", with_sh=True)       # upload + LBVH build (outside the timed region)
    acc = shape.accel()
    sensors = scene.sensors()
    params = integ._vp_params(scene, None)
    cams = [s.vp_camera() for s in sensors]
    my_view = lambda k: (k + rank * max(1, V // max(world, 1))) % V      # launch k of this rank

    # `value`: inputs resident in HBM before the timed region = the primitive cloud + LBVH; the rays of a view are
    # generated inside the trace kernel from the 76-byte sensor description (no ray buffers to keep resident)
    def launch(k):
        return acc.render_forward(params, RaySource(camera=cams[my_view(k)], spp=1), want_nhits=False)

    hits_per_view = {}
    for k in range(max(args.warmup, 3) * VPS if args.warmup_full else max(V, 3)):
        launch(k)
    for v in range(V):
        acc.render_forward(params, RaySource(camera=cams[v], spp=1), want_nhits=False)
        torch.cuda.synchronize()
        hits_per_view[v] = acc.stats()["hits"]
    for s in range(args.warmup):
        for j in range(VPS):
            launch(s * VPS + j)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- device-resident throughput (`value`) + per-launch kernel time (roofline) --------------------
    n_launch = args.steps * VPS
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(n_launch + 1)]
    barrier()
    with ClockSampler(local_rank) as clocks:
        for k in range(n_launch):
            ev[k].record()
            launch(k)
        ev[n_launch].record()
        barrier()
    total_ms = ev[0].elapsed_time(ev[n_launch])
    kern_ms = [ev[k].elapsed_time(ev[k + 1]) for k in range(n_launch)]
    clock_summary = clocks.summary()
    if world > 1:
        tmax = torch.tensor([total_ms], device=dev)
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
        total_ms = float(tmax.item())
    value = world * R * n_launch / (total_ms * 1e-3) / 1e6

    algo_bytes = [R * RAY_IO_BYTES + hits_per_view[my_view(k)] * EVAL_BYTES_SH3 for k in range(n_launch)]
    achieved = sum(algo_bytes) / (sum(kern_ms) * 1e-3) / 1e9
    peak, peak_src = measured_peak()
    mean_hits = sum(hits_per_view[my_view(k)] for k in range(n_launch)) / (n_launch * R)

    # ---- forward (recording compressed hit lists) + gather adjoint, ms / view ------------------------------------
    shf = cloud.sh_coeffs.shape[1]
    id_cap = 128 if wl.get("max_depth", 128) > 0 else 1024
    acc.hits_per_ray_estimate = max(acc.hits_per_ray_estimate, mean_hits * 1.05)
    fa = time_fwd_adjoint(torch, acc, params, sensors, [my_view(k) for k in range(V)], id_cap, R, cloud.n, shf,
                          reps=min(max(args.steps, 4), 8))
    barrier()

    # ---- end to end through the public API: render_to_host() ------------------------------------------------------
    # one call for the views of the timed region; every view's camera goes host->device (inside the kernel arguments)
    # and every view's image device->host (pinned ring of two), the copy of view i overlapping the trace of view i+1
    host_ring = [torch.empty((H, W, 3), dtype=torch.float32).pin_memory() for _ in range(2)]
    vp.render_to_host(scene, sensors=[my_view(k) for k in range(4)], out=host_ring, spp=1, jitter=False)
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    vp.render_to_host(scene, sensors=[my_view(k) for k in range(n_launch)], out=host_ring, spp=1, jitter=False)
    e1.record()
    barrier()
    e2e_ms = e0.elapsed_time(e1)
    if world > 1:
        tmax = torch.tensor([e2e_ms], device=dev)
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
        e2e_ms = float(tmax.item())
    e2e_value = world * R * n_launch / (e2e_ms * 1e-3) / 1e6
    import ctypes
    cam_bytes = ctypes.sizeof(vp._cabi.vp_camera)

    # ---- ONE view cut into row strips over the ranks, image gathered on rank 0 (strong scaling of a single view) ---
    tiles = None
    if not args.no_extras or world > 1:
        spr = 4 if world > 1 else 1
        one = lambda v: parallel.render_tiles(scene, sensors[v % V], vp.render, spp=1, jitter=False, strips_per_rank=spr)
        for v in range(3):
            one(v)
        barrier()
        t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        n_tv = 16
        t0.record()
        for v in range(n_tv):
            one(v)
        t1.record()
        barrier()
        tms = torch.tensor([t0.elapsed_time(t1) / n_tv], device=dev)
        if world > 1:
            dist.all_reduce(tms, op=dist.ReduceOp.MAX)
        tiles = {"what": f"one {W}x{H} view at a time through parallel.render_tiles(scene, sensor, render): {spr} interleaved "
                         f"row strip(s) per rank over {world} GPU(s), strips gathered into the image on rank 0 (NCCL gather)",
                 "scaling": "strong", "ms_per_view": float(tms.item()), "Mrays_per_s": R / float(tms.item()) / 1e3,
                 "strips_per_rank": spr}

    # ---- training step on a FIXED batch of TRAIN_VIEWS views, sharded over the ranks (strong scaling) --------------
    train = None
    if not args.no_train and wl.get("kernel", "gaussian") == "gaussian" and wl["n"] <= 3_000_000:
        train = run_train_step(args, torch, dist, vp, training, parallel, Ellipsoid, wl, cloud, scene, dev, rank, world)
    del scene, shape, acc
    torch.cuda.empty_cache()

    north = None
    build = None
    if world == 1 and args.workload == "cfg2" and not args.no_extras:
        wl3 = WORKLOADS["cfg3"]
        cloud3 = build_cloud(wl3)
        scene3 = make_scene(vp, wl3, cloud3, dev)
        sh3 = scene3.ellipsoids()
        sh3.bind("opacities", with_sh=True)
        acc3 = sh3.accel()
        p3 = scene3.integrator._vp_params(scene3, None)
        R3 = wl3["W"] * wl3["H"]
        sens3 = scene3.sensors()
        for v in range(3):
            acc3.render_forward(p3, RaySource(camera=sens3[v].vp_camera(), spp=1), want_nhits=False)
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        h3 = 0
        for v in range(8):
            acc3.render_forward(p3, RaySource(camera=sens3[v].vp_camera(), spp=1), want_nhits=False)
        b.record()
        torch.cuda.synchronize()
        fwd3 = a.elapsed_time(b) / 8
        for v in range(8):
            acc3.render_forward(p3, RaySource(camera=sens3[v].vp_camera(), spp=1), want_nhits=False)
            torch.cuda.synchronize()
            h3 += acc3.stats()["hits"]
        acc3.hits_per_ray_estimate = h3 / 8 / R3 * 1.08
        fa3 = time_fwd_adjoint(torch, acc3, p3, sens3, list(range(8)), 128, R3, cloud3.n, 48, reps=8)
        bytes3 = R3 * RAY_IO_BYTES + h3 / 8 * EVAL_BYTES_SH3
        north = {"workload": wl3["desc"], "primitives": wl3["n"], "hits_per_ray": h3 / 8 / R3,
                 "forward_ms_per_view": fwd3, "forward_Mrays_per_s": R3 / fwd3 / 1e3,
                 "forward_roofline_frac": bytes3 / (fwd3 * 1e-3) / 1e9 / peak, **fa3,
                 "fwd_adjoint_algorithmic_GBps": (bytes3 + h3 / 8 * REC_HIT_BYTES + adjoint_bytes(R3, h3 / 8, wl3["n"]))
                 / (fa3["fwd_adjoint_ms_per_view"] * 1e-3) / 1e9}
        north["fwd_adjoint_roofline_frac"] = north["fwd_adjoint_algorithmic_GBps"] / peak
        del scene3, sh3, acc3, cloud3
        torch.cuda.empty_cache()
        build = measure_build(torch, vp, dev)

    if rank != 0:
        return
    cpu = None
    if world == 1 and not args.no_cpu_baseline:
        # bounded sample: one full view of the workload if the host does it in ~25 s, else a contiguous crop
        ref = CpuReference(wl, cloud)
        rows, _ = ref.crop_for(25.0)
        n, dt, h = ref.time_view(0, rows)
        what = "the full view 0" if rows is None else f"rows {rows[0]}..{rows[1]} of view 0 (contiguous crop)"
        cpu = {"value": n / dt / 1e6, "unit": "Mrays/s", "cores": ref.cores, "kind": "port",
               "sample": f"CPU restatement of the reference loop (oracle/volprim_oracle.c, OpenMP, one BVH closest-hit query per "
                         f"hit) on {what}: {n} rays, mean hits/ray {h:.1f}, {dt:.2f} s"}
    adj_ms = fa["adjoint_ray_pass_ms"] + fa["adjoint_primitive_pass_ms"]
    adj_bytes = adjoint_bytes(R, fa["hits_last_view"], wl["n"])
    line = {
        "metric": "volprim_rf forward Mrays/s", "value": value, "unit": "Mrays/s", "n_gpus": world,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": total_ms / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": workload_config(wl),
        "step_definition": f"one step = {VPS} views (one k_trace_forward launch each), one view per launch per GPU; "
                           f"view-sharded x{world}, primitives replicated",
        "rays_per_step_per_gpu": R * VPS, "mean_hits_per_ray": round(mean_hits, 2),
        "clocks": clock_summary,
        "gpu_launches": n_launch,  # one k_trace_forward launch per view in the timed region (rays generated in-kernel)
        "e2e": {"value": e2e_value, "unit": "Mrays/s", "h2d_bytes_per_step": cam_bytes * VPS,
                "d2h_bytes_per_step": R * 12 * VPS, "ms_per_step": e2e_ms / args.steps,
                "api": "volprim_balance_b200.render_to_host(scene, sensors=[views], out=<2 pinned images>): per view the sensor "
                       "(76 B) host->device, trace with in-kernel ray generation, image device->host; the copy of view i overlaps the "
                       "trace of view i+1"},
        "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                     "traffic": None, "kernel": "k_trace_forward<RF,%s,SH3,tile>" % wl.get("kernel", "gaussian").upper(),
                     "algorithmic_bytes_per_launch": sum(algo_bytes) / len(algo_bytes),
                     "kernel_ms": sum(kern_ms) / len(kern_ms), "peak_source": peak_src, "source_stamp": source_stamp()},
        "primitive_evals_per_s": mean_hits * R * n_launch * world / (total_ms * 1e-3),
        "fwd_adjoint_ms_per_view": fa["fwd_adjoint_ms_per_view"],
        "adjoint": {"formulation": "gather: counting pass over the hit record (k_bucket_ranks_dense), ray-major replay into "
                                   "per-primitive buckets (k_adjoint_rows_dense), one warp per primitive (k_adjoint_gather); no "
                                   "global float reductions", **fa,
                    "byte_model": f"{ADJ_HIT_BYTES} B/recorded hit + {ADJ_PRIM_BYTES_SH3} B/primitive + {ADJ_RAY_BYTES} B/ray",
                    "kernel_ms": adj_ms, "algorithmic_bytes_per_launch": adj_bytes,
                    "achieved": adj_bytes / (adj_ms * 1e-3) / 1e9, "unit": "GB/s",
                    "frac": adj_bytes / (adj_ms * 1e-3) / 1e9 / peak,
                    "forward_recording_overhead": fa["forward_record_ms"] / (sum(kern_ms) / len(kern_ms)) - 1.0},
    }
    prof = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(prof) and args.workload == "cfg2":
        try:
            t = json.load(open(prof))
            # only a figure measured on THIS build counts
            if t.get("source_stamp") == line["roofline"]["source_stamp"]:
                line["roofline"]["traffic"] = t.get("k_trace_forward_dram_bytes_per_launch")
                line["roofline"]["traffic_source"] = t.get("source")
        except Exception:
            pass
    if tiles:
        line["single_view_tiles"] = tiles
    if train:
        line["train_step"] = train
    if north:
        line["north_star"] = north
    if build:
        line["build"] = build
    if cpu:
        line["cpu_baseline"] = cpu
    print(json.dumps(line), flush=True)


def run_train_step(args, torch, dist, vp, training, parallel, Ellipsoid, wl, cloud, ref_scene, dev, rank, world):
    """examples/refine_3dg_dataset.py:170-189 on a fixed batch of TRAIN_VIEWS views of the workload, sharded by view."""
    from volprim_balance_b200 import synthetic
    W, H, V = wl["W"], wl["H"], wl["views"]
    sensors = ref_scene.sensors()[:TRAIN_VIEWS]
    mine = parallel.shard_views(len(sensors), rank, world)
    targets = {i: vp.render(ref_scene, sensor=sensors[i], spp=1, jitter=False).clone() for i in mine}
    rng = __import__("numpy").random.default_rng(7)
    start = cloud.data.copy()
    start[:, :3] += rng.normal(0, 2e-3, (cloud.n, 3)).astype("float32")
    scene = make_scene(vp, wl, cloud, dev, data=start, opacities=(cloud.opacities * 0.8).clip(1e-4, 1 - 1e-4),
                       sh=cloud.sh_coeffs * 0.9)
    p = vp.traverse(scene)
    opt = vp.optimizers.BoundedAdam()
    e = Ellipsoid.unravel(p["primitives.data"])
    opt["centers"], opt["scales"], opt["quats"] = e.center, e.scale, e.quat
    opt["opacities"], opt["sh_coeffs"] = p["primitives.opacities"], p["primitives.sh_coeffs"]
    opt.set_learning_rate({"centers": 1e-4, "scales": 1e-4, "quats": 1e-4, "opacities": 1e-2, "sh_coeffs": 1e-3})
    opt.set_bounds("scales", lower=1e-6)
    opt.set_bounds("opacities", lower=1e-6, upper=1.0 - 1e-6)
    step = training.RefineStep(scene, sensors, targets, opt, n_chunks=args.train_chunks, rebuild=args.train_rebuild,
                               rebuild_every=args.train_rebuild_every)
    losses = []
    for _ in range(2):
        losses.append(float(step.step()[0]))
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    n_steps = args.train_steps
    exposed, optim = [], []
    t0.record()
    for _ in range(n_steps):
        losses.append(float(step.step()[0]))
        exposed.append(step.timing["exposed_allreduce_ms"])
        optim.append(step.timing["optimizer_and_rebuild_ms"])
    t1.record()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    ms = t0.elapsed_time(t1) / n_steps
    if world > 1:
        tmax = torch.tensor([ms], device=dev)
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
        ms = float(tmax.item())
    return {"workload": f"refine_3dg_dataset-style step (BASELINE configs[3]): {wl['n']} Gaussian primitives SH3, a fixed batch of "
                        f"{len(sensors)} views of {W}x{H} per step sharded by view over {world} GPU(s); per view forward (recording) + "
                        f"gather adjoint, L1 loss; gradient all-reduce in {len(step.ranges)} primitive ranges overlapped with the last "
                        f"view's accumulation; fused L1 loss + gradient; BoundedAdam; LBVH {args.train_rebuild}"
                        + (f" (full build every {args.train_rebuild_every} steps)" if args.train_rebuild == "refit" else ""),
            "scaling": "strong", "views_per_step": len(sensors), "views_per_gpu": len(mine), "ms_per_step": ms,
            "views_per_s": len(sensors) / (ms * 1e-3), "Mrays_per_s": len(sensors) * W * H / (ms * 1e-3) / 1e6,
            "allreduce_bytes": step.bucket.flat.numel() * 4 if world > 1 else 0, "allreduce_ranges": len(step.ranges),
            "exposed_allreduce_ms": sum(exposed) / len(exposed), "optimizer_and_rebuild_ms": sum(optim) / len(optim),
            "loss_first_last": [losses[0], losses[-1]], "timed_steps": n_steps}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="cfg2", choices=sorted(WORKLOADS))
    ap.add_argument("--views-per-step", type=int, default=0, help="views per step (default: 32 for cfg2, fewer for the larger workloads)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-train", action="store_true")
    ap.add_argument("--no-extras", action="store_true", help="skip the cfg3 (north-star) and build-time measurements")
    ap.add_argument("--warmup-full", action="store_true")
    ap.add_argument("--train-steps", type=int, default=6)
    ap.add_argument("--train-chunks", type=int, default=4)
    ap.add_argument("--train-rebuild", default="refit", choices=["rebuild", "refit"],
                    help="LBVH after every optimiser step: full rebuild, or refit with a full build every --train-rebuild-every steps")
    ap.add_argument("--train-rebuild-every", type=int, default=8)
    args = ap.parse_args()
    wl = WORKLOADS[args.workload]
    if not args.views_per_step:
        args.views_per_step = {"cfg2": VIEWS_PER_STEP, "small": VIEWS_PER_STEP, "cfg3": 16, "cfg5": 2}[args.workload]
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, wl, rank, world)
        return
    if world > 1:
        import torch
        import torch.distributed as dist
        torch.cuda.set_device(local_rank)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    try:
        run_ours(args, wl, rank, world, local_rank)
    finally:
        if world > 1:
            import torch.distributed as dist
            dist.destroy_process_group()


if __name__ == "__main__":
    main()
