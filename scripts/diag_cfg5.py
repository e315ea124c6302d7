#!/usr/bin/env python
"""Diagnostic: robust rays of the cfg5 subsample whose GPU hit list differs from the oracle's."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
import volprim_balance_b200 as vp  # noqa: E402
from volprim_balance_b200 import synthetic  # noqa: E402
from volprim_balance_b200.accel import RaySource  # noqa: E402
from tests.parity_utils import gpu_scene, make_params, oracle_scene, record_lists, robust_mask  # noqa: E402
from tests.test_gpu_at_size import _sensor, _strided  # noqa: E402

name = sys.argv[1] if len(sys.argv) > 1 else "cfg5"
wl = bench.WORKLOADS[name]
cloud = bench.build_cloud(wl)
W, H = wl["W"], wl["H"]
cam = synthetic.ring_camera(0, wl["views"], W, H)
acc = gpu_scene(cloud)
acc.hits_per_ray_estimate = 162.0 if name == "cfg5" else 56.0
cap = 1024 if name == "cfg5" else 128
p, op = make_params(0, 1 if wl.get("kernel") == "epanechnikov" else 0, wl.get("max_depth", 128))
sensor = _sensor(cam)
rays = RaySource(camera=sensor.vp_camera())
fwd = acc.render_forward(p, rays, record=True, id_cap=cap)
print("record", fwd.record.totals(), "capacity", fwd.record.capacity, "usable", fwd.record.usable())
o, d, mt = (x.cpu().numpy() for x in acc.raygen_perspective(sensor.vp_camera(), 1, None))
sel = np.flatnonzero(_strided(16)(W, H).reshape(-1))
osc = oracle_scene(cloud)
ref = osc.forward(op, o[sel], d[sel], mt[sel], cap=cap, fragility=True)
ids_g, cnt_g = record_lists(fwd.record, sel, cap)
nh_g = fwd.nhits.cpu().numpy()[sel]
same = (ids_g == ref.hit_ids[:, :cap]).all(1) & (nh_g == ref.nhits)
rob = robust_mask(ref)
bad = np.flatnonzero(rob & ~same)
print(len(bad), "robust rays differ;", int((~same).sum()), "rays differ in total of", len(sel))
# per-ray walker on the same rays (explicit batch)
tsel = torch.from_numpy(sel).cuda()
to, td, tm = (torch.from_numpy(x[sel]) for x in (o, d, mt))
pr = acc.trace_forward(p, to, td, tm, record_cap=cap)
ids_r = pr.hit_ids.t().cpu().numpy()
same_r = (ids_r == ref.hit_ids[:, :cap]).all(1)
print("per-ray walker: robust rays that differ:", int((rob & ~same_r).sum()))
for r in bad[:8]:
    a, b = ids_g[r], ref.hit_ids[r, :cap]
    k = int(np.argmax(a != b))
    print(f"ray {r} (pixel {sel[r] % W},{sel[r] // W}) nhits gpu {nh_g[r]} orc {ref.nhits[r]} first diff at {k}: gpu {a[k:k+4]} orc {b[k:k+4]} per-ray {ids_r[r][k:k+4]}")
    print("    oracle entry t", ref.hit_t[r, max(k-1,0):k+3], "fragility", ref.fragility[r])
    # is it a swap, an insertion or a deletion?
    sa, sb = set(a[a >= 0].tolist()), set(b[b >= 0].tolist())
    print("    only gpu:", sorted(sa - sb)[:6], "only oracle:", sorted(sb - sa)[:6])
