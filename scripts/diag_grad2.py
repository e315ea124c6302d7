#!/usr/bin/env python
"""Diagnostic: per-ray decomposition of one at-size gradient element (gather adjoint vs oracles):
python scripts/diag_grad2.py cfg5 <primitive> <component 0..9>"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from oracle import oracle as O  # noqa: E402
from volprim_balance_b200 import synthetic  # noqa: E402
from volprim_balance_b200.accel import RaySource  # noqa: E402
from tests.parity_utils import gpu_scene, make_params, oracle_scene  # noqa: E402

name, prim, comp = sys.argv[1], int(sys.argv[2]), int(sys.argv[3])
wl = bench.WORKLOADS[name]
CAP = 128 if wl.get("max_depth", 128) > 0 else 1024
cloud = bench.build_cloud(wl)
W, H = wl["W"], wl["H"]
cam = synthetic.ring_camera(0, wl["views"], W, H)
acc = gpu_scene(cloud)
p, op = make_params(0, 1 if wl.get("kernel") == "epanechnikov" else 0, wl.get("max_depth", 128))
acc.hits_per_ray_estimate = 260.0 if name == "cfg5" else 48.0
o, d, mt = synthetic.camera_rays(cam)
m = np.zeros((H, W), bool)
m[8::16, 8::16] = True
if name == "cfg2":
    m[460:620, 832:1088] = True
sel = np.flatnonzero(m.reshape(-1))
o, d, mt = o[sel], d[sel], mt[sel]
s32, s64 = oracle_scene(cloud), oracle_scene(cloud, precision="f64")
ref = s32.forward(op, o, d, mt, cap=CAP)
rays = np.flatnonzero((ref.hit_ids == prim).any(1))
print("primitive", prim, "record", cloud.data[prim], "opacity", cloud.opacities[prim], "hit by", len(rays), "selected rays")
o, d, mt, state = o[rays], d[rays], mt[rays], ref.rgb[rays]
dL = np.random.default_rng(7).normal(size=(len(sel), 3)).astype(np.float32)[rays]
to, td, tm = (torch.from_numpy(x).cuda() for x in (o, d, mt))
src = RaySource(o=to, d=td, maxt=tm)
fwd = acc.render_forward(p, src, record=True, id_cap=CAP)
assert fwd.record.usable()
g = acc.render_adjoint(p, src, torch.from_numpy(dL), torch.from_numpy(state), fwd.record)
g32 = s32.adjoint(op, o, d, dL, state, mt)
g64 = s64.adjoint(op, o, d, dL, state, mt)
print("all rays: gpu", g[0].view(-1, 10)[prim, comp].item(), "orc32", g32[0][prim, comp], "orc64", g64[0][prim, comp])
rows = []
for k in range(len(rays)):
    sub = slice(k, k + 1)
    s1 = RaySource(o=to[sub], d=td[sub], maxt=tm[sub])
    f1 = acc.render_forward(p, s1, record=True, id_cap=CAP)
    a = acc.render_adjoint(p, s1, torch.from_numpy(dL[sub]), torch.from_numpy(state[sub]), f1.record)[0].view(-1, 10)[prim, comp].item()
    b = s32.adjoint(op, o[sub], d[sub], dL[sub], state[sub], mt[sub])[0][prim, comp]
    c = s64.adjoint(op, o[sub], d[sub], dL[sub], state[sub], mt[sub])[0][prim, comp]
    hit = int(np.flatnonzero(ref.hit_ids[rays[k]] == prim)[0])
    rows.append((abs(a - c), k, hit, a, b, c))
rows.sort(reverse=True)
print("largest per-ray |gpu - orc64| (ray, hit index, gpu, orc32, orc64):")
for r in rows[:10]:
    print("  %.3e ray %d hit %d gpu %.6e orc32 %.6e orc64 %.6e" % r)
print("sums: gpu", sum(r[3] for r in rows), "orc32", sum(r[4] for r in rows), "orc64", sum(r[5] for r in rows))
