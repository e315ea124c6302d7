#!/usr/bin/env python
"""Diagnostic (not a test): where do the largest GPU-vs-oracle gradient discrepancies come from?  For the worst element of
the centre-gradient block, every ray that hits that primitive is run alone through the GPU adjoint and both oracles."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import oracle as O  # noqa: E402
from volprim_balance_b200 import synthetic  # noqa: E402
from tests.parity_utils import gpu_scene, make_params  # noqa: E402


def ratio(a, b):
    a, b = np.asarray(a, np.float64).ravel(), np.asarray(b, np.float64).ravel()
    rms = np.sqrt(np.mean(b ** 2))
    return np.abs(a - b) / (np.abs(b) + rms), rms


n = 6000
cloud = synthetic.make_cloud(n, synthetic.sigma0_for_hits(n, 35), seed=1, sh_degree=3)
o, d, mt = synthetic.camera_rays(synthetic.ring_camera(1, 8, 64, 32))
acc = gpu_scene(cloud)
p, op = make_params(0, 0, 128)
to, td, tm = (torch.from_numpy(x) for x in (o, d, mt))
fwd = acc.trace_forward(p, to, td, tm, record_cap=128)
s32 = O.Scene(cloud.data, cloud.opacities, cloud.sh_coeffs, 3.0, precision="f32")
s64 = O.Scene(cloud.data, cloud.opacities, cloud.sh_coeffs, 3.0, precision="f64")
ref = s32.forward(op, o, d, mt, cap=128)
ref64 = s64.forward(op, o, d, mt, cap=128)
ids_g = fwd.hit_ids.t().cpu().numpy()
same = (ids_g == ref.hit_ids).all(1) & (ref.hit_ids == ref64.hit_ids).all(1)
dL = np.random.default_rng(7).normal(size=(o.shape[0], 3)).astype(np.float32)
dL[~same] = 0
state = ref.rgb
g = acc.trace_adjoint(p, to, td, tm, torch.from_numpy(dL), torch.from_numpy(state), hit_ids=fwd.hit_ids, hit_counts=fwd.nhits)
g32 = s32.adjoint(op, o, d, dL, state, mt)
g64 = s64.adjoint(op, o, d, dL, state.astype(np.float64), mt)
gd = g[0].cpu().numpy().reshape(-1, 10)
for name, sl in (("center", slice(0, 3)), ("scale", slice(3, 6)), ("quat", slice(6, 10))):
    r32, rms = ratio(gd[:, sl], g32[0][:, sl])
    r64, _ = ratio(gd[:, sl], g64[0][:, sl])
    ro, _ = ratio(g32[0][:, sl], g64[0][:, sl])
    print(f"{name}: gpu-vs-orc32 {r32.max():.2e}  gpu-vs-orc64 {r64.max():.2e}  orc32-vs-orc64 {ro.max():.2e}  rms {rms:.3e}")
r32, rms = ratio(gd[:, 0:3], g32[0][:, 0:3])
worst = int(r32.argmax())
prim, comp = worst // 3, worst % 3
print("worst element: primitive", prim, "component", comp, "gpu", gd[prim, comp], "orc32", g32[0][prim, comp], "orc64", g64[0][prim, comp])
print("primitive record", cloud.data[prim], "opacity", cloud.opacities[prim])
rays = np.flatnonzero((ids_g == prim).any(1) & same)
print(len(rays), "rays hit it")
rows = []
for r in rays:
    one = np.zeros_like(dL)
    one[r] = dL[r]
    sub = slice(r, r + 1)
    gg = acc.trace_adjoint(p, to[sub], td[sub], tm[sub], torch.from_numpy(one[sub]), torch.from_numpy(state[sub]),
                           hit_ids=fwd.hit_ids[:, sub].contiguous(), hit_counts=fwd.nhits[sub])
    a = gg[0].cpu().numpy().reshape(-1, 10)[prim, comp]
    b = s32.adjoint(op, o[sub], d[sub], one[sub], state[sub], mt[sub])[0][prim, comp]
    c = s64.adjoint(op, o[sub], d[sub], one[sub], state[sub].astype(np.float64), mt[sub])[0][prim, comp]
    k = int(np.flatnonzero(ids_g[r] == prim)[0])
    rows.append((abs(a - b), r, k, a, b, c))
rows.sort(reverse=True)
print("largest per-ray discrepancies (|gpu - orc32|, ray, hit index, gpu, orc32, orc64):")
for row in rows[:8]:
    print("  %.3e ray %d hit %d gpu %.6e orc32 %.6e orc64 %.6e" % row)
print("sum of per-ray: gpu", sum(r[3] for r in rows), "orc32", sum(r[4] for r in rows), "orc64", sum(r[5] for r in rows))
