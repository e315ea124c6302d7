#!/usr/bin/env python
"""Diagnostic: one cfg5 ray whose GPU list misses a primitive -- reduce the scene to the primitives of the oracle's list."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from oracle import oracle as O  # noqa: E402
from volprim_balance_b200 import synthetic  # noqa: E402
from tests.parity_utils import gpu_scene, make_params, oracle_scene  # noqa: E402

px, py = int(sys.argv[1]), int(sys.argv[2])
wl = bench.WORKLOADS["cfg5"]
cloud = bench.build_cloud(wl)
W, H = wl["W"], wl["H"]
cam = synthetic.ring_camera(0, wl["views"], W, H)
acc = gpu_scene(cloud)
from tests.test_gpu_at_size import _sensor
o, d, mt = (x.cpu().numpy() for x in acc.raygen_perspective(_sensor(cam).vp_camera(), 1, None))
r = py * W + px
o, d, mt = o[r:r + 1], d[r:r + 1], mt[r:r + 1]
p, op = make_params(0, 0, -1)
osc = oracle_scene(cloud)
ref = osc.forward(op, o, d, mt, cap=1024, fragility=True)
ids = ref.hit_ids[0][ref.hit_ids[0] >= 0]
g = acc.trace_forward(p, torch.from_numpy(o), torch.from_numpy(d), torch.from_numpy(mt), record_cap=1024)
gi = g.hit_ids[:, 0].cpu().numpy()
gi = gi[gi >= 0]
k = int(np.argmax(gi[:len(ids)] != ids[:len(gi)]))
print("full scene: oracle", len(ids), "hits, gpu", len(gi), "first diff", k, "oracle", ids[k:k + 3], "gpu", gi[k:k + 3], "stats", acc.stats())
missing = ids[k]
print("missing primitive", missing, cloud.data[missing], "entry t", ref.hit_t[0, k - 1:k + 2])
for name, sub in (("oracle list only", ids), ("list window", ids[max(k - 3, 0):k + 4]), ("pair", ids[k:k + 2]), ("single", ids[k:k + 1])):
    sub = np.asarray(sub)
    c2 = synthetic.Cloud(cloud.data[sub].copy(), cloud.opacities[sub].copy(), cloud.sh_coeffs[sub].copy(), 3.0)
    a2 = gpu_scene(c2)
    g2 = a2.trace_forward(p, torch.from_numpy(o), torch.from_numpy(d), torch.from_numpy(mt), record_cap=1024)
    l2 = g2.hit_ids[:, 0].cpu().numpy()
    l2 = sub[l2[l2 >= 0]]
    r2 = oracle_scene(c2).forward(op, o, d, mt, cap=1024)
    lo = sub[r2.hit_ids[0][r2.hit_ids[0] >= 0]]
    print(f"{name}: gpu finds missing: {missing in l2}, oracle finds it: {missing in lo}; gpu {len(l2)} hits, oracle {len(lo)}")
# the primitive seen from the origin the loop has when it meets it
e = cloud.data[missing]
print("ray_ellipsoid from the ORIGINAL origin:", O.ray_ellipsoid(o[0], d[0], e, 3.0), "f64:", O.ray_ellipsoid(o[0], d[0], e, 3.0, precision="f64"))
