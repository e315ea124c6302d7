"""Calibrate sigma0 of the synthetic clouds so that the mean number of PROCESSED hits per ray meets the target of
BASELINE.md section 4 (+-10 %), using the GPU path at reduced resolution.  Prints the values hard-coded in bench.py."""
import sys, os, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from volprim_balance_b200 import synthetic, _cabi
from volprim_balance_b200.accel import EllipsoidAccel

def mean_hits(n, sigma0, centers, kernel, mu_o, max_depth, seed, W=480, H=272):
    cloud = synthetic.make_cloud(n, sigma0, seed=seed, sh_degree=3, centers=centers, mu_opacity=mu_o)
    acc = EllipsoidAccel()
    acc.set_primitives(torch.from_numpy(cloud.data), torch.from_numpy(cloud.opacities), torch.from_numpy(cloud.sh_coeffs), 3.0)
    acc.build()
    p = _cabi.vp_params(); p.integrator = 0; p.kernel = kernel; p.max_depth = 0xFFFFFFFF if max_depth < 0 else max_depth
    p.srgb_primitives = 1; p.t_cutoff = 0.01; p.eps_advance = 1e-4; p.image_width = W; p.image_height = H
    tot = 0.0
    for v in (0, 3):
        o, d, mt = synthetic.camera_rays(synthetic.ring_camera(v, 8, W, H))
        acc.trace_forward(p, torch.from_numpy(o).cuda(), torch.from_numpy(d).cuda(), torch.from_numpy(mt).cuda())
        tot += acc.stats()["hits"] / (W * H)
    acc.close()
    return tot / 2

for name, n, centers, kernel, mu_o, md, seed, target in (("cfg3", 3_000_000, "truck", 1, -1.0, 128, 2, 48.0),
                                                          ("cfg5", 10_000_000, "dense", 0, -4.0, -1, 4, 200.0)):
    lo, hi = 1e-4, 2e-2
    for it in range(9):
        mid = (lo * hi) ** 0.5
        h = mean_hits(n, mid, centers, kernel, mu_o, md, seed)
        print(name, "sigma0", mid, "mean hits", h, flush=True)
        if abs(h - target) / target < 0.05: break
        if h < target: lo = mid
        else: hi = mid
    print(json.dumps({"workload": name, "sigma0": mid, "mean_hits": h}), flush=True)
