"""Per-workload walk statistics (passes, candidates, node tests per ray) for the bench workloads."""
import sys, os, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import bench
from volprim_balance_b200 import synthetic, _cabi
from volprim_balance_b200.accel import EllipsoidAccel
name = sys.argv[1]
wl = bench.WORKLOADS[name]
cloud = bench.build_cloud(wl)
W, H = wl["W"], wl["H"]
acc = EllipsoidAccel()
acc.set_primitives(torch.from_numpy(cloud.data), torch.from_numpy(cloud.opacities), torch.from_numpy(cloud.sh_coeffs), 3.0)
acc.build()
p = _cabi.vp_params(); p.integrator = 0; p.kernel = 1 if wl.get("kernel") == "epanechnikov" else 0
md = wl.get("max_depth", 128); p.max_depth = 0xFFFFFFFF if md < 0 else md
p.srgb_primitives = 1; p.t_cutoff = 0.01; p.eps_advance = 1e-4; p.image_width = W; p.image_height = H
for view in (0, 1):
    o, d, mt = synthetic.camera_rays(synthetic.ring_camera(view, 8, W, H))
    o, d, mt = torch.from_numpy(o).cuda(), torch.from_numpy(d).cuda(), torch.from_numpy(mt).cuda()
    ms = 1e9
    for rep in range(3):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); res = acc.trace_forward(p, o, d, mt); e1.record(); torch.cuda.synchronize()
        ms = min(ms, e0.elapsed_time(e1))
    st = acc.stats(); R = W * H
    print(json.dumps({"workload": name, "view": view, "ms": round(ms, 3), "hits": round(st["hits"] / R, 2), "cands": round(st["candidates"] / R, 1),
                      "nodes": round(st["node_visits"] / R, 1), "passes": round(st["passes"] / R, 2), "overflow": st["stack_overflows"], "retries": st["interval_retries"]}))
