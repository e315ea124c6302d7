"""Quick GPU probe: build + forward timing on a cfg2-like cloud (not a benchmark; see bench.py)."""
import sys, time, json, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from volprim_balance_b200 import synthetic, _cabi
from volprim_balance_b200.accel import EllipsoidAccel

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
cross = float(sys.argv[2]) if len(sys.argv) > 2 else 60
W, H = 1920, 1080
cloud = synthetic.make_cloud(n, synthetic.sigma0_for_hits(n, cross), seed=1)
acc = EllipsoidAccel()
acc.set_primitives(torch.from_numpy(cloud.data), torch.from_numpy(cloud.opacities), torch.from_numpy(cloud.sh_coeffs), 3.0)
torch.cuda.synchronize(); t = time.time(); acc.build(); torch.cuda.synchronize(); print("build s", time.time() - t)
t = time.time(); acc.build(); torch.cuda.synchronize(); print("build2 s", time.time() - t)
p = _cabi.vp_params(); p.integrator = 0; p.kernel = 0; p.max_depth = 128; p.srgb_primitives = 1
p.t_cutoff = 0.01; p.eps_advance = 1e-4; p.image_width = W; p.image_height = H
for view in range(3):
    cam = synthetic.ring_camera(view, 8, W, H)
    o, d, mt = synthetic.camera_rays(cam)
    o, d, mt = torch.from_numpy(o).cuda(), torch.from_numpy(d).cuda(), torch.from_numpy(mt).cuda()
    for rec in (0, 128):
        ms = 1e9
        for rep in range(3):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); res = acc.trace_forward(p, o, d, mt, record_cap=rec); e1.record(); torch.cuda.synchronize()
            ms = min(ms, e0.elapsed_time(e1))
        st = acc.stats()
        print(json.dumps({"view": view, "record": rec, "ms": ms, "mrays_s": W * H / ms / 1e3, "mean_hits": st["hits"] / (W * H),
                          "cand_per_ray": st["candidates"] / (W * H), "nodes_per_ray": st["node_visits"] / (W * H),
                          "passes_per_ray": st["passes"] / (W * H), "overflow": st["stack_overflows"], "retries": st["interval_retries"],
                          "beta_mean": float(res.beta.mean())}))
    if view == 0:
        dL = torch.randn(W * H, 3, device="cuda")
        for replay in (True, False):
            ms = 1e9
            for rep in range(2):
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                g = acc.trace_adjoint(p, o, d, mt, dL, res.rgb, res.hit_ids if replay else None, res.nhits if replay else None)
                e1.record(); torch.cuda.synchronize()
                ms = min(ms, e0.elapsed_time(e1))
            print(json.dumps({"adjoint_replay": replay, "ms": ms}))
