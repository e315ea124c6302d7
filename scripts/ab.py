#!/usr/bin/env python
"""A/B probe of one build of the trace kernels (VOLPRIM_CUDA_LIB=build_var/libvp_X.so python scripts/ab.py [cfg2 cfg3 ...]):
forward ms/view (best of 3 passes over the 8 views), forward with recording, gather adjoint passes, and a checksum of the
images so that variants can be checked for identical output.  One JSON line per workload."""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
import volprim_balance_b200 as vp  # noqa: E402
from volprim_balance_b200.accel import RaySource  # noqa: E402


def main():
    names = [a for a in sys.argv[1:] if not a.startswith("-")] or ["cfg2"]
    adjoint = "--no-adjoint" not in sys.argv
    explicit = "--explicit" in sys.argv      # rays from vp_raygen_perspective in HBM instead of in-kernel generation
    dev = torch.device("cuda", 0)
    for name in names:
        wl = bench.WORKLOADS[name]
        cloud = bench.build_cloud(wl)
        scene = bench.make_scene(vp, wl, cloud, dev)
        shape = scene.ellipsoids()
        shape.bind("opacities", with_sh=True)
        acc = shape.accel()
        params = scene.integrator._vp_params(scene, None)
        sens = scene.sensors()
        R = wl["W"] * wl["H"]
        V = len(sens) if name != "cfg5" else 2
        if explicit:
            srcs = []
            for v in range(V):
                o_, d_, m_ = acc.raygen_perspective(sens[v].vp_camera(), 1, None)
                srcs.append(RaySource(o=o_, d=d_, maxt=m_))
            params.image_width, params.image_height = wl["W"], wl["H"]
            fwd = lambda v: acc.render_forward(params, srcs[v], want_nhits=False)
        else:
            fwd = lambda v: acc.render_forward(params, RaySource(camera=sens[v].vp_camera()), want_nhits=False)
        for v in range(min(V, 3)):
            fwd(v)
        torch.cuda.synchronize()
        best = 1e9
        chk = 0.0
        for _ in range(3 if name != "cfg5" else 1):
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            for v in range(V):
                r = fwd(v)
            b.record()
            torch.cuda.synchronize()
            best = min(best, a.elapsed_time(b) / V)
        hits = 0
        st_sum = {}
        for v in range(V):
            r = acc.render_forward(params, RaySource(camera=sens[v].vp_camera()), want_nhits=False)
            torch.cuda.synchronize()
            st = acc.stats()
            hits += st["hits"]
            for k2, v2 in st.items():
                st_sum[k2] = st_sum.get(k2, 0) + v2
            chk += float(r.rgb.double().sum())
        out = {"lib": os.path.basename(vp._cabi.LIB_PATH), "workload": name, "rays": "explicit" if explicit else "in-kernel", "forward_ms_per_view": round(best, 4),
               "hits_per_ray": round(hits / V / R, 2), "checksum": chk,
               "per_ray": {k2: round(v2 / V / R, 2) for k2, v2 in st_sum.items() if k2 != "rays"},
               "roofline_frac": round((R * 44 + hits / V * 236) / (best * 1e-3) / 1e9 / bench.measured_peak()[0], 4)}
        if adjoint:
            acc.hits_per_ray_estimate = hits / V / R * 1.08
            id_cap = 128 if wl.get("max_depth", 128) > 0 else 1024
            fa = bench.time_fwd_adjoint(torch, acc, params, sens, list(range(V)), id_cap, R, cloud.n, 48, reps=V)
            if "--buckets" in sys.argv:
                rec = acc.new_record(R, id_cap, with_state=False)
                acc.render_forward(params, RaySource(camera=sens[0].vp_camera()), record=rec, id_cap=id_cap)
                tot = rec.totals()[0]
                cnt = torch.bincount(rec.ids[:tot].long(), minlength=cloud.n)
                edges = [0, 1, 9, 33, 65, 257, 1025, 1 << 30]
                out["buckets"] = {f"{a}..{b - 1}": [round(float(((cnt >= a) & (cnt < b)).float().mean()), 4),
                                                     round(float(cnt[(cnt >= a) & (cnt < b)].sum()) / tot, 4)] for a, b in zip(edges, edges[1:])}
            out.update({k: round(v, 4) if isinstance(v, float) else v for k, v in fa.items()})
        print(json.dumps(out), flush=True)
        del scene, shape, acc, cloud
        torch.cuda.empty_cache()


if __name__ == "__main__":
    main()
