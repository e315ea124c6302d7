"""GPU: parity AT THE SIZES BASELINE.json names -- cfg2 (1M Gaussian, 1080p) gradients where thousands of rays reduce
into one primitive, cfg3 (3M Epanechnikov, 1080p) and cfg5 (10M, 3840x2160, max_depth = -1, ~200 hits per ray) hit
lists / radiance / gradients -- against the oracle on pixel subsamples it can afford, with the contract's tolerances."""
import os
import sys

import numpy as np
import pytest
import torch

import volprim_balance_b200 as vp
from oracle import oracle as O
from volprim_balance_b200 import synthetic
from volprim_balance_b200.accel import RaySource, TraceResult
from tests.parity_utils import (check_gradients, compare_forward, f64_reference, gpu_scene, grad_close, make_params, oracle_scene,
                                record_lists)

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402  (the workload definitions of the benchmark ARE the configurations under test)

pytestmark = pytest.mark.gpu


def _sensor(cam):
    return vp.PerspectiveSensor({"type": "perspective", "fov": cam.fov_x_deg, "fov_axis": "x",
                                 "to_world": vp.Transform4f(cam.to_world), "near_clip": cam.near_clip, "far_clip": cam.far_clip,
                                 "film": {"type": "hdrfilm", "width": cam.width, "height": cam.height, "rfilter": {"type": "box"}}})


def _check_at_size(name, sel_fn, id_cap, hits_estimate, max_fragile_frac=5e-3, grad=True):
    wl = bench.WORKLOADS[name]
    cloud = bench.build_cloud(wl)
    W, H = wl["W"], wl["H"]
    cam = synthetic.ring_camera(0, wl["views"], W, H)
    acc = gpu_scene(cloud)
    acc.hits_per_ray_estimate = hits_estimate
    kernel = 1 if wl.get("kernel") == "epanechnikov" else 0
    p, op = make_params(0, kernel, wl.get("max_depth", 128))
    sensor = _sensor(cam)
    rays = RaySource(camera=sensor.vp_camera())
    fwd = acc.render_forward(p, rays, record=True, id_cap=id_cap)
    rec = fwd.record
    entries, cut = rec.totals()
    assert rec.usable(), f"record unusable: {entries} entries for capacity {rec.capacity}, {cut} rays cut at {id_cap}"
    nh = fwd.nhits.cpu().numpy().astype(np.int64)
    assert entries == int(nh.sum())                       # kept bytes per view = 4 x sum of the hit counts (+ offsets)
    st = acc.stats()
    assert st["stack_overflows"] == 0
    o, d, mt = (x.cpu().numpy() for x in acc.raygen_perspective(sensor.vp_camera(), 1, None))
    sel = np.flatnonzero(sel_fn(W, H).reshape(-1))
    osc = oracle_scene(cloud)
    ref = osc.forward(op, o[sel], d[sel], mt[sel], cap=id_cap, fragility=True)
    ids_g, cnt_g = record_lists(rec, sel, id_cap)
    tsel = torch.from_numpy(sel).cuda()
    part = TraceResult(fwd.rgb[tsel], fwd.beta[tsel], fwd.nhits[tsel])
    ref64 = oracle_scene(cloud, precision="f64").forward(op, o[sel], d[sel], mt[sel], cap=id_cap)
    out = compare_forward(part, ref, id_cap, max_fragile_frac=max_fragile_frac, replay=(osc, op, o[sel], d[sel], mt[sel]), ids_g=ids_g,
                          res_orc64=ref64)
    del ref64
    info = {k: v for k, v in out.items() if k[0] != "_"}
    info.update(primitives=cloud.n, rays=W * H, hits_per_ray=float(nh.mean()), max_hits=int(nh.max()), record_GB=rec.nbytes() / 2**30,
                entries=entries)
    if grad:
        dL = np.zeros((W * H, 3), np.float32)
        dsel = np.random.default_rng(7).normal(size=(len(sel), 3)).astype(np.float32)
        del osc
        osc64, same64 = f64_reference(cloud, op, o[sel], d[sel], mt[sel], ids_g, id_cap)
        dsel[~(out["_same"] & same64)] = 0
        dL[sel] = dsel
        state = fwd.rgb.clone()
        state[tsel] = torch.from_numpy(ref.rgb).cuda()
        gd, ga, gs = acc.render_adjoint(p, rays, torch.from_numpy(dL), state, rec)
        want, noise = osc64.adjoint(op, o[sel], d[sel], dsel, ref.rgb, mt[sel])
        info["grad_err"] = check_gradients((gd, ga, gs), want, noise, name)
        per_prim = np.bincount(ids_g[ids_g >= 0], minlength=cloud.n)
        info["max_selected_rays_per_primitive"] = int(per_prim.max())
    print(name, info)
    return info


def _strided(step):
    def f(W, H):
        m = np.zeros((H, W), bool)
        m[step // 2::step, step // 2::step] = True
        return m
    return f


def test_cfg2_adjoint_at_size_where_thousands_of_rays_meet_one_primitive():
    """1M Gaussian primitives, 1920x1080.  delta-L is non-zero on a dense 256x160 pixel crop (plus every 16th pixel of
    the rest): neighbouring rays hit the same primitives, so single primitives collect hundreds to thousands of per-hit
    terms -- the regime the gather adjoint exists for.  All three gradient blocks against the oracle, elementwise."""
    def sel(W, H):
        m = _strided(16)(W, H)
        m[460:620, 832:1088] = True
        return m
    info = _check_at_size("cfg2", sel, id_cap=128, hits_estimate=40.0)
    assert 25 < info["hits_per_ray"] < 40 and info["max_selected_rays_per_primitive"] > 300


def test_cfg3_3m_epanechnikov_1080p_tile_walker_against_oracle():
    """BASELINE configs[2] / the north-star target configuration: 3M Epanechnikov primitives at 1080p."""
    info = _check_at_size("cfg3", _strided(16), id_cap=128, hits_estimate=56.0)
    assert info["primitives"] == 3_000_000 and 35 < info["hits_per_ray"] < 55


def test_cfg5_10m_4k_unbounded_depth_against_oracle():
    """BASELINE configs[4]: 10M overlapping primitives, 3840x2160, max_depth = -1, ~200 hits per ray.  The record keeps
    4 bytes per hit (< 7 GB for the view); the dense scratch it is compacted from is bounded by row bands."""
    info = _check_at_size("cfg5", _strided(16), id_cap=1024, hits_estimate=215.0, max_fragile_frac=2e-2)
    assert info["primitives"] == 10_000_000 and 150 < info["hits_per_ray"] < 260 and info["max_hits"] > 400
    assert info["record_GB"] < 9.5 and info["entries"] * 4 / 2**30 < 7.0
