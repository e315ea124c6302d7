"""GPU: the public host API (load_dict / render / autograd adjoint) and full-size, size-independent properties."""
import os

import numpy as np
import pytest
import torch

import volprim_balance_b200 as vp
from oracle import oracle as O
from volprim_balance_b200 import synthetic
from tests.parity_utils import (RGB_ATOL, RGB_RTOL, GradientReference, check_gradients, compare_forward, gpu_scene, grad_close,
                                make_params, oracle_scene, robust_mask)

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def _sensor_dict(cam, rfilter="box"):
    return {"type": "perspective", "fov": cam.fov_x_deg, "fov_axis": "x", "to_world": vp.Transform4f(cam.to_world),
            "near_clip": cam.near_clip, "far_clip": cam.far_clip,
            "film": {"type": "hdrfilm", "width": cam.width, "height": cam.height, "rfilter": {"type": rfilter}}}


def test_shared_reciprocal_division_is_correctly_rounded():
    """exact_isect's divisions share reciprocals (vp_div_rn); the hit ORDER rests on them being the IEEE quotients the
    oracle's C code computes.  2^28 random operand pairs of the magnitudes that occur there, against __fdiv_rn."""
    import ctypes as C
    from volprim_balance_b200 import _cabi
    lib = _cabi.load_library()
    torch.cuda.set_device(0)
    for seed in (1, 0xC0FFEE):
        bad = C.c_int64(-1)
        _cabi.check(lib.vp_debug_selftest(0, 1 << 27, seed, C.byref(bad)))
        assert bad.value == 0, f"{bad.value} of 2^27 quotients differ from __fdiv_rn"


def test_raygen_matches_perspective_sensor_restatement():
    cam = synthetic.ring_camera(3, 8, 72, 40)
    acc = gpu_scene(synthetic.make_cloud(4, 0.1, seed=0))
    s = vp.PerspectiveSensor(_sensor_dict(cam))
    o, d, mt = acc.raygen_perspective(s.vp_camera(), 1, None)
    ro, rd, rmt = synthetic.camera_rays(cam)
    np.testing.assert_allclose(o.cpu().numpy(), ro, atol=2e-6)
    np.testing.assert_allclose(d.cpu().numpy(), rd, atol=2e-6)
    np.testing.assert_allclose(mt.cpu().numpy(), rmt, rtol=1e-5)
    jit = torch.rand(72 * 40 * 2, 2)
    o2, d2, _ = acc.raygen_perspective(s.vp_camera(), 2, jit)
    px = (np.repeat(np.arange(72 * 40), 2) % 72 + jit[:, 0].numpy()) / 72
    py = (np.repeat(np.arange(72 * 40), 2) // 72 + jit[:, 1].numpy()) / 40
    ro2, rd2, _ = synthetic.rays_from_samples(cam, px, py)
    np.testing.assert_allclose(d2.cpu().numpy(), rd2, atol=2e-6)


def test_cfg1_smoke_ply_tomography_256x256_16spp():
    """BASELINE configs[0]: volprim_tomography on resources/smoke.ply, 256x256, 16 spp (4x4 stratified offsets)."""
    scene = vp.load_dict({"type": "scene", "integrator": {"type": "volprim_tomography", "max_depth": -1},
                          "primitives": {"type": "ellipsoidsmesh", "filename": os.path.join(GOLD, "smoke.ply"), "extent": 3.0},
                          "environment": {"type": "constant"}})
    shape = scene.ellipsoids()
    assert shape.count == 835
    cam = synthetic.Camera(synthetic.look_at([0, 0, 4], [0, 0, 0], [0, 1, 0]), 40.0, 256, 256)
    offs = [((i + 0.5) / 4, (j + 0.5) / 4) for j in range(4) for i in range(4)]
    o, d, mt = synthetic.camera_rays(cam, offs)
    ray = vp.Ray3f(torch.from_numpy(o).cuda(), torch.from_numpy(d).cuda(), torch.from_numpy(mt).cuda())
    scene.integrator.record_cap = 256
    L, valid, aovs, state = scene.integrator.sample(vp.ADMode.Primal, scene, None, ray, record=True)
    res = scene.integrator.last
    data = shape.data.cpu().numpy().reshape(-1, 10)
    sig = shape.attributes["sigma_t"].cpu().numpy()
    ref = O.Scene(data, sig, None, 3.0).forward(O.Params(integrator=O.TOMO, max_depth=-1), o, d, mt, cap=256, fragility=True)
    st = compare_forward(res, ref, 256)
    assert st["rays"] == 256 * 256 * 16 and ref.nhits.max() > 20
    img = L.reshape(256, 256, 16, 3).mean(2)
    assert 0.0 < float(img.min()) and float(img.max()) <= 1.0 + 1e-6
    print(st)


def test_render_autograd_matches_oracle_adjoint_reference_exact_and_corrected():
    n = 3000
    cloud = synthetic.make_cloud(n, synthetic.sigma0_for_hits(n, 25), seed=21, sh_degree=2)
    cam = synthetic.ring_camera(5, 8, 48, 32)
    scene = vp.load_dict({"type": "scene", "integrator": {"type": "volprim_rf", "max_depth": 64, "rr_depth": 64},
                          "primitives": {"type": "ellipsoidsmesh", "centers": cloud.data[:, :3], "scales": cloud.data[:, 3:6],
                                         "quaternions": cloud.data[:, 6:], "opacities": cloud.opacities[:, None],
                                         "sh_coeffs": cloud.sh_coeffs, "extent": 3.0},
                          "cam": _sensor_dict(cam)})
    params = vp.traverse(scene)
    keys = ["primitives.data", "primitives.opacities", "primitives.sh_coeffs"]
    o, d, mt = synthetic.camera_rays(cam)
    osc = oracle_scene(cloud)
    gref = GradientReference(cloud)
    w_np = np.random.default_rng(2).normal(size=(32 * 48, 3)).astype(np.float32)
    # gradients are compared elementwise at 1e-3: rays whose list may legally differ in fp32 carry no weight
    w_np[~robust_mask(osc.forward(O.Params(integrator=O.RF, kernel=O.GAUSS, max_depth=64), o, d, mt, cap=64, fragility=True))] = 0
    w = torch.from_numpy(w_np.reshape(32, 48, 3)).cuda()
    for mode in ("reference_exact", "corrected"):
        for k in keys:
            params[k].requires_grad_(True)
            params[k].grad = None
        img = vp.render(scene, params, sensor=0, spp=1, jitter=False, adjoint_mode=mode)
        (img * w).sum().backward()
        op = O.Params(integrator=O.RF, kernel=O.GAUSS, max_depth=64, srgb_primitives=(mode == "reference_exact"))
        ref = osc.forward(op, o, d, mt, cap=64, fragility=True)
        dL = w.reshape(-1, 3).cpu().numpy()
        if mode == "reference_exact":   # state_in = linear output, delta-L applied in sRGB space (quirk Q3)
            got = img.detach().reshape(-1, 3).cpu().numpy()
            rob = robust_mask(ref)
            assert (np.abs(got - ref.rgb) <= RGB_ATOL + RGB_RTOL * np.abs(ref.rgb))[rob].all()
            want, noise = gref.adjoint(op, o, d, dL, ref.rgb, mt)
        else:                           # true gradient: chain through srgb_to_linear, state_in in sRGB space
            deriv = O.srgb_to_linear_deriv(ref.rgb)
            want, noise = gref.adjoint(op, o, d, dL * deriv, ref.rgb, mt)
        print(mode, check_gradients([params[k].grad for k in keys], want, noise, mode))


def test_params_update_rebuild_and_refit_agree():
    n = 20000
    cloud = synthetic.make_cloud(n, synthetic.sigma0_for_hits(n, 30), seed=8)
    o, d, mt = synthetic.camera_rays(synthetic.ring_camera(0, 8, 64, 32))
    p, _ = make_params(0, 0, 64)
    acc = gpu_scene(cloud)
    moved = cloud.data.copy()
    moved[:, :3] += np.random.default_rng(0).normal(0, 2e-3, (n, 3)).astype(np.float32)
    to, td, tm = torch.from_numpy(o), torch.from_numpy(d), torch.from_numpy(mt)
    acc.set_primitives(torch.from_numpy(moved), torch.from_numpy(cloud.opacities), torch.from_numpy(cloud.sh_coeffs), 3.0)
    acc.refit()
    a = acc.trace_forward(p, to, td, tm, record_cap=64)
    acc.build()
    b = acc.trace_forward(p, to, td, tm, record_cap=64)
    assert torch.equal(a.hit_ids, b.hit_ids) and torch.equal(a.rgb, b.rgb)


@pytest.fixture(scope="module")
def cfg2():
    n = 1_000_000
    cloud = synthetic.make_cloud(n, synthetic.sigma0_for_hits(n, 60.0), seed=1)
    cam = synthetic.ring_camera(0, 8, 1920, 1080)
    o, d, mt = synthetic.camera_rays(cam)
    return cloud, tuple(torch.from_numpy(x).cuda() for x in (o, d, mt)), (o, d, mt)


def test_full_size_cfg2_properties(cfg2):
    """1M primitives, 1920x1080: determinism, permutation invariance of the primitive order (different Morton
    ties / BVH input order, same image and same hit lists after relabelling), max_depth prefix property, replayed
    vs re-traced adjoint, and a 1/256 pixel subsample against the oracle."""
    cloud, (o, d, mt), (on, dn, mtn) = cfg2
    p, op = make_params(0, 0, 128, image=(1920, 1080))
    acc = gpu_scene(cloud)
    a = acc.trace_forward(p, o, d, mt, record_cap=128)
    b = acc.trace_forward(p, o, d, mt, record_cap=128)
    assert torch.equal(a.rgb, b.rgb) and torch.equal(a.hit_ids, b.hit_ids)            # bit-stable
    st = acc.stats()
    assert st["stack_overflows"] == 0 and 25 < st["hits"] / o.shape[0] < 40
    # permutation of the primitive numbering
    perm = np.random.default_rng(5).permutation(cloud.n)
    acc2 = vp.accel.EllipsoidAccel()
    acc2.set_primitives(torch.from_numpy(cloud.data[perm]), torch.from_numpy(cloud.opacities[perm]),
                        torch.from_numpy(cloud.sh_coeffs[perm]), 3.0)
    acc2.build()
    c = acc2.trace_forward(p, o, d, mt, record_cap=128)
    tperm = torch.from_numpy(perm).cuda()
    relabelled = torch.where(c.hit_ids >= 0, tperm[c.hit_ids.clamp_min(0).long()].int(), c.hit_ids)
    same = (relabelled == a.hit_ids).all(0)
    assert float(same.float().mean()) > 0.9999                                        # exact-tie order may differ
    assert torch.equal(c.rgb[same], a.rgb[same])
    # max_depth prefix property
    p8, _ = make_params(0, 0, 8, image=(1920, 1080))
    e = acc.trace_forward(p8, o, d, mt, record_cap=8)
    assert torch.equal(e.hit_ids, a.hit_ids[:8]) and torch.equal(e.nhits, a.nhits.clamp_max(8))
    # replayed and re-traced adjoint see the same hit sequence
    sub = slice(0, 1920 * 64)
    dL = torch.randn(1920 * 64, 3, device="cuda")
    pa, _ = make_params(0, 0, 128, image=(1920, 64))   # same (tile) walker as the recorded lists
    g1 = acc.trace_adjoint(pa, o[sub], d[sub], mt[sub], dL, a.rgb[sub], a.hit_ids[:, sub].contiguous(), a.nhits[sub])
    g2 = acc.trace_adjoint(pa, o[sub], d[sub], mt[sub], dL, a.rgb[sub])
    for x, y in zip(g1, g2):
        grad_close(x.cpu().numpy(), y.cpu().numpy(), rtol=1e-3, what="replay vs retrace")
    # oracle on every 16th pixel in x and y
    sel = np.zeros((1080, 1920), bool)
    sel[8::16, 8::16] = True
    sel = sel.reshape(-1)
    ref = oracle_scene(cloud).forward(op, on[sel], dn[sel], mtn[sel], cap=128, fragility=True)
    tsel = torch.from_numpy(sel).cuda()
    from volprim_balance_b200.accel import TraceResult
    part = TraceResult(a.rgb[tsel], a.beta[tsel], a.nhits[tsel], a.hit_ids[:, tsel].contiguous())
    print(compare_forward(part, ref, 128))


def test_fused_bounded_adam_matches_torch_reference_rule():
    """csrc/vp_optim.cu against the torch restatement of BoundedAdam.step (optimizers.py:72-146) run on the CPU."""
    from volprim_balance_b200 import optimizers
    torch.manual_seed(3)
    n = 100_003
    p0 = torch.rand(n) * 0.2 + 1e-3
    a, b = optimizers.BoundedAdam(lr=0.05), optimizers.BoundedAdam(lr=0.05)
    a["x"], b["x"] = p0.cuda(), p0.clone()
    for o in (a, b):
        o.set_bounds("x", lower=1e-6, upper=0.25)
    for step in range(4):
        g = torch.randn(n)
        g[::97] = float("nan")
        g[5::13] *= 30.0                       # large steps -> both bounds are hit
        a["x"].grad, b["x"].grad = g.cuda(), g.clone()
        a.step()
        b.step()
        # elements whose step lands within rounding of a bound may take the other branch (and reset their moments):
        # compare with fp32 op-order tolerance and allow a handful of such flips
        for got, want in ((a["x"].detach().cpu(), b["x"].detach()), (a.state["x"][0].cpu(), b.state["x"][0]),
                          (a.state["x"][1].cpu(), b.state["x"][1])):
            bad = ~torch.isclose(got, want, rtol=1e-5, atol=1e-6)
            assert int(bad.sum()) <= 20, f"{int(bad.sum())} elements differ"
    x = a["x"].detach()
    assert float(x.min()) >= 1e-6 and float(x.max()) <= 0.25 and bool(((a.state["x"][0] == 0).sum() > 0))


def test_batch_sensor_tent_filter_and_tomography_autograd():
    """Batch sensor (views side by side, refine_3dg_dataset.py:96-107) with jittered samples + tent filter, and the
    tomography plugin through render() / autograd against the oracle adjoint on pixel-centre rays."""
    n = 2000
    cloud = synthetic.make_cloud(n, synthetic.sigma0_for_hits(n, 15), seed=31, sh_degree=0)
    sig = np.random.default_rng(1).uniform(0.0005, 0.01, n).astype(np.float32)
    cams = [synthetic.ring_camera(i, 8, 40, 24) for i in (0, 3, 5)]
    sd = {"type": "scene", "integrator": {"type": "volprim_tomography", "max_depth": -1},
          "primitives": {"type": "ellipsoidsmesh", "centers": cloud.data[:, :3], "scales": cloud.data[:, 3:6],
                         "quaternions": cloud.data[:, 6:], "sigma_t": sig[:, None], "albedo": np.ones((n, 3), np.float32),
                         "extent": 3.0},
          "environment": {"type": "constant", "radiance": 0.8}}
    scene = vp.load_dict(sd)
    batch = vp.load_dict({"type": "batch", "film": {"type": "hdrfilm", "width": 120, "height": 24, "filter": {"type": "tent"}},
                          **{f"cam_{i}": _sensor_dict(c, "tent") for i, c in enumerate(cams)}})
    img = vp.render(scene, sensor=batch, spp=4, seed=3)
    assert img.shape == (24, 120, 3) and bool(torch.isfinite(img).all())
    singles = [vp.render(scene, sensor=vp.load_dict(_sensor_dict(c, "box")), spp=1, jitter=False) for c in cams]
    ref_strip = vp.utils.concatenate_tensors(singles)
    # a jittered, tent-filtered estimate of the same (high-frequency, coarse) picture: compare the view means
    for v in range(3):
        a, b = img[:, 40 * v:40 * (v + 1)].mean(), ref_strip[:, 40 * v:40 * (v + 1)].mean()
        assert abs(float(a - b)) < 0.03
    # autograd through the tomography adjoint
    params = vp.traverse(scene)
    for k in ("primitives.data", "primitives.sigma_t"):
        params[k].requires_grad_(True)
    s0 = vp.load_dict(_sensor_dict(cams[1], "box"))
    im = vp.render(scene, params, sensor=s0, spp=1, jitter=False)
    o, d, mt = synthetic.camera_rays(cams[1])
    osc = O.Scene(cloud.data, sig, None, 3.0)
    gref = GradientReference(cloud, attr=sig, sh=False)
    op = O.Params(integrator=O.TOMO, kernel=O.GAUSS, max_depth=-1, env=(0.8, 0.8, 0.8))
    ref = osc.forward(op, o, d, mt, cap=256, fragility=True)
    w_np = np.random.default_rng(5).normal(size=(24 * 40, 3)).astype(np.float32)
    w_np[~robust_mask(ref)] = 0
    w = torch.from_numpy(w_np.reshape(24, 40, 3)).cuda()
    (im * w).sum().backward()
    got = im.detach().reshape(-1, 3).cpu().numpy()
    assert (np.abs(got - ref.rgb) <= RGB_ATOL + RGB_RTOL * np.abs(ref.rgb))[robust_mask(ref)].all()
    want, noise = gref.adjoint(op, o, d, w.reshape(-1, 3).cpu().numpy(), ref.rgb, mt)
    print(check_gradients((params["primitives.data"].grad, params["primitives.sigma_t"].grad, None), want, noise, "tomography autograd"))


def test_explicit_ray_batches_nonunit_directions_and_finite_maxt():
    """The radiance-cache calling convention (scripts/radiosity/radiance_cache.py:252-266): arbitrary rays, not from a
    sensor -- random origins on a sphere, un-normalised directions, finite maxt that ends some rays inside the cloud."""
    n = 30000
    cloud = synthetic.make_cloud(n, synthetic.sigma0_for_hits(n, 40), seed=41, sh_degree=3)
    rng = np.random.default_rng(9)
    R = 4096
    o = rng.normal(size=(R, 3))
    o = (2.5 * o / np.linalg.norm(o, axis=1, keepdims=True)).astype(np.float32)
    tgt = rng.uniform(-0.8, 0.8, size=(R, 3))
    d = ((tgt - o) * rng.uniform(0.5, 2.0, size=(R, 1))).astype(np.float32)        # |d| != 1
    mt = rng.uniform(0.3, 2.0, size=R).astype(np.float32)                          # in units of |d|
    p, op = make_params(0, 0, max_depth=64)
    acc = gpu_scene(cloud)
    res = acc.trace_forward(p, torch.from_numpy(o), torch.from_numpy(d), torch.from_numpy(mt), record_cap=64)
    ref = oracle_scene(cloud).forward(op, o, d, mt, cap=64, fragility=True)
    st = compare_forward(res, ref, 64)
    assert 2 < st["mean_hits"] < 40
    print(st)


@pytest.mark.gpu
def test_render_to_host_pipelines_views_and_matches_render():
    """render_to_host(): the images that land in pinned host memory (copy of view i overlapped with the trace of view
    i+1, ring of two buffers, consumer callback) are bit-identical to render() of each view."""
    n = 3000
    cloud = synthetic.make_cloud(n, synthetic.sigma0_for_hits(n, 12), seed=77, sh_degree=1)
    cams = [synthetic.ring_camera(i, 8, 64, 32) for i in range(5)]
    sd = {"type": "scene", "integrator": {"type": "volprim_rf", "max_depth": 64},
          "primitives": {"type": "ellipsoidsmesh", "centers": cloud.data[:, :3], "scales": cloud.data[:, 3:6],
                         "quaternions": cloud.data[:, 6:], "opacities": cloud.opacities[:, None],
                         "sh_coeffs": cloud.sh_coeffs, "extent": 3.0},
          **{f"sensor_{i}": _sensor_dict(c, "box") for i, c in enumerate(cams)}}
    scene = vp.load_dict(sd)
    expected = [vp.render(scene, sensor=i, spp=1, jitter=False).cpu() for i in range(5)]
    # default: one fresh pinned image per view
    got = vp.render_to_host(scene, spp=1, jitter=False)
    assert len(got) == 5 and all(g.is_pinned() and not g.is_cuda for g in got)
    for e, g in zip(expected, got):
        assert torch.equal(e, g)
    # ring of two pinned buffers + consumer callback (the callback sees every image before its buffer is reused)
    ring = [torch.empty((32, 64, 3), dtype=torch.float32).pin_memory() for _ in range(2)]
    seen = []
    vp.render_to_host(scene, sensors=[4, 0, 3, 1, 2], out=ring, spp=1, jitter=False,
                      on_image=lambda i, h: seen.append((i, h.clone())))
    assert [i for i, _ in seen] == [0, 1, 2, 3, 4]
    for (i, h), v in zip(seen, [4, 0, 3, 1, 2]):
        assert torch.equal(h, expected[v])
    with pytest.raises(Exception):
        vp.render_to_host(scene, out=ring[:1], spp=1, jitter=False)


@pytest.mark.gpu
def test_backward_replaying_the_primal_records_equals_rerendering():
    """render() keeps the primal's hit lists for the backward pass when the gradient pass would draw the same samples;
    the gradients must equal those of the reference's scheme (render the primal again, RBIntegrator.render_backward),
    for pixel-centre samples and for jittered samples with seed_grad == seed."""
    from volprim_balance_b200 import scene as scene_mod
    n = 4000
    cloud = synthetic.make_cloud(n, synthetic.sigma0_for_hits(n, 14), seed=41, sh_degree=2)
    cam = synthetic.ring_camera(2, 8, 64, 32)
    sd = {"type": "scene", "integrator": {"type": "volprim_rf", "max_depth": 64},
          "primitives": {"type": "ellipsoidsmesh", "centers": cloud.data[:, :3], "scales": cloud.data[:, 3:6],
                         "quaternions": cloud.data[:, 6:], "opacities": cloud.opacities[:, None],
                         "sh_coeffs": cloud.sh_coeffs, "extent": 3.0},
          "sensor": _sensor_dict(cam, "box")}
    w = torch.from_numpy(np.random.default_rng(2).normal(size=(32, 64, 3)).astype(np.float32)).cuda()
    keys = ("primitives.data", "primitives.opacities", "primitives.sh_coeffs")

    def grads(reuse, **kw):
        scene = vp.load_dict(sd)
        params = vp.traverse(scene)
        for k in keys:
            params[k].requires_grad_(True)
        scene_mod.REUSE_PRIMAL_RECORDS = reuse
        try:
            img = vp.render(scene, params, sensor=0, **kw)
            (img * w).sum().backward()
        finally:
            scene_mod.REUSE_PRIMAL_RECORDS = True
        return img.detach(), [params[k].grad.clone() for k in keys]

    for kw in (dict(spp=1, jitter=False), dict(spp=2, seed=5, seed_grad=5, jitter=True)):
        img_a, g_a = grads(True, **kw)
        img_b, g_b = grads(False, **kw)
        assert torch.equal(img_a, img_b)
        for a, b, k in zip(g_a, g_b, keys):
            grad_close(a.cpu().numpy(), b.cpu().numpy(), what=k)   # gather (replay) vs scatter (re-trace): two fp32 evaluations
            assert float(b.abs().max()) > 0
