"""GPU: sensor-fused ray generation, compressed hit records, the gather adjoint, Russian roulette, the film kernels and
row-band rendering -- every new C-ABI entry point of round 2 against the oracle / a float64 restatement."""
import numpy as np
import pytest
import torch

import volprim_balance_b200 as vp
from oracle import oracle as O
from volprim_balance_b200 import _cabi, synthetic
from volprim_balance_b200.accel import HitRecord, RaySource
from tests.parity_utils import (RGB_ATOL, RGB_RTOL, GradientReference, check_gradients, compare_forward, f64_reference, gpu_scene,
                                grad_close, make_params, oracle_scene, record_lists, robust_mask)

pytestmark = pytest.mark.gpu


def _cloud(n=20000, seed=1, crossings=40, deg=3):
    return synthetic.make_cloud(n, synthetic.sigma0_for_hits(n, crossings), seed=seed, sh_degree=deg)


def _sensor(cam, rfilter="box"):
    return vp.PerspectiveSensor({"type": "perspective", "fov": cam.fov_x_deg, "fov_axis": "x",
                                 "to_world": vp.Transform4f(cam.to_world), "near_clip": cam.near_clip, "far_clip": cam.far_clip,
                                 "film": {"type": "hdrfilm", "width": cam.width, "height": cam.height, "rfilter": {"type": rfilter}}})


def _gpu_rays(acc, sensor, spp=1, jit=None):
    """The rays the fused kernels generate, as numpy arrays for the oracle (bit-identical: one generation routine)."""
    return tuple(x.cpu().numpy() for x in acc.raygen_perspective(sensor.vp_camera(), spp, jit))


@pytest.mark.parametrize("spp,jitter", [(1, False), (2, True)])
def test_fused_raygen_equals_explicit_rays(spp, jitter):
    """vp_render_forward with a camera (rays generated inside the trace kernel) == vp_raygen_perspective followed by
    vp_trace_forward on the explicit rays: bit-identical radiance, transmittance and hit counts."""
    cloud = _cloud()
    cam = synthetic.ring_camera(2, 8, 64, 48)
    acc = gpu_scene(cloud)
    s = _sensor(cam)
    jit = torch.rand((64 * 48 * spp, 2), device="cuda") if jitter else None
    p, _ = make_params(0, 0, 64, image=(64 * spp, 48))
    o, d, mt = acc.raygen_perspective(s.vp_camera(), spp, jit)
    a = acc.trace_forward(p, o, d, mt)
    p0, _ = make_params(0, 0, 64)
    b = acc.render_forward(p0, RaySource(camera=s.vp_camera(), spp=spp, jitter=jit))
    assert torch.equal(a.rgb, b.rgb) and torch.equal(a.beta, b.beta) and torch.equal(a.nhits, b.nhits)
    # a row band of the sensor == the same rows of the full image
    c = acc.render_forward(p0, RaySource(camera=s.vp_camera(), spp=spp, jitter=None if jit is None else jit[16 * 64 * spp:40 * 64 * spp],
                                         rows=(16, 24)))
    assert torch.equal(c.rgb, b.rgb[16 * 64 * spp:40 * 64 * spp])


def test_repeated_builds_are_bit_reproducible():
    """Replicas of a multi-GPU job build their own LBVH; everything the walk derives from the build (initial interval
    width, tile-vs-per-ray heuristic) must come out bit-identical however the build's warps were scheduled -- otherwise
    fragile rays order their hits differently from replica to replica.  Eight independent builds of one cloud render
    one view to the same bits (the sums the build takes over all leaves are integer accumulations)."""
    cloud = synthetic.make_cloud(60_000, synthetic.sigma0_for_hits(60_000, 30.0), seed=11)
    cam = synthetic.ring_camera(3, 8, 128, 96)
    p, _ = make_params(0, 0, 64)
    ref = None
    for _ in range(8):
        acc = gpu_scene(cloud)
        r = acc.render_forward(p, RaySource(camera=_sensor(cam).vp_camera()))
        got = (r.rgb.clone(), r.beta.clone(), r.nhits.clone())
        if ref is None:
            ref = got
        assert all(torch.equal(a, b) for a, b in zip(ref, got))
        del acc


@pytest.mark.parametrize("scratch", [None, 1 << 20])
def test_compressed_hit_records_equal_the_dense_lists(scratch):
    """vp_render_forward(record): ray_offsets = exclusive scan of the hit counts, ids = the dense lists without padding,
    the same with a 1 MiB scratch (many row bands)."""
    cloud = _cloud()
    cam = synthetic.ring_camera(1, 8, 128, 64)
    acc = gpu_scene(cloud)
    if scratch:
        acc.set_option("record_scratch_bytes", scratch)       # 1 MiB / (64 * 4 B) = 4096 rays = 32 rows per band
    s = _sensor(cam)
    p, _ = make_params(0, 0, 64, image=(128, 64))
    o, d, mt = acc.raygen_perspective(s.vp_camera(), 1, None)
    dense = acc.trace_forward(p, o, d, mt, record_cap=64)
    p0, _ = make_params(0, 0, 64)
    res = acc.render_forward(p0, RaySource(camera=s.vp_camera()), record=acc.new_record(128 * 64, 64, with_state=False, dense=False), id_cap=64)
    rec = res.record
    assert rec.usable() and torch.equal(res.rgb, dense.rgb) and torch.equal(res.nhits, dense.nhits)
    nh = dense.nhits.cpu().numpy().astype(np.int64)
    off = rec.ray_offsets.cpu().numpy()
    assert off[0] == 0 and (np.diff(off) == nh).all() and rec.totals() == (int(nh.sum()), 0)
    ids = rec.ids.cpu().numpy()[:off[-1]]
    want = dense.hit_ids.t().cpu().numpy()
    assert (ids == want[want >= 0]).all()                       # row-major, no padding
    # bytes kept per view: 4 per recorded hit (+ the ray offsets)
    assert rec.ids.numel() >= off[-1] and rec.nbytes() == rec.ids.numel() * 4 + (128 * 64 + 1) * 8
    # the state record: (colour, transmittance) per hit -- the product of (1 - T) along a row is the ray's final beta,
    # and compositing the recorded colours reproduces the image (sRGB -> linear at the end)
    rich = acc.render_forward(p0, RaySource(camera=s.vp_camera()), record=acc.new_record(128 * 64, 64, with_state=True, dense=False), id_cap=64)
    assert rich.record.usable() and torch.equal(rich.record.ids[:off[-1]], rec.ids[:off[-1]]) and torch.equal(rich.rgb, res.rgb)
    st = rich.record.state[:off[-1]].cpu().numpy().astype(np.float64)
    row = np.repeat(np.arange(128 * 64), nh)
    logT = np.zeros(128 * 64)
    np.add.at(logT, row, np.log(st[:, 3]))
    np.testing.assert_allclose(np.exp(logT), res.beta.cpu().numpy(), rtol=2e-5, atol=1e-7)
    beta_before = np.exp(np.concatenate([np.cumsum(np.log(st[off[r]:off[r + 1], 3])) - np.log(st[off[r]:off[r + 1], 3]) for r in range(0, 128 * 64, 97)]))
    rows97 = np.concatenate([np.arange(off[r], off[r + 1]) for r in range(0, 128 * 64, 97)])
    Ls = np.zeros((128 * 64, 3))
    np.add.at(Ls, row[rows97], (beta_before * (1 - st[rows97, 3]))[:, None] * st[rows97, :3])
    lin = np.where(Ls <= 0.04045, Ls / 12.92, ((np.maximum(Ls, 0.04045) + 0.055) / 1.055) ** 2.4)[::97]
    np.testing.assert_allclose(lin, res.rgb.cpu().numpy()[::97], rtol=1e-4, atol=1e-6)
    # the dense record: the same lists and states hit-major in the record's own buffers, no compaction pass
    dn = acc.render_forward(p0, RaySource(camera=s.vp_camera()), record=acc.new_record(128 * 64, 64, dense=True), id_cap=64)
    assert dn.record.dense and dn.record.usable() and dn.record.totals() == (int(nh.sum()), 0)
    assert torch.equal(dn.record.counts, dense.nhits) and torch.equal(dn.rgb, res.rgb)
    valid = torch.arange(64, device="cuda")[:, None] < dn.record.counts[None, :]
    assert torch.equal(dn.record.ids[valid], dense.hit_ids[valid])
    assert torch.equal(dn.record.state.permute(1, 0, 2)[valid.t()], rich.record.state[:off[-1]])
    # a record that is too small is reported, not silently truncated
    small = acc.new_record(128 * 64, 64, capacity=1000, dense=False)
    acc.render_forward(p0, RaySource(camera=s.vp_camera()), record=small, id_cap=64)
    assert not small.usable() and small.totals()[0] == int(nh.sum())
    cut = acc.new_record(128 * 64, 8, dense=False)
    acc.render_forward(p0, RaySource(camera=s.vp_camera()), record=cut, id_cap=8)
    assert not cut.usable() and cut.totals()[1] == int((nh > 8).sum())


@pytest.mark.parametrize("kind", ["dense", "rows_state", "rows"])
@pytest.mark.parametrize("kernel,deg", [(0, 3), (1, 3), (0, 1), (1, 0), (0, 2)])
def test_gather_adjoint_matches_oracle_and_scatter_adjoint(kernel, deg, kind):
    """volprim_rf adjoint through vp_render_adjoint (ray-major pass into per-primitive buckets + one warp per
    primitive, no global reductions) against the oracle, elementwise at 1e-3, and against the scatter formulation."""
    cloud = _cloud(n=6000, crossings=35, deg=deg)
    cam = synthetic.ring_camera(1, 8, 64, 32)
    acc = gpu_scene(cloud)
    s = _sensor(cam)
    o, d, mt = _gpu_rays(acc, s)
    p, op = make_params(0, kernel, 128)
    rays = RaySource(camera=s.vp_camera())
    fwd = acc.render_forward(p, rays, record=acc.new_record(o.shape[0], 128, with_state=kind != "rows", dense=kind == "dense"), id_cap=128)
    assert (fwd.record.state is not None) == (kind != "rows") and fwd.record.dense == (kind == "dense") and fwd.record.usable()
    osc = oracle_scene(cloud)
    ref = osc.forward(op, o, d, mt, cap=128, fragility=True)
    ids_g, _ = record_lists(fwd.record, range(o.shape[0]), 128)
    st = compare_forward(fwd, ref, 128, replay=(osc, op, o, d, mt), ids_g=ids_g)
    dL = np.random.default_rng(7).normal(size=(o.shape[0], 3)).astype(np.float32)
    osc64, same64 = f64_reference(cloud, op, o, d, mt, ids_g, 128)
    dL[~(st["_same"] & same64)] = 0                                 # keep the comparison on rays all sides agree on
    dL[::7] = 0                                                     # and exercise the "gradient is zero: skip" branch
    gd, ga, gs = acc.render_adjoint(p, rays, torch.from_numpy(dL), torch.from_numpy(ref.rgb), fwd.record)
    want, noise = osc64.adjoint(op, o, d, dL, ref.rgb, mt)
    e = check_gradients((gd, ga, gs), want, noise, "gather adjoint")
    # scatter formulation (vector reductions) replaying the SAME lists from a dense record
    to, td, tm = (torch.from_numpy(x) for x in (o, d, mt))
    p_img, _ = make_params(0, kernel, 128, image=(64, 32))
    dense = acc.trace_forward(p_img, to, td, tm, record_cap=128)
    assert (dense.hit_ids.t().cpu().numpy() == ids_g).all()
    sd, sa, ss = acc.trace_adjoint(p, to, td, tm, torch.from_numpy(dL), torch.from_numpy(ref.rgb),
                                   hit_ids=dense.hit_ids, hit_counts=dense.nhits)
    check_gradients((sd, sa, ss), want, noise, "scatter adjoint on the same lists")
    # the call ADDS: a second pass doubles the buffers; primitive ranges compose
    out = (gd.clone(), ga.clone(), gs.clone())
    acc.adjoint_begin(p, rays, torch.from_numpy(dL), torch.from_numpy(ref.rgb), fwd.record, out)
    for p0, p1 in ((0, 1000), (1000, 1001), (1001, cloud.n)):
        acc.adjoint_finish(p, rays, fwd.record, p0, p1, out)
    for x, y in zip(out, (gd, ga, gs)):
        grad_close(x.cpu().numpy(), 2 * y.cpu().numpy(), rtol=1e-5, what="accumulate + ranges")
    print("gather adjoint errors", e)


def test_tomography_adjoint_replays_compressed_records():
    cloud = _cloud(n=3000, crossings=20, deg=0)
    sig = np.random.default_rng(5).uniform(0.0005, 0.02, cloud.n).astype(np.float32)
    cam = synthetic.ring_camera(2, 8, 48, 32)
    acc = gpu_scene(cloud, attr=sig, sh=False)
    o, d, mt = _gpu_rays(acc, _sensor(cam))
    p, op = make_params(1, 0, -1)
    rays = RaySource(camera=_sensor(cam).vp_camera())
    fwd = acc.render_forward(p, rays, record=True, id_cap=256)
    osc = oracle_scene(cloud, attr=sig, sh=False)
    ref = osc.forward(op, o, d, mt, cap=256, fragility=True)
    ids_g, _ = record_lists(fwd.record, range(o.shape[0]), 256)
    st = compare_forward(fwd, ref, 256, replay=(osc, op, o, d, mt), srgb=False, ids_g=ids_g)
    dL = np.random.default_rng(9).normal(size=(o.shape[0], 3)).astype(np.float32)
    osc64, same64 = f64_reference(cloud, op, o, d, mt, ids_g, 256, attr=sig, sh=False)
    dL[~(st["_same"] & same64)] = 0
    gd, ga, _ = acc.render_adjoint(p, rays, torch.from_numpy(dL), torch.from_numpy(ref.rgb), fwd.record)
    want, noise = osc64.adjoint(op, o, d, dL, ref.rgb, mt)
    print(check_gradients((gd, ga, None), want, noise, "tomography replay"))


@pytest.mark.parametrize("tile", [False, True], ids=["per_ray", "tile"])
def test_russian_roulette_matches_oracle(tile):
    """rr_depth < max_depth: the primal pass terminates rays with beta in (0.01, 0.1) with probability 0.9 and
    rescales the survivors (volprim_rf.py:177-183); PCG32 stream per ray, identical on both sides."""
    cloud = _cloud(n=20000, crossings=45)
    cam = synthetic.ring_camera(0, 8, 64, 48)
    o, d, mt = synthetic.camera_rays(cam)
    acc = gpu_scene(cloud)
    p, op = make_params(0, 0, 128, rr_depth=3, rr_seed=17, rr_skip=2, image=(64, 48) if tile else None)
    assert p.use_rr == 1
    res = acc.trace_forward(p, torch.from_numpy(o), torch.from_numpy(d), torch.from_numpy(mt), record_cap=128)
    osc = oracle_scene(cloud)
    ref = osc.forward(op, o, d, mt, cap=128, fragility=True)
    p_off, op_off = make_params(0, 0, 128)
    ref_off = osc.forward(op_off, o, d, mt, cap=128)
    assert (ref.nhits != ref_off.nhits).mean() > 0.2           # roulette really changes where rays end ...
    assert np.abs(ref.rgb.mean() - ref_off.rgb.mean()) < 0.02  # ... and stays unbiased on average
    # (the fragile-ray replay is off here: it evaluates a list without the roulette's rescaling)
    st = compare_forward(res, ref, 128)
    print({k: v for k, v in st.items() if k[0] != "_"})
    # the plugin accepts the configuration and produces the same image through sample()
    scene = vp.load_dict({"type": "scene", "integrator": {"type": "volprim_rf", "max_depth": 128, "rr_depth": 3, "rr_seed": 17, "rr_skip": 2},
                          "primitives": {"type": "ellipsoidsmesh", "centers": cloud.data[:, :3], "scales": cloud.data[:, 3:6],
                                         "quaternions": cloud.data[:, 6:], "opacities": cloud.opacities[:, None],
                                         "sh_coeffs": cloud.sh_coeffs, "extent": 3.0}})
    assert scene.integrator.use_rr
    L, *_ = scene.integrator.sample(vp.ADMode.Primal, scene, None, vp.Ray3f(torch.from_numpy(o).cuda(), torch.from_numpy(d).cuda(),
                                                                              torch.from_numpy(mt).cuda()))
    if not tile:
        assert torch.equal(L, res.rgb)


def _film_reference(W, H, spp, rfilter, jit, L):
    """float64 restatement of the film: separable filter weights, weighted sum / sum of weights."""
    rad = {"box": 0.5, "tent": 1.0, "gaussian": 2.0}[rfilter]

    def f(dist):
        dist = np.abs(dist)
        if rfilter == "box":
            return (dist <= 0.5).astype(np.float64)
        if rfilter == "tent":
            return np.maximum(0.0, 1.0 - dist)
        return np.maximum(0.0, np.exp(-2.0 * dist * dist) - np.exp(-8.0))

    acc = np.zeros((H, W, 4))
    pix = np.repeat(np.arange(W * H), spp)
    sx = pix % W + jit[:, 0].astype(np.float64)
    sy = pix // W + jit[:, 1].astype(np.float64)
    contrib = []
    for i in range(len(pix)):
        x0, x1 = int(np.floor(sx[i] - rad - 0.5)) + 1, int(np.ceil(sx[i] + rad - 0.5)) - 1
        y0, y1 = int(np.floor(sy[i] - rad - 0.5)) + 1, int(np.ceil(sy[i] + rad - 0.5)) - 1
        mine = []
        for py in range(max(y0, 0), min(y1, H - 1) + 1):
            for px in range(max(x0, 0), min(x1, W - 1) + 1):
                w = f(sx[i] - (px + 0.5)) * f(sy[i] - (py + 0.5))
                if w > 0:
                    acc[py, px, :3] += w * L[i]
                    acc[py, px, 3] += w
                    mine.append((py, px, w))
        contrib.append(mine)
    img = np.where(acc[..., 3:] > 0, acc[..., :3] / np.maximum(acc[..., 3:], 1e-300), 0.0)
    return img, acc, contrib


@pytest.mark.parametrize("rfilter", ["box", "tent", "gaussian"])
@pytest.mark.parametrize("spp", [1, 3])
def test_film_kernels_match_float64_restatement_per_pixel(rfilter, spp):
    """vp_film_splat / develop / adjoint: every pixel and every sample gradient against the float64 restatement (true
    filter radii: tent 1, gaussian 2 = 4 standard deviations of 0.5)."""
    W, H = 24, 16
    rng = np.random.default_rng(3)
    jit = rng.random((W * H * spp, 2)).astype(np.float32)
    L = rng.random((W * H * spp, 3)).astype(np.float32)
    acc = gpu_scene(synthetic.make_cloud(4, 0.1, seed=0))
    f = _cabi.RFILTERS[rfilter]
    tj, tL = torch.from_numpy(jit).cuda(), torch.from_numpy(L).cuda()
    accum = torch.zeros(W * H * 4, device="cuda")
    acc.film_splat(W, H, spp, f, tj, tL, accum)
    wide = torch.zeros((H, 3 * W, 3), device="cuda")                 # develop into the middle block of a batch film
    acc.film_develop(W, H, accum, wide[:, W:2 * W])
    img, ref_acc, contrib = _film_reference(W, H, spp, rfilter, jit, L.astype(np.float64))
    np.testing.assert_allclose(wide[:, W:2 * W].cpu().numpy(), img, rtol=2e-5, atol=2e-6)
    assert float(wide[:, :W].abs().max()) == 0 and float(wide[:, 2 * W:].abs().max()) == 0
    d_img = rng.normal(size=(H, 3 * W, 3)).astype(np.float32)
    dL = acc.film_adjoint(W, H, spp, f, tj, accum, torch.from_numpy(d_img).cuda()[:, W:2 * W]).cpu().numpy()
    want = np.zeros((W * H * spp, 3))
    for i, mine in enumerate(contrib):
        for py, px, w in mine:
            want[i] += d_img[py, W + px].astype(np.float64) * w / ref_acc[py, px, 3]
    np.testing.assert_allclose(dL, want, rtol=2e-5, atol=2e-6)
    # pixel centres (jitter = NULL): box and tent return the samples themselves
    if spp == 1 and rfilter != "gaussian":
        accum.zero_()
        acc.film_splat(W, H, 1, f, None, tL, accum)
        out = torch.empty((H, W, 3), device="cuda")
        acc.film_develop(W, H, accum, out)
        assert torch.equal(out.reshape(-1, 3), tL)


def test_render_batch_sensor_with_filters_matches_film_restatement_per_pixel():
    """render() of a batch sensor with jittered samples: every pixel of every view equals the float64 film
    restatement applied to the radiance of the very samples render() drew (tent and gaussian at their true radii),
    and the autograd gradient equals the oracle adjoint fed with the restated film adjoint."""
    from volprim_balance_b200 import scene as scene_mod
    n = 3000
    cloud = synthetic.make_cloud(n, synthetic.sigma0_for_hits(n, 20), seed=31, sh_degree=1)
    cams = [synthetic.ring_camera(i, 8, 40, 24) for i in (0, 3)]
    sd = {"type": "scene", "integrator": {"type": "volprim_rf", "max_depth": 64, "srgb_primitives": False},
          "primitives": {"type": "ellipsoidsmesh", "centers": cloud.data[:, :3], "scales": cloud.data[:, 3:6],
                         "quaternions": cloud.data[:, 6:], "opacities": cloud.opacities[:, None],
                         "sh_coeffs": cloud.sh_coeffs, "extent": 3.0}}
    scene = vp.load_dict(sd)
    osc = oracle_scene(cloud)
    gref = GradientReference(cloud)
    op = O.Params(integrator=O.RF, kernel=O.GAUSS, max_depth=64, srgb_primitives=False)
    for rfilter in ("tent", "gaussian"):
        sens = {f"cam_{i}": {"type": "perspective", "fov": c.fov_x_deg, "fov_axis": "x", "to_world": vp.Transform4f(c.to_world),
                             "near_clip": c.near_clip, "far_clip": c.far_clip,
                             "film": {"type": "hdrfilm", "width": 40, "height": 24}} for i, c in enumerate(cams)}
        batch = vp.load_dict({"type": "batch", "film": {"type": "hdrfilm", "width": 80, "height": 24, "rfilter": {"type": rfilter}}, **sens})
        spp, seed = 2, 3
        acc = scene.ellipsoids().accel()
        scene.ellipsoids().bind("opacities", with_sh=True)
        # oracle first: pixels that receive a sample whose hit list may legally differ in fp32 carry no gradient weight
        w = np.random.default_rng(11).normal(size=(24, 80, 3)).astype(np.float32)
        views = []
        for vi, c in enumerate(cams):
            jit_t = scene_mod._sample_positions(40, 24, spp, seed * 7919 + vi, True, torch.device("cuda"))
            o, d, mt = _gpu_rays(acc, batch.sensors[vi], spp, jit_t)
            ref = osc.forward(op, o, d, mt, cap=64, fragility=True)
            want, ref_acc, contrib = _film_reference(40, 24, spp, rfilter, jit_t.cpu().numpy(), ref.rgb.astype(np.float64))
            okpix = np.ones((24, 40), bool)
            for i in np.flatnonzero(~robust_mask(ref)):
                for py, px, _ in contrib[i]:
                    okpix[py, px] = False
            assert okpix.mean() > 0.6      # (the gaussian filter spreads every excluded sample over 16 pixels)
            w[:, 40 * vi:40 * (vi + 1)][~okpix] = 0
            views.append((o, d, mt, ref, want, ref_acc, contrib, okpix))
        params = vp.traverse(scene)
        keys = ["primitives.data", "primitives.opacities", "primitives.sh_coeffs"]
        for k in keys:
            params[k].requires_grad_(True)
            params[k].grad = None
        img = vp.render(scene, params, sensor=batch, spp=spp, seed=seed, seed_grad=seed)
        assert img.shape == (24, 80, 3)
        (img * torch.from_numpy(w).cuda()).sum().backward()
        tot = [np.zeros((n, 10)), np.zeros(n), np.zeros((n, 12))]
        tot_noise = [np.zeros((n, 10)), np.zeros(n), np.zeros((n, 12))]
        for vi, (o, d, mt, ref, want, ref_acc, contrib, okpix) in enumerate(views):
            got = img.detach()[:, 40 * vi:40 * (vi + 1)].cpu().numpy()
            ok = np.abs(got - want) <= RGB_ATOL + RGB_RTOL * np.abs(want)
            assert ok[okpix].all(), f"{rfilter} view {vi}: max pixel diff {np.abs(got - want)[okpix].max()}"
            dL = np.zeros((o.shape[0], 3))
            for i, mine in enumerate(contrib):
                for py, px, wt in mine:
                    dL[i] += w[py, 40 * vi + px].astype(np.float64) * wt / ref_acc[py, px, 3]
            g, nz = gref.adjoint(op, o, d, dL, ref.rgb, mt)
            for k in range(3):
                tot[k] += g[k]
                tot_noise[k] += nz[k]
        check_gradients([params[k].grad for k in keys], tot, tot_noise, rfilter + " batch film")


def test_render_rows_band_equals_rows_of_the_full_image_and_gradients_sum():
    """render(rows=(y0, y1)): what every rank of an image-tile-sharded view computes.  The bands tile the image
    exactly, and the band gradients add up to the gradient of the full image."""
    n = 4000
    cloud = synthetic.make_cloud(n, synthetic.sigma0_for_hits(n, 25), seed=5, sh_degree=2)
    cam = synthetic.ring_camera(4, 8, 64, 48)
    sd = {"type": "scene", "integrator": {"type": "volprim_rf", "max_depth": 64},
          "primitives": {"type": "ellipsoidsmesh", "centers": cloud.data[:, :3], "scales": cloud.data[:, 3:6],
                         "quaternions": cloud.data[:, 6:], "opacities": cloud.opacities[:, None],
                         "sh_coeffs": cloud.sh_coeffs, "extent": 3.0}}
    scene = vp.load_dict(sd)
    s = _sensor(cam)
    full = vp.render(scene, sensor=s, spp=1, jitter=False)
    from volprim_balance_b200 import parallel
    bands = [parallel.shard_rows(48, r, 5) for r in range(5)]
    got = torch.cat([vp.render(scene, sensor=s, spp=1, jitter=False, rows=b) for b in bands if b[1] > b[0]], dim=0)
    assert torch.equal(got, full)
    w = torch.from_numpy(np.random.default_rng(2).normal(size=(48, 64, 3)).astype(np.float32)).cuda()
    keys = ["primitives.data", "primitives.opacities", "primitives.sh_coeffs"]

    def grads(rows_list):
        params = vp.traverse(scene)
        tot = None
        for rows in rows_list:
            for k in keys:
                params[k].requires_grad_(True)
                params[k].grad = None
            img = vp.render(scene, params, sensor=s, spp=1, jitter=False, rows=rows)
            ww = w if rows is None else w[rows[0]:rows[1]]
            (img * ww).sum().backward()
            g = [params[k].grad.clone() for k in keys]
            tot = g if tot is None else [a + b for a, b in zip(tot, g)]
        return tot

    a, b = grads([None]), grads([bd for bd in bands if bd[1] > bd[0]])
    for x, y, k in zip(a, b, keys):
        grad_close(y.cpu().numpy(), x.cpu().numpy(), rtol=1e-4, what="bands vs full " + k)
    assert parallel.render_tiles(scene, s, vp.render, spp=1, jitter=False).shape == (48, 64, 3)   # world size 1: the band is the image


def test_backward_falls_back_to_retracing_when_the_record_does_not_fit():
    """A hit record that overflows its capacity is not replayed: render()'s backward pass re-traces, same gradients."""
    n = 4000
    cloud = synthetic.make_cloud(n, synthetic.sigma0_for_hits(n, 25), seed=9, sh_degree=1)
    cam = synthetic.ring_camera(1, 8, 64, 32)
    sd = {"type": "scene", "integrator": {"type": "volprim_rf", "max_depth": 64},
          "primitives": {"type": "ellipsoidsmesh", "centers": cloud.data[:, :3], "scales": cloud.data[:, 3:6],
                         "quaternions": cloud.data[:, 6:], "opacities": cloud.opacities[:, None],
                         "sh_coeffs": cloud.sh_coeffs, "extent": 3.0}}
    w = torch.from_numpy(np.random.default_rng(2).normal(size=(32, 64, 3)).astype(np.float32)).cuda()
    keys = ("primitives.data", "primitives.opacities", "primitives.sh_coeffs")
    out = []
    for estimate in (48.0, 0.01):          # 0.01 hits per ray: the record cannot hold the lists
        scene = vp.load_dict(sd)
        params = vp.traverse(scene)
        for k in keys:
            params[k].requires_grad_(True)
        scene.ellipsoids().accel().hits_per_ray_estimate = estimate
        img = vp.render(scene, params, sensor=_sensor(cam), spp=1, jitter=False)
        rec = scene.integrator.last.record
        assert rec is not None and rec.usable() == (estimate > 1)
        (img * w).sum().backward()
        out.append([params[k].grad.clone() for k in keys])
        if estimate < 1:
            assert scene.ellipsoids().accel().hits_per_ray_estimate > 4      # the next record is sized from this one
    for a, b, k in zip(out[0], out[1], keys):     # two fp32 evaluations (gather from the far origin vs scatter): 2e-3
        grad_close(b.cpu().numpy(), a.cpu().numpy(), rtol=2e-3, what="replay vs re-trace fallback " + k)


def test_misaligned_gradient_buffers_are_refused_and_unaligned_adam_grads_work():
    """vector reductions need aligned gradient rows (VP_E_INVALID instead of a misaligned-address fault), and
    BoundedAdam accepts gradients that are slices of a packed buffer at any offset."""
    cloud = _cloud(n=1001, crossings=10)
    acc = gpu_scene(cloud)
    o, d, mt = synthetic.camera_rays(synthetic.ring_camera(0, 8, 16, 8))
    p, _ = make_params(0, 0, 16)
    flat = torch.zeros(cloud.n * 59 + 8, device="cuda")
    bad = (flat[1:1 + cloud.n * 10], flat[cloud.n * 10 + 4:cloud.n * 11 + 4], flat[cloud.n * 11 + 4:cloud.n * 59 + 4])
    with pytest.raises(vp.VolprimCudaError, match="aligned"):
        acc.trace_adjoint(p, torch.from_numpy(o), torch.from_numpy(d), torch.from_numpy(mt), torch.ones(128, 3), torch.zeros(128, 3), out=bad)
    from volprim_balance_b200 import optimizers, parallel
    opt = optimizers.BoundedAdam(lr=0.1)
    opt["x"] = torch.rand(1001, device="cuda")
    before = opt["x"].detach().clone()
    opt["x"].grad = torch.ones(1002, device="cuda")[1:]           # 4-byte offset
    opt.step()
    assert float((before - opt["x"].detach()).abs().min()) > 0.05
    g = {"a": torch.ones(1001, device="cuda"), "b": torch.ones(7, device="cuda"), "c": torch.ones(13, device="cuda")}
    flat = parallel.pack_gradients(g, sorted(g))
    un = parallel.unpack_gradients(flat, g, sorted(g))
    assert all(v.data_ptr() % 16 == 0 for v in un.values()) and all(torch.equal(un[k], g[k]) for k in g)
