"""GPU parity: CUDA path (through the C ABI) against the CPU oracle on identical seeded inputs."""
import numpy as np
import pytest
import torch

from oracle import oracle as O
from volprim_balance_b200 import synthetic
from tests.parity_utils import (RGB_ATOL, RGB_RTOL, check_gradients, compare_forward, f64_reference, gpu_scene, grad_close,
                                make_params, oracle_scene)
from volprim_balance_b200.accel import EllipsoidAccel

pytestmark = pytest.mark.gpu


def _cloud(n=20000, seed=1, crossings=40, deg=3, centers="uniform"):
    return synthetic.make_cloud(n, synthetic.sigma0_for_hits(n, crossings), seed=seed, sh_degree=deg, centers=centers)


def _rays(view=0, w=64, h=48):
    cam = synthetic.ring_camera(view, 8, w, h)
    return synthetic.camera_rays(cam)


@pytest.mark.parametrize("tile", [False, True], ids=["per_ray", "tile"])
@pytest.mark.parametrize("kernel", [0, 1])
@pytest.mark.parametrize("deg", [0, 1, 2, 3])
def test_rf_forward_matches_oracle(kernel, deg, tile):
    """tile=True: image-shaped launch -> warp-cooperative walker; False: explicit ray batch -> per-ray walker."""
    cloud = _cloud(deg=deg)
    o, d, mt = _rays()
    p, op = make_params(0, kernel, max_depth=128, image=(64, 48) if tile else None)
    acc = gpu_scene(cloud)
    res = acc.trace_forward(p, torch.from_numpy(o), torch.from_numpy(d), torch.from_numpy(mt), record_cap=128)
    osc = oracle_scene(cloud)
    ref = osc.forward(op, o, d, mt, cap=128, fragility=True)
    st = compare_forward(res, ref, 128, replay=(osc, op, o, d, mt))
    assert st["mean_hits"] > 5
    print({k: v for k, v in st.items() if k[0] != "_"}, acc.stats())


@pytest.mark.parametrize("tile", [False, True], ids=["per_ray", "tile"])
@pytest.mark.parametrize("kernel", [0, 1])
def test_tomography_forward_matches_oracle(kernel, tile):
    cloud = _cloud(n=5000, crossings=20)
    cloud.extent = 3.0 if kernel == 0 else 1.0  # Epanechnikov integral is clamped to 0 at extent 3 (quirk Q4)
    sig = np.random.default_rng(5).uniform(0.0005, 0.02, cloud.n).astype(np.float32)
    o, d, mt = _rays(view=3)
    p, op = make_params(1, kernel, max_depth=-1, env=(1.0, 0.5, 0.25), image=(64, 48) if tile else None)
    acc = gpu_scene(cloud, attr=sig, sh=False)
    res = acc.trace_forward(p, torch.from_numpy(o), torch.from_numpy(d), torch.from_numpy(mt), record_cap=256)
    osc = oracle_scene(cloud, attr=sig, sh=False)
    ref = osc.forward(op, o, d, mt, cap=256, fragility=True)
    st = compare_forward(res, ref, 256, replay=(osc, op, o, d, mt), srgb=False)
    assert ref.beta.min() < 0.99
    print({k: v for k, v in st.items() if k[0] != "_"})


@pytest.mark.parametrize("kernel,replay,tile", [(0, True, False), (0, False, False), (1, True, True), (0, False, True)])
def test_rf_adjoint_matches_oracle(kernel, replay, tile):
    cloud = _cloud(n=4000, crossings=30)
    o, d, mt = _rays(view=1, w=48, h=32)
    p, op = make_params(0, kernel, max_depth=128, image=(48, 32) if tile else None)
    acc = gpu_scene(cloud)
    to, td, tm = torch.from_numpy(o), torch.from_numpy(d), torch.from_numpy(mt)
    fwd = acc.trace_forward(p, to, td, tm, record_cap=128)
    osc = oracle_scene(cloud)
    ref = osc.forward(op, o, d, mt, cap=128, fragility=True)
    compare_forward(fwd, ref, 128)
    rng = np.random.default_rng(7)
    dL = rng.normal(size=(o.shape[0], 3)).astype(np.float32)
    ids_g = fwd.hit_ids.t().cpu().numpy()
    osc64, same64 = f64_reference(cloud, op, o, d, mt, ids_g, 128)
    same = (ids_g == ref.hit_ids).all(1) & same64
    assert same.mean() > 0.995
    dL[~same] = 0          # keep the comparison on the rays all sides agree on
    # reference_exact: state_in = the primal's state_out (volprim_rf.py:192), here the oracle's own
    gd, ga, gs = acc.trace_adjoint(p, to, td, tm, torch.from_numpy(dL), torch.from_numpy(ref.rgb),
                                   hit_ids=fwd.hit_ids if replay else None,
                                   hit_counts=fwd.nhits if replay else None)
    want, noise = osc64.adjoint(op, o, d, dL, ref.rgb, mt)
    print(check_gradients((gd, ga, gs), want, noise, "rf adjoint"))


@pytest.mark.parametrize("kernel", [0, 1])
def test_tomography_adjoint_matches_oracle(kernel):
    cloud = _cloud(n=3000, crossings=20)
    cloud.extent = 3.0 if kernel == 0 else 1.0
    sig = np.random.default_rng(5).uniform(0.0005, 0.02, cloud.n).astype(np.float32)
    o, d, mt = _rays(view=2, w=48, h=32)
    p, op = make_params(1, kernel, max_depth=-1)
    acc = gpu_scene(cloud, attr=sig, sh=False)
    to, td, tm = torch.from_numpy(o), torch.from_numpy(d), torch.from_numpy(mt)
    osc = oracle_scene(cloud, attr=sig, sh=False)
    ref = osc.forward(op, o, d, mt, cap=256, fragility=True)
    fwd = acc.trace_forward(p, to, td, tm, record_cap=256)
    compare_forward(fwd, ref, 256)
    ids_g = fwd.hit_ids.t().cpu().numpy()
    osc64, same64 = f64_reference(cloud, op, o, d, mt, ids_g, 256, attr=sig, sh=False)
    same = (ids_g == ref.hit_ids).all(1) & same64
    dL = np.random.default_rng(9).normal(size=(o.shape[0], 3)).astype(np.float32)
    dL[~same] = 0
    gd, ga, _ = acc.trace_adjoint(p, to, td, tm, torch.from_numpy(dL), torch.from_numpy(ref.rgb))
    want, noise = osc64.adjoint(op, o, d, dL, ref.rgb, mt)
    print(check_gradients((gd, ga, None), want, noise, "tomography adjoint"))


def test_bvh_is_a_valid_hierarchy():
    cloud = _cloud(n=10000)
    acc = gpu_scene(cloud)
    nodes, perm = acc.debug_bvh()
    nodes, perm = nodes.cpu().numpy(), perm.cpu().numpy()
    n = cloud.n
    assert sorted(perm.tolist()) == list(range(n))
    links = nodes[:, 12:14].copy().view(np.int32)
    seen_leaf, seen_int = np.zeros(n, bool), np.zeros(n - 1, bool)
    seen_int[0] = True
    for i in range(n - 1):
        for c in links[i]:
            if c < 0:
                assert not seen_leaf[~c]
                seen_leaf[~c] = True
            else:
                assert not seen_int[c]
                seen_int[c] = True
    assert seen_leaf.all() and seen_int.all()
    # a parent's child box encloses that child's own two boxes
    for i in range(n - 1):
        for side, c in enumerate(links[i]):
            if c >= 0:
                lo = np.minimum(nodes[c, 0:3], nodes[c, 6:9])
                hi = np.maximum(nodes[c, 3:6], nodes[c, 9:12])
                assert (nodes[i, 6 * side:6 * side + 3] <= lo).all() and (nodes[i, 6 * side + 3:6 * side + 6] >= hi).all()


def test_edge_cases_empty_single_and_missing_rays():
    p, op = make_params(0, 0, max_depth=16)
    one = synthetic.make_cloud(1, 0.2, seed=0, sh_degree=3)
    one.data[0, :3] = 0
    acc = gpu_scene(one)
    o = np.array([[0, 0, -4], [0, 0, -4], [0.1, 0, 0]], np.float32)  # last origin is INSIDE the ellipsoid
    d = np.array([[0, 0, 1], [0, 1, 0], [0, 0, 1]], np.float32)
    res = acc.trace_forward(p, torch.from_numpy(o), torch.from_numpy(d), None, record_cap=4)
    ref = oracle_scene(one).forward(op, o, d, None, cap=4)
    assert res.nhits.cpu().tolist() == ref.nhits.tolist() == [1, 0, 0]
    np.testing.assert_allclose(res.rgb.cpu().numpy(), ref.rgb, atol=1e-5)
    empty = synthetic.make_cloud(0, 0.2, seed=0)
    acc0 = gpu_scene(empty)
    res0 = acc0.trace_forward(p, torch.from_numpy(o), torch.from_numpy(d), None)
    assert res0.nhits.cpu().tolist() == [0, 0, 0] and float(res0.rgb.abs().max()) == 0.0
    r_empty = acc.trace_forward(p, torch.zeros((0, 3)), torch.zeros((0, 3)))
    assert r_empty.rgb.shape == (0, 3)


def test_dense_overlap_falls_back_to_closest_hit_search():
    """More candidates in one interval than the shared-memory list holds (48): 300 nested ellipsoids."""
    rng = np.random.default_rng(3)
    n = 300
    cloud = synthetic.make_cloud(n, 0.05, seed=3, sh_degree=1)
    cloud.data[:, 0:3] = rng.normal(0, 0.003, size=(n, 3))
    cloud.data[:, 3:6] = (0.05 + 0.25 * rng.random((n, 1))) * (1 + 0.2 * rng.random((n, 3)))
    cloud.opacities[:] = 0.02
    o, d, mt = _rays(view=0, w=32, h=16)
    acc = gpu_scene(cloud)
    ref = None
    for image in (None, (32, 16)):      # per-ray walker, and the tile walker handing over to it
        p, op = make_params(0, 0, max_depth=-1, image=image)
        res = acc.trace_forward(p, torch.from_numpy(o), torch.from_numpy(d), torch.from_numpy(mt), record_cap=320)
        ref = ref or oracle_scene(cloud).forward(op, o, d, mt, cap=320, fragility=True)
        st = compare_forward(res, ref, 320, max_fragile_frac=0.02)
        assert ref.nhits.max() > 100
        print(st)


def test_sparse_tiny_primitives_long_intervals():
    """Tiny primitives in a large box: intervals are thousands of primitive radii long, most rays graze or miss.  The
    tile walker's candidate cull works in each primitive's unit-sphere space, where such a segment start lies thousands
    of units from the origin -- its rounding margin must not reject grazing hits.  Hit lists against the oracle."""
    n = 30000
    cloud = synthetic.make_cloud(n, 2e-3, seed=12, sh_degree=0)
    cloud.opacities[:] = 0.5
    o, d, mt = _rays(view=1, w=128, h=64)
    acc = gpu_scene(cloud)
    ref = None
    for image in (None, (128, 64)):
        p, op = make_params(0, 0, max_depth=64, image=image)
        res = acc.trace_forward(p, torch.from_numpy(o), torch.from_numpy(d), torch.from_numpy(mt), record_cap=16)
        ref = ref or oracle_scene(cloud).forward(op, o, d, mt, cap=16, fragility=True)
        st = compare_forward(res, ref, 16)
        assert 0.2 < st["mean_hits"] < 5 and st["robust"] > 0.9 * st["rays"]
        print(st)


def test_far_away_dense_cluster_terminates_and_matches():
    """A heavily overlapping cluster seen from 2000 units away: the interval width that fits the lists is below one
    ulp of t there, so `t_start + delta == t_start` in fp32 (regression guard: the walkers must still advance and
    finish; the closest-hit search takes over).  Both walkers against the oracle."""
    rng = np.random.default_rng(9)
    n = 200
    cloud = synthetic.make_cloud(n, 0.05, seed=9, sh_degree=0)
    cloud.data[:, 0:3] = rng.normal(0, 0.002, size=(n, 3))
    cloud.data[:, 3:6] = (0.04 + 0.1 * rng.random((n, 1))) * (1 + 0.2 * rng.random((n, 3)))
    cloud.opacities[:] = 0.01
    # an 16x8 pixel bundle of nearly parallel rays from z = -2000 through the cluster
    ys, xs = np.meshgrid(np.linspace(-0.25, 0.25, 8), np.linspace(-0.35, 0.35, 16), indexing="ij")
    o = np.stack([xs.ravel(), ys.ravel(), np.full(xs.size, -2000.0)], 1).astype(np.float32)
    d = np.tile(np.array([[0.0, 0.0, 1.0]], np.float32), (xs.size, 1))
    mt = np.full(xs.size, np.finfo(np.float32).max, np.float32)
    acc = gpu_scene(cloud)
    ref = None
    for image in (None, (16, 8)):
        p, op = make_params(0, 0, max_depth=-1, image=image)
        res = acc.trace_forward(p, torch.from_numpy(o), torch.from_numpy(d), torch.from_numpy(mt), record_cap=256)
        torch.cuda.synchronize()
        ref = ref or oracle_scene(cloud).forward(op, o, d, mt, cap=256, fragility=True)
        assert ref.nhits.max() > 50
        st = compare_forward(res, ref, 256, max_fragile_frac=1.0)   # at t = 2000 most lists are fragile; robust ones must match
        print(st)
        assert acc.stats()["stack_overflows"] == 0


def test_tile_walker_fallback_does_not_disturb_neighbour_warps():
    """Coarse image over a dense cloud: tiles are wide, their candidate lists overflow and single warps hand over to
    the per-ray walker while the other warps of the block keep using their shared-memory queues (regression: the
    fallback lists once overlapped the neighbours' queues -> illegal address at 2M primitives)."""
    n = 400_000
    cloud = synthetic.make_cloud(n, 0.0025, seed=2, sh_degree=0)
    o, d, mt = _rays(view=0, w=64, h=32)
    acc = gpu_scene(cloud)
    pt, op = make_params(0, 0, max_depth=128, image=(64, 32))
    pr, _ = make_params(0, 0, max_depth=128)
    to, td, tm = torch.from_numpy(o), torch.from_numpy(d), torch.from_numpy(mt)
    a = acc.trace_forward(pt, to, td, tm, record_cap=128)
    b = acc.trace_forward(pr, to, td, tm, record_cap=128)
    same = (a.hit_ids == b.hit_ids).all(0)
    assert float(same.float().mean()) > 0.995
    ref = oracle_scene(cloud).forward(op, o, d, mt, cap=128, fragility=True)
    print(compare_forward(a, ref, 128))


def test_mixed_block_some_warps_hand_over_to_per_ray_walker():
    """A nested cluster in the image centre inside an ordinary cloud: within one thread block some warps overflow
    their lists and continue with the per-ray walker while their neighbours keep walking cooperatively (regression:
    the hand-over must only touch the thread's own shared-memory slots)."""
    rng = np.random.default_rng(11)
    base = _cloud(n=20000, seed=6, crossings=30)
    m = 300
    nest = synthetic.make_cloud(m, 0.05, seed=7, sh_degree=3)
    nest.data[:, 0:3] = rng.normal(0, 0.003, size=(m, 3))
    nest.data[:, 3:6] = (0.02 + 0.06 * rng.random((m, 1))) * (1 + 0.2 * rng.random((m, 3)))
    nest.opacities[:] = 0.02
    cloud = synthetic.Cloud(np.concatenate([base.data, nest.data]), np.concatenate([base.opacities, nest.opacities]),
                            np.concatenate([base.sh_coeffs, nest.sh_coeffs]), 3.0)
    o, d, mt = _rays(view=2, w=128, h=64)
    acc = gpu_scene(cloud)
    p, op = make_params(0, 0, max_depth=-1, image=(128, 64))
    res = acc.trace_forward(p, torch.from_numpy(o), torch.from_numpy(d), torch.from_numpy(mt), record_cap=400)
    ref = oracle_scene(cloud).forward(op, o, d, mt, cap=400, fragility=True)
    st = compare_forward(res, ref, 400, max_fragile_frac=0.02)
    assert ref.nhits.max() > 150 and np.median(ref.nhits) < 60      # only the centre tiles are dense
    print(st)


GOLDEN_SAMPLES = ["sample_rf_gaussian", "sample_rf_epanechnikov", "sample_rf_gaussian_deg1_depth5", "sample_tomo_gaussian",
                  "sample_tomo_epanechnikov_extent1", "sample_tomo_epanechnikov_extent3_depth7",
                  "sample_tomo_gaussian_hide_maxt", "sample_rf_gaussian_deg2_unnorm_maxt", "sample_rf_epanechnikov_deg0_depth3"]


@pytest.mark.parametrize("tile", [False, True], ids=["per_ray", "tile"])
@pytest.mark.parametrize("name", GOLDEN_SAMPLES)
def test_cuda_path_against_reference_source_fixtures(name, tile):
    """The CUDA path on the inputs of tests/golden/*.npz -- produced by EXECUTING the reference's own volprim_rf.py /
    volprim_tomography.py / common.py in float64 (tests/golden/make_golden.py) -- against the reference's outputs:
    ordered hit lists, radiance, and the PRB gradients.  fp32 vs fp64: a near-tie may swap two hits on a few rays, so
    >= 97 % of the rays must have identical lists, and those must agree in radiance to the stated tolerance."""
    import os
    z = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", name + ".npz"))
    rf = "_rf_" in name
    kernel = {"gaussian": 0, "epanechnikov": 1}[str(z["kernel"])]
    hide = bool(z["hide_emitters"]) if "hide_emitters" in z.files else False
    f32 = lambda a: torch.from_numpy(np.ascontiguousarray(a, dtype=np.float32))
    acc = EllipsoidAccel()
    acc.set_primitives(f32(z["data"]), f32(z["attr"]), f32(z["sh"]) if rf else None, float(z["extent"]))
    acc.build()
    p, _ = make_params(0 if rf else 1, kernel, max_depth=int(z["max_depth"]), srgb=bool(z["srgb"]), hide_emitters=hide,
                       env=tuple(float(x) for x in z["env"]), image=(16, 12) if tile else None)   # fixtures: 16x12 views
    hits = z["hits"]
    cap = hits.shape[1]
    o, d, mt = f32(z["o"]), f32(z["d"]), f32(np.minimum(z["maxt"], np.finfo(np.float32).max))
    res = acc.trace_forward(p, o, d, mt, record_cap=cap)
    ids = res.hit_ids.t().cpu().numpy()[:, :cap]
    nh = res.nhits.cpu().numpy()
    same = np.array([list(hits[r][hits[r] >= 0]) == list(ids[r][:nh[r]]) for r in range(hits.shape[0])])
    assert same.mean() >= 0.97, f"only {same.mean():.3f} of the rays reproduce the reference's hit lists"
    rgb = res.rgb.cpu().numpy()
    ok = np.abs(rgb - z["L"]) <= RGB_ATOL + RGB_RTOL * np.abs(z["L"])
    assert ok[same].all(), f"radiance: max abs diff {np.abs(rgb - z['L'])[same].max()}"
    # Gradients, elementwise at the contract's tolerance.  If fp32 swapped a near-tie on some rays, those rays are taken
    # out on BOTH sides: the float64 oracle (pinned to these fixtures with all rays by tests/test_oracle_golden.py)
    # supplies the gradient of the remaining rays.
    dL = np.array(z["dL"], np.float64)
    dL[~same] = 0
    from oracle import oracle as O
    maxt64 = np.minimum(z["maxt"], np.finfo(np.float32).max)
    op = O.Params(integrator=O.RF if rf else O.TOMO, kernel=kernel, max_depth=int(z["max_depth"]), srgb_primitives=bool(z["srgb"]),
                  hide_emitters=hide, env=tuple(float(x) for x in z["env"]))
    sc = {pr: O.Scene(z["data"], z["attr"], z["sh"] if rf else None, float(z["extent"]), precision=pr) for pr in ("f32", "f64")}
    want = sc["f64"].adjoint(op, z["o"], z["d"], dL, z["L"], maxt64)
    if same.all():      # the float64 oracle IS the fixture (tests/test_oracle_golden.py); say so here too
        np.testing.assert_allclose(want[0], np.asarray(z["g_data"]).reshape(-1, 10), rtol=1e-6, atol=1e-9)
    g32 = sc["f32"].adjoint(op, z["o"], z["d"], dL, z["L"], maxt64)
    noise = tuple(None if a is None else np.abs(a - b) for a, b in zip(g32, want))
    g = acc.trace_adjoint(p, o, d, mt, f32(dL), f32(z["L"]))
    print(name, "identical lists", same.mean(), check_gradients(g, want, noise, name, split_data=False))
