"""CPU: known-answer tests of the oracle from the closed forms in the reference source (SURVEY.md section 8c i-v)."""
import math

import numpy as np
import pytest

from oracle import oracle as O
from volprim_balance_b200 import synthetic


def one_prim(sigma=0.1, opacity=0.7, f_dc=(0.3, -0.2, 0.1), center=(0, 0, 0)):
    data = np.array([[*center, sigma, sigma, sigma, 0, 0, 0, 1]], np.float64)
    sh = np.zeros((1, 48))
    sh[0, :3] = f_dc
    return data, np.array([opacity]), sh


def test_central_hit_rf_closed_form():
    data, op, sh = one_prim()
    sc = O.Scene(data, op, sh, 3.0, precision="f64", bvh=False)
    p = O.Params(srgb_primitives=False)
    r = sc.forward(p, [[0, 0, -2]], [[0, 0, 1]], cap=4)
    alpha = min(0.7, 0.9999)                                  # G = 1 at the centre (volprim_rf.py:63-80)
    col = np.maximum(0.28209479177387814 * np.array([0.3, -0.2, 0.1]) + 0.5, 0)   # volprim_rf.py:91-96
    np.testing.assert_allclose(r.rgb[0], alpha * col, rtol=1e-12)
    assert r.nhits[0] == 1 and r.hit_ids[0, 0] == 0
    np.testing.assert_allclose(r.beta[0], 1 - alpha, rtol=1e-12)
    # entry distance: front face of the bounding ellipsoid at extent * sigma
    np.testing.assert_allclose(r.hit_t[0, 0], 2 - 0.3, rtol=1e-12)


def test_offaxis_gaussian_and_srgb():
    data, op, sh = one_prim(opacity=1.0)
    sc = O.Scene(data, op, sh, 3.0, precision="f64", bvh=False)
    b = 0.15
    r = sc.forward(O.Params(srgb_primitives=False), [[b, 0, -2]], [[0, 0, 1]])
    G = math.exp(-b * b / (2 * 0.1 ** 2))                      # common.py:153-159
    col = np.maximum(0.28209479177387814 * np.array([0.3, -0.2, 0.1]) + 0.5, 0)
    np.testing.assert_allclose(r.rgb[0], G * col, rtol=1e-12)
    r2 = sc.forward(O.Params(srgb_primitives=True), [[b, 0, -2]], [[0, 0, 1]])
    x = G * col
    lin = np.where(x <= 0.04045, x / 12.92, ((x + 0.055) / 1.055) ** 2.4)
    np.testing.assert_allclose(r2.rgb[0], lin, rtol=1e-12)


def test_tomography_central_hit_closed_form():
    sigma, sigma_t = 0.1, 0.05
    data, _, _ = one_prim(sigma)
    sc = O.Scene(data, np.array([sigma_t]), None, 3.0, precision="f64", bvh=False)
    p = O.Params(integrator=O.TOMO, max_depth=-1, env=(1.0, 1.0, 1.0))
    r = sc.forward(p, [[0, 0, -2]], [[0, 0, 1]])
    T = math.exp(-sigma_t / (2 * math.pi * sigma ** 2))        # common.py:204-206
    np.testing.assert_allclose(r.rgb[0], [T, T, T], rtol=1e-12)
    # a ray that misses everything sees the environment, unless hide_emitters (volprim_tomography.py:105-111)
    r = sc.forward(p, [[5, 0, -2]], [[0, 0, 1]])
    np.testing.assert_allclose(r.rgb[0], [1, 1, 1])
    r = sc.forward(O.Params(integrator=O.TOMO, max_depth=-1, hide_emitters=True), [[5, 0, -2], [0, 0, -2]], [[0, 0, 1]] * 2)
    np.testing.assert_allclose(r.rgb[0], [0, 0, 0])
    assert r.rgb[1, 0] > 0


def test_epanechnikov_support_edge():
    data, op, sh = one_prim(sigma=0.1, opacity=1.0)
    for b, inside in ((0.299, True), (0.2999999, True)):
        v = O.rf_transmission(O.EPAN, [b, 0, -2], [0, 0, 1], data[0], 1.0, precision="f64")
        expect = 1 - 0.75 * (1 - (b / 0.3) ** 2)               # common.py:251-259: support 3 s, peak 0.75
        np.testing.assert_allclose(v, expect, rtol=1e-9)
    assert O.kernel_eval(O.EPAN, [0.31, 0, 0], data[0], precision="f64") == 0.0
    assert O.kernel_eval(O.EPAN, [0, 0, 0], data[0], precision="f64") == 0.75


def test_epsilon_advance_skips_entries_and_origin_inside_is_culled():
    # two concentric-ish primitives whose entry points are 5e-5 apart: the second is never reported (Q1)
    d1 = [0, 0, 0, 0.1, 0.1, 0.1, 0, 0, 0, 1]
    d2 = [0, 0, 5e-5, 0.1, 0.1, 0.1, 0, 0, 0, 1]
    d3 = [0, 0, 2e-4, 0.1, 0.1, 0.1, 0, 0, 0, 1]
    sc = O.Scene(np.array([d1, d2, d3], np.float64), np.full(3, 0.2), np.zeros((3, 3)), 3.0, precision="f64", bvh=False)
    r = sc.forward(O.Params(), [[0, 0, -2]], [[0, 0, 1]], cap=4)
    assert list(r.hit_ids[0]) == [0, 2, -1, -1]
    # a primitive that contains the ray origin is never hit (front face behind the origin, back face culled)
    r = sc.forward(O.Params(), [[0, 0, 0.05]], [[0, 0, 1]], cap=4)
    assert r.nhits[0] == 0


def test_max_depth_and_transmittance_cutoff():
    n = 40
    data = np.zeros((n, 10))
    data[:, 2] = np.linspace(0, 3.9, n)
    data[:, 3:6] = 0.02
    data[:, 9] = 1
    sc = O.Scene(data, np.full(n, 0.3), np.zeros((n, 3)), 3.0, precision="f64", bvh=False)
    o, d = [[0, 0, -1]], [[0, 0, 1]]
    assert sc.forward(O.Params(max_depth=5), o, d).nhits[0] == 5
    assert sc.forward(O.Params(max_depth=0), o, d).nhits[0] == 1        # the depth test runs after the first hit (rf:186)
    r = sc.forward(O.Params(max_depth=-1), o, d)
    k = math.ceil(math.log(0.01) / math.log(0.7))                       # beta = 0.7^k <= 0.01  (rf:173-174)
    assert r.nhits[0] == k and abs(r.beta[0] - 0.7 ** k) < 1e-12
    r = sc.forward(O.Params(integrator=O.TOMO, max_depth=-1), o, d)     # tomography has no cut-off (tomo:121-122)
    assert r.nhits[0] == n


def test_bvh_equals_brute_force_and_f32_tracks_f64():
    n = 20000
    cloud = synthetic.make_cloud(n, synthetic.sigma0_for_hits(n, 40), seed=4)
    o, d, mt = synthetic.camera_rays(synthetic.ring_camera(2, 8, 48, 32))
    sc = O.Scene(cloud.data, cloud.opacities, cloud.sh_coeffs, 3.0)
    a = sc.forward(O.Params(max_depth=128), o, d, mt, cap=128)
    b = sc.forward(O.Params(max_depth=128, brute_force=True), o, d, mt, cap=128)
    assert np.array_equal(a.hit_ids, b.hit_ids) and np.array_equal(a.rgb, b.rgb)
    c = O.Scene(cloud.data, cloud.opacities, cloud.sh_coeffs, 3.0, precision="f64").forward(O.Params(max_depth=128), o, d, mt, cap=128)
    same = (a.hit_ids == c.hit_ids).all(1)
    assert same.mean() > 0.99
    assert np.abs(a.rgb - c.rgb)[same].max() < 2e-5


def test_sh_basis_matches_3dgs_constants():
    rng = np.random.default_rng(0)
    d = rng.normal(size=3)
    d /= np.linalg.norm(d)
    x, y, z = d
    Y = O.sh_eval(d, 3, precision="f64")
    C1, C2 = 0.4886025119029199, [1.0925484305920792, -1.0925484305920792, 0.31539156525252005, -1.0925484305920792, 0.5462742152960396]
    C3 = [-0.5900435899266435, 2.890611442640554, -0.4570457994644658, 0.3731763325901154, -0.4570457994644658, 1.445305721320277, -0.5900435899266435]
    ref = [0.28209479177387814, -C1 * y, C1 * z, -C1 * x, C2[0] * x * y, C2[1] * y * z, C2[2] * (2 * z * z - x * x - y * y),
           C2[3] * x * z, C2[4] * (x * x - y * y), C3[0] * y * (3 * x * x - y * y), C3[1] * x * y * z,
           C3[2] * y * (4 * z * z - x * x - y * y), C3[3] * z * (2 * z * z - 3 * x * x - 3 * y * y),
           C3[4] * x * (4 * z * z - x * x - y * y), C3[5] * z * (x * x - y * y), C3[6] * x * (x * x - 3 * y * y)]
    np.testing.assert_allclose(Y, ref, rtol=1e-12, atol=1e-14)


def test_pcg32_matches_the_published_known_answer_vector():
    """PCG32 (pcg-c-basic demo, pcg32_srandom(42, 54)): the generator behind the Russian-roulette restatement."""
    want = [0xa15c02b7, 0x7b47f409, 0xba1d3330, 0x83d2f293, 0xbfa4784b, 0xcbed606e]
    assert [O.pcg32_uint_at(42, 54, k) for k in range(6)] == want
    u = [O.pcg32_float_at(3, 7, k) for k in range(64)]
    assert all(0.0 <= x < 1.0 for x in u) and len(set(u)) == 64
    assert O.pcg32_float_at(3, 7, 5) != O.pcg32_float_at(3, 8, 5) != O.pcg32_float_at(4, 7, 5)


def test_russian_roulette_and_replay_of_the_oracle():
    """volprim_rf.py:177-183: with rr_depth < max_depth, rays with beta in (0.01, 0.1) survive with probability 0.1
    and are rescaled by 10; unbiased.  orc_replay_forward re-evaluates the loop's own hit lists bit for bit."""
    from volprim_balance_b200 import synthetic
    n = 4000
    cloud = synthetic.make_cloud(n, synthetic.sigma0_for_hits(n, 40), seed=2, sh_degree=1)
    o, d, mt = synthetic.camera_rays(synthetic.ring_camera(0, 8, 48, 32))
    sc = O.Scene(cloud.data, cloud.opacities, cloud.sh_coeffs, 3.0)
    off = O.Params(integrator=O.RF, max_depth=64)
    assert not off.use_rr and not O.Params(max_depth=64, rr_depth=64).use_rr and O.Params(max_depth=-1, rr_depth=0).use_rr
    a = sc.forward(off, o, d, mt, cap=64)
    rp = sc.replay(off, o, d, mt, a.hit_ids, a.nhits)
    assert rp["valid"].all() and np.array_equal(rp["rgb"], a.rgb) and np.array_equal(rp["beta"], a.beta)
    assert (rp["hit_beta"][:, 0] == 1).all() and (np.diff(rp["hit_beta"], axis=1)[a.hit_ids[:, 1:] >= 0] <= 0).all()
    means = []
    for seed in range(8):
        b = sc.forward(O.Params(integrator=O.RF, max_depth=64, rr_depth=2, rr_seed=seed), o, d, mt, cap=64)
        assert (b.nhits < a.nhits).mean() > 0.2          # most rays that reach beta < 0.1 are ended there
        means.append(b.rgb.mean())
    assert abs(np.mean(means) - a.rgb.mean()) < 0.01
    # a list with a primitive the ray never enters is reported invalid
    bad = a.hit_ids.copy()
    far = int(np.argmax(np.linalg.norm(cloud.data[:, :3] - (o[0] + d[0] * 4), axis=1)))
    bad[0, 0] = far
    assert not sc.replay(off, o, d, mt, bad, a.nhits)["valid"][0]
