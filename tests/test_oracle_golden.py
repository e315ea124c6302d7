"""CPU: the oracle (oracle/volprim_oracle.c, float64 build) against fixtures produced by EXECUTING the reference's
own Python source over a torch stand-in for drjit/mitsuba (tests/golden/make_golden.py, refshim.py).
This is what pins the oracle for every formula that lives in /root/reference; the third-party pieces
(closest-hit semantics, SH basis, quat_to_matrix, sRGB) are restated on both sides and remain unpinned."""
import os

import numpy as np
import pytest

from oracle import oracle as O

G = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
KID = {"gaussian": O.GAUSS, "epanechnikov": O.EPAN}


def test_kernel_formulas_match_reference_source():
    z = np.load(os.path.join(G, "kernels.npz"))
    rec, o, d, p = z["rec"], z["o"], z["d"], z["p"]
    n = rec.shape[0]
    for name, kid in KID.items():
        got = np.array([O.kernel_eval(kid, p[i], rec[i], precision="f64") for i in range(n)])
        np.testing.assert_allclose(got, z[f"{name}_eval"], rtol=1e-10, atol=1e-14, err_msg=f"{name}.eval (common.py)")
        for ext in (3.0, 1.0):
            got = np.array([O.density_integral(kid, o[i], d[i], rec[i], ext, precision="f64") for i in range(n)])
            ref = z[f"{name}_density_integral_e{ext:g}"]
            np.testing.assert_allclose(got, ref, rtol=1e-8, atol=1e-12, err_msg=f"{name}.density_integral extent {ext}")
    for ext in (3.0, 1.0):
        res = [O.ray_ellipsoid(o[i], d[i], rec[i], ext, precision="f64") for i in range(n)]
        valid = np.array([r[0] for r in res])
        ref_valid = z[f"isect_valid_e{ext:g}"]
        assert (valid == ref_valid).all()
        assert ref_valid.sum() > 20
        np.testing.assert_allclose(np.array([r[1] for r in res])[valid], z[f"isect_near_e{ext:g}"][valid], rtol=1e-10)
        np.testing.assert_allclose(np.array([r[2] for r in res])[valid], z[f"isect_far_e{ext:g}"][valid], rtol=1e-10)
    # the Epanechnikov full-range integral is clamped to zero at extent 3 (reference quirk Q4)
    assert (z["epanechnikov_density_integral_e3"] == 0).all() and (z["epanechnikov_density_integral_e1"] > 0).any()


SAMPLES = ["sample_rf_gaussian", "sample_rf_epanechnikov", "sample_rf_gaussian_deg1_depth5", "sample_tomo_gaussian",
           "sample_tomo_epanechnikov_extent1", "sample_tomo_epanechnikov_extent3_depth7",
           "sample_tomo_gaussian_hide_maxt", "sample_rf_gaussian_deg2_unnorm_maxt", "sample_rf_epanechnikov_deg0_depth3"]


@pytest.mark.parametrize("name", SAMPLES)
def test_sample_loop_and_adjoint_match_reference_source(name):
    z = np.load(os.path.join(G, name + ".npz"))
    rf = "_rf_" in name
    kernel = KID[str(z["kernel"])]
    sc = O.Scene(z["data"], z["attr"], z["sh"] if rf else None, float(z["extent"]), precision="f64", bvh=False)
    prm = O.Params(integrator=O.RF if rf else O.TOMO, kernel=kernel, max_depth=int(z["max_depth"]),
                   srgb_primitives=bool(z["srgb"]), env=tuple(z["env"]), brute_force=True,
                   hide_emitters=bool(z["hide_emitters"]) if "hide_emitters" in z.files else False)
    hits = z["hits"]
    cap = hits.shape[1]
    res = sc.forward(prm, z["o"], z["d"], z["maxt"], cap=cap)
    # ordered hit lists: identical
    for r in range(hits.shape[0]):
        ref_list = hits[r][hits[r] >= 0]
        got_list = res.hit_ids[r][res.hit_ids[r] >= 0]
        assert list(ref_list) == list(got_list), f"ray {r}: {ref_list} vs {got_list}"
    assert (hits >= 0).sum() > 100
    np.testing.assert_allclose(res.rgb, z["L"], rtol=1e-9, atol=1e-12, err_msg="radiance (sample(), Primal)")
    # PRB adjoint with state_in = the primal's state_out, exactly as RBIntegrator.render_backward calls it
    gd, ga, gs = sc.adjoint(prm, z["o"], z["d"], z["dL"], z["L"], z["maxt"])
    scale = lambda a: max(np.abs(a).max(), 1e-300)
    assert np.abs(gd - z["g_data"]).max() <= 1e-7 * scale(z["g_data"]) + 1e-14, "d primitives.data"
    assert np.abs(ga - z["g_attr"]).max() <= 1e-7 * scale(z["g_attr"]) + 1e-14, "d opacities / sigma_t"
    if rf:
        assert np.abs(gs - z["g_sh"]).max() <= 1e-7 * scale(z["g_sh"]) + 1e-14, "d sh_coeffs"
        assert np.abs(z["g_sh"]).max() > 0
    if "extent3" not in name:
        assert np.abs(z["g_data"]).max() > 0
