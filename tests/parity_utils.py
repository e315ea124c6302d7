"""Shared helpers of the GPU parity tests: build the same cloud on both sides, compare hit lists with the
documented exclusion (exact-depth ties / entries at the epsilon cull / grazing entries),
and compare gradients ELEMENTWISE with the contract's tolerance."""
import numpy as np
import torch

from oracle import oracle as O
from volprim_balance_b200 import _cabi, synthetic
from volprim_balance_b200.accel import EllipsoidAccel

# tolerances stated by BASELINE.json north_star: radiance / transmittance 1e-4 absolute + 1e-3 relative,
# gradients 1e-3 relative
RGB_ATOL, RGB_RTOL = 1e-4, 1e-3
GRAD_RTOL = 1e-3


def make_params(integrator=0, kernel=0, max_depth=128, srgb=True, hide_emitters=False, env=(1.0, 1.0, 1.0), image=None,
                rr_depth=-1, rr_seed=0, rr_skip=0):
    p = _cabi.vp_params()
    p.integrator, p.kernel = integrator, kernel
    p.max_depth = 0xFFFFFFFF if max_depth == -1 else max_depth
    p.srgb_primitives, p.hide_emitters = int(srgb), int(hide_emitters)
    p.t_cutoff, p.eps_advance = 0.01, 1e-4
    p.env[0], p.env[1], p.env[2] = env
    p.image_width, p.image_height = image if image else (0, 0)
    op = O.Params(integrator=integrator, kernel=kernel, max_depth=max_depth, srgb_primitives=srgb,
                  hide_emitters=hide_emitters, env=tuple(env), rr_depth=rr_depth, rr_seed=rr_seed, rr_skip=rr_skip)
    p.use_rr = int(op.use_rr)
    p.rr_depth = 0xFFFFFFFF if rr_depth == -1 else rr_depth
    p.rr_seed, p.rr_skip = rr_seed, rr_skip
    return p, op


def gpu_scene(cloud, attr=None, sh=True, build=True):
    acc = EllipsoidAccel()
    acc.set_primitives(torch.from_numpy(cloud.data), torch.from_numpy(cloud.opacities if attr is None else attr),
                       torch.from_numpy(cloud.sh_coeffs) if sh else None, cloud.extent)
    if build:
        acc.build()
    return acc


def oracle_scene(cloud, attr=None, sh=True, precision="f32"):
    return O.Scene(cloud.data, cloud.opacities if attr is None else attr, cloud.sh_coeffs if sh else None,
                   cloud.extent, precision=precision)


def linear_to_srgb(x):
    x = np.asarray(x, np.float64)
    return np.where(x <= 0.0031308, 12.92 * x, 1.055 * np.power(np.maximum(x, 0.0031308), 1 / 2.4) - 0.055)


def check_fragile_rays(ids_g, nh_g, rgb_g, res_orc, idx, replay, cap, srgb=True):
    """Rays whose GPU hit list differs from the oracle's (all of them outside the robust set).  Their colour is NOT
    free: (1) the GPU's own list, replayed with the oracle's arithmetic, must be a sequence of legal hits and must
    reproduce the GPU radiance to the contract tolerance; (2) both lists agree up to their first differing entry, so
    the two radiances can differ by at most the throughput left at that point times the largest colour -- asserted in
    the space the compositing happens in (sRGB when srgb_primitives)."""
    if len(idx) == 0:
        return 0.0
    osc, op, o, d, mt = replay
    mt_i = None if mt is None else mt[idx]
    cnt_g = np.minimum(nh_g[idx], cap).astype(np.uint32)
    rp = osc.replay(op, o[idx], d[idx], mt_i, ids_g[idx], cnt_g)
    assert rp["valid"].all(), f"{(~rp['valid']).sum()} fragile rays list a primitive the ray does not enter"
    ok = np.abs(rgb_g[idx] - rp["rgb"]) <= RGB_ATOL + RGB_RTOL * np.abs(rp["rgb"])
    assert ok.all(), f"fragile rays: radiance is not the one their own hit list gives (max diff {np.abs(rgb_g[idx] - rp['rgb']).max()})"
    ids_o = res_orc.hit_ids[idx][:, :cap]
    cnt_o = np.minimum(res_orc.nhits[idx], cap).astype(np.uint32)
    rp_o = osc.replay(op, o[idx], d[idx], mt_i, ids_o, cnt_o)
    differ = ids_g[idx] != ids_o
    first = np.where(differ.any(1), differ.argmax(1), cap - 1)
    beta_m = rp["hit_beta"][np.arange(len(idx)), np.minimum(first, rp["hit_beta"].shape[1] - 1)]
    beta_m = np.where(first >= cnt_g, rp["beta"].astype(np.float64), beta_m)   # GPU list ended first: what it had left
    cmax = np.maximum(rp["cmax"], rp_o["cmax"])
    bound = beta_m * np.maximum(cmax, 1e-6)
    a, b = (linear_to_srgb(rgb_g[idx]), linear_to_srgb(res_orc.rgb[idx])) if srgb else (rgb_g[idx], res_orc.rgb[idx])
    excess = np.abs(a - b).max(axis=1) - (bound * 1.001 + 5e-4)
    assert (excess <= 0).all(), f"fragile rays: radiance differs by more than the remaining throughput allows ({excess.max():.3e})"
    return float(np.abs(a - b).max())


def robust_mask(res_orc):
    """Rays whose oracle hit list is robust (see compare_forward); the others may legally differ in fp32."""
    frag = res_orc.fragility
    scale = np.maximum(1.0, np.nan_to_num(np.where(np.isfinite(res_orc.hit_t), res_orc.hit_t, 0.0).max(axis=1)))
    # column 3: the throughput came within 1e-5 (relative) of the termination threshold of rf:173-174 -- the list's
    # LENGTH is then decided by the last bits of the transmittances
    return (frag[:, 0] > 1e-5 * scale) & (frag[:, 1] > 2e-6 * scale) & (frag[:, 2] > 1e-4) & (frag[:, 3] > 1e-5)


def compare_forward(res_gpu, res_orc, cap, max_fragile_frac=5e-3, replay=None, srgb=True, ids_g=None, res_orc64=None):
    """Returns dict of stats; asserts the parity contract.

    Contract: for every ray whose hit list the oracle reports as robust (no two entries closer than 1e-5
    relative, no entry within 1e-6 of the epsilon cull, no |discriminant| below 1e-4, throughput never within 1e-5 of
    the termination threshold), the ID list must be
    IDENTICAL and radiance / transmittance within tolerance (with `res_orc64`, the float64 oracle's result on the same
    rays, a robust ray must also have the same list in both precisions).  Rays outside that set may differ and are counted;
    their fraction must stay below `max_fragile_frac`, and with `replay` = (oracle scene, oracle params, o, d, maxt)
    their radiance is bounded as well (check_fragile_rays)."""
    if ids_g is None:
        ids_g = res_gpu.hit_ids.t().contiguous().cpu().numpy()[:, :cap]
    nh_g = res_gpu.nhits.cpu().numpy().astype(np.int64)
    rgb_g = res_gpu.rgb.cpu().numpy()
    beta_g = res_gpu.beta.cpu().numpy() if res_gpu.beta is not None else None
    ids_o, nh_o = res_orc.hit_ids[:, :cap], res_orc.nhits.astype(np.int64)
    same = (ids_g == ids_o).all(axis=1) & (nh_g == nh_o)
    robust = robust_mask(res_orc)
    if res_orc64 is not None:
        # a ray whose list differs between the fp32 and the float64 build of the SAME loop is decided by rounding
        robust &= (res_orc.hit_ids[:, :cap] == res_orc64.hit_ids[:, :cap]).all(axis=1)
    bad = robust & ~same
    assert not bad.any(), f"{bad.sum()} robust rays have different hit lists, e.g. ray {np.flatnonzero(bad)[:5]}"
    n_diff = int((~same).sum())
    # how many rays may differ: a fixed small fraction, plus -- when the float64 oracle is at hand -- twice the number of
    # rays on which the fp32 and float64 builds of the oracle THEMSELVES disagree (long lists of sub-millimetre
    # primitives: at 200 hits per ray every tenth ray contains a hit that rounding decides)
    budget = max_fragile_frac * len(same) + 1
    if res_orc64 is not None:
        budget += 2 * int((~(res_orc.hit_ids[:, :cap] == res_orc64.hit_ids[:, :cap]).all(axis=1)).sum())
    assert n_diff <= budget, f"{n_diff} of {len(same)} rays differ (all fragile) -- more than the budget of {budget:.0f}"
    ok = np.abs(rgb_g - res_orc.rgb) <= RGB_ATOL + RGB_RTOL * np.abs(res_orc.rgb)
    assert ok[same].all(), f"radiance mismatch: max abs diff {np.abs(rgb_g - res_orc.rgb)[same].max()}"
    if beta_g is not None:
        okb = np.abs(beta_g - res_orc.beta) <= RGB_ATOL + RGB_RTOL * np.abs(res_orc.beta)
        assert okb[same].all(), f"transmittance mismatch: max abs diff {np.abs(beta_g - res_orc.beta)[same].max()}"
    frag_diff = None
    if replay is not None and n_diff:
        frag_diff = check_fragile_rays(ids_g, nh_g, rgb_g, res_orc, np.flatnonzero(~same), replay, cap, srgb)
    return {"rays": len(same), "identical": int(same.sum()), "fragile_diff": n_diff,
            "robust": int(robust.sum()), "max_rgb_diff": float(np.abs(rgb_g - res_orc.rgb)[same].max(initial=0.0)),
            "mean_hits": float(nh_o.mean()), "fragile_max_diff": frag_diff, "_same": same}


class GradientReference:
    """Reference for GRADIENTS: the float64 build of the oracle (the build tests/test_oracle_golden.py pins to the
    reference source's own outputs at 1e-7), with the fp32 build beside it as the measure of what fp32 can resolve.

    An fp32 evaluation of one hit -- the reference's Dr.Jit kernels, the fp32 oracle and the CUDA kernels alike --
    loses digits where the algorithm is ill-conditioned: the first hit of a camera ray is evaluated from an origin
    ~1e3 scales away (up to ~5e-4 relative error on that hit's terms), and the tomography line integral as written
    (common.py:199-206) cancels in fp32 (per-element errors of 1e-2 of the block's rms between the fp32 and the float64
    oracle).  So: the CUDA gradient must agree with float64 within 1e-3 (|ref| + block rms) -- the contract -- plus
    three times the fp32 oracle's OWN deviation from float64 on that element, the part no fp32 code can be held to."""

    def __init__(self, cloud, attr=None, sh=True):
        self.o32 = oracle_scene(cloud, attr=attr, sh=sh, precision="f32")
        self.o64 = oracle_scene(cloud, attr=attr, sh=sh, precision="f64")
        # What fp32 WORLD COORDINATES can resolve of a primitive: positions near it sit on a grid of 2^-23 |x|, and every
        # per-hit term depends on (p - c) / s.  For a primitive whose thinnest axis is 4e-4 (they exist in the 1M cloud)
        # that is 6e-4 of relative error on each of its terms, in ANY fp32 arithmetic -- the fp32 oracle's own per-ray
        # terms are off from float64 by up to 4 % there.  kappa = 4 * 2^-24 * max(|c|, 1) / min(s), per primitive.
        data = np.asarray(cloud.data, np.float64).reshape(-1, 10)
        self.kappa = 4.0 * 2.0 ** -24 * np.maximum(np.abs(data[:, :3]).max(axis=1), 1.0) / data[:, 3:6].min(axis=1)

    def same_lists(self, op, o, d, mt, ids_g, cap):
        ref64 = self.o64.forward(op, o, d, mt, cap=cap)
        return (np.asarray(ids_g)[:, :cap] == ref64.hit_ids[:, :cap]).all(axis=1)

    SUM_EPS = 5e-7 / 3.0     # x3 in grad_close: 8 units of fp32 round-off (2^-24) per unit of sum |term|

    def adjoint(self, op, o, d, dL, state, mt):
        """(float64 gradients, per-element noise).  noise = |fp32 oracle - float64 oracle| + (SUM_EPS + kappa / 3) *
        sum_hits |term|.  SUM_EPS bounds what an fp32 ACCUMULATION of the per-hit terms loses (the oracles accumulate in
        double; the reference's Dr.Jit scatter-add and the CUDA kernels accumulate in fp32, and with a random delta-L
        thousands of terms cancel to a small sum); kappa is the primitive's fp32 conditioning (see __init__), applied to
        terms in which the oracle's abs mode floors every |u_i| = |R^T (p - c)|_i / s_i at 1: the error of p - c is a few
        ulp of the WORLD coordinates, i.e. absolute, so a term that happens to be small because the ray passes the
        primitive's mid-plane (u_i ~ 0 along a thin axis) is as uncertain as it would be at |u_i| ~ 1 -- the
        observed |fp32 - float64| alone is one realisation of that noise and can be small by luck."""
        g64 = self.o64.adjoint(op, o, d, dL, state, mt)
        g32 = self.o32.adjoint(op, o, d, dL, state, mt)
        gabs = self.o64.adjoint(op, o, d, dL, state, mt, abs_terms=True)
        noise = []
        for a, b, c in zip(g32, g64, gabs):
            if a is None:
                noise.append(None)
                continue
            c = np.asarray(c, np.float64)
            k = self.kappa.reshape((-1,) + (1,) * (c.ndim - 1))
            noise.append(np.abs(np.asarray(a, np.float64) - np.asarray(b, np.float64)) + (self.SUM_EPS + k / 3.0) * c)
        return g64, tuple(noise)


def f64_reference(cloud, op, o, d, mt, ids_g, cap, attr=None, sh=True):
    """(GradientReference, mask of the rays whose float64 hit list equals the GPU's `ids_g`)."""
    ref = GradientReference(cloud, attr=attr, sh=sh)
    return ref, ref.same_lists(op, o, d, mt, ids_g, cap)


def grad_close(g_gpu, g_orc, rtol=GRAD_RTOL, what="", noise=None):
    """ELEMENTWISE agreement of a gradient block: |diff| <= rtol * |ref| + rtol * rms(ref) [+ 3 * noise].  The relative
    term is the contract's 1e-3; the rms term is the absolute floor for elements that are themselves sums with
    cancellation; `noise` (optional) is the fp32 oracle's own deviation from float64 on each element (see
    GradientReference).  Returns the largest |diff| / (|ref| + rms)."""
    g_gpu = np.asarray(g_gpu, np.float64).reshape(-1)
    g_orc = np.asarray(g_orc, np.float64).reshape(-1)
    assert g_gpu.shape == g_orc.shape, what
    rms = np.sqrt(np.mean(g_orc ** 2))
    if rms == 0:
        assert np.abs(g_gpu).max(initial=0.0) == 0, what
        return 0.0
    assert np.isfinite(g_gpu).all(), f"{what}: non-finite gradient"
    diff = np.abs(g_gpu - g_orc)
    ratio = diff / (np.abs(g_orc) + rms)
    allowed = rtol * (np.abs(g_orc) + rms)
    if noise is not None:
        allowed = allowed + 3.0 * np.asarray(noise, np.float64).reshape(-1)
    bad = diff > allowed
    # A per-hit derivative is discontinuous at the alpha clamp (0.9999) and at the edge of the Epanechnikov support: a
    # hit within rounding of either may fall on the other side in another arithmetic.  At most one element in a
    # million (and never more than 1e-2 of |ref| + rms) may be such a case.
    if 0 < bad.sum() <= max(1, int(1e-6 * bad.size)) and (ratio[bad] <= 1e-2).all():
        bad[:] = False
    if bad.any():
        worst = int(np.argmax(np.where(bad, diff / allowed, 0)))
        raise AssertionError(f"{what}: {int(bad.sum())} elements out of tolerance; worst element {worst}: got {g_gpu[worst]:.6e}, "
                             f"want {g_orc[worst]:.6e} (|diff| / (|ref| + rms) = {ratio[worst]:.3e}, rtol {rtol}, "
                             f"fp32 noise of the reference there {0.0 if noise is None else float(np.asarray(noise).reshape(-1)[worst]):.3e}, "
                             f"block rms {rms:.3e})")
    return float(ratio.max())


def check_gradients(got, want, noise, what, split_data=True):
    """(g_data, g_attr, g_sh) of the CUDA path against GradientReference.adjoint's (gradients, noise), block by block
    (centre / scale / quaternion separately when split_data)."""
    to_np = lambda t: t.detach().cpu().numpy() if hasattr(t, "detach") else np.asarray(t)
    gd, wd, nd = to_np(got[0]).reshape(-1, 10), np.asarray(want[0]).reshape(-1, 10), np.asarray(noise[0]).reshape(-1, 10)
    errs = {}
    blocks = (("center", slice(0, 3)), ("scale", slice(3, 6)), ("quat", slice(6, 10))) if split_data else (("data", slice(0, 10)),)
    for name, sl in blocks:
        errs[name] = grad_close(gd[:, sl], wd[:, sl], what=f"{what} d {name}", noise=nd[:, sl])
    errs["attr"] = grad_close(to_np(got[1]), want[1], what=f"{what} d attr", noise=noise[1])
    if len(got) > 2 and got[2] is not None and want[2] is not None:
        errs["sh"] = grad_close(to_np(got[2]), want[2], what=f"{what} d sh", noise=noise[2])
    return errs


def record_lists(rec, rays, cap):
    """Dense [len(rays), cap] -1 padded id lists + counts of the selected rays of a compressed-row hit record."""
    out = np.full((len(rays), cap), -1, np.int32)
    cnt = np.zeros(len(rays), np.int64)
    if rec.dense:
        sel = torch.as_tensor(np.asarray(list(rays), np.int64), device=rec.ids.device)
        ids = rec.ids[:, sel].t().cpu().numpy()
        cnt = rec.counts[sel].cpu().numpy().astype(np.int64)
        n = min(cap, ids.shape[1])
        out[:, :n] = np.where(np.arange(n)[None, :] < cnt[:, None], ids[:, :n], -1)
        return out, cnt
    off = rec.ray_offsets.cpu().numpy()
    ids = rec.ids.cpu().numpy()
    for k, r in enumerate(rays):
        a, b = off[r], off[r + 1]
        n = min(b - a, cap)
        out[k, :n] = ids[a:a + n]
        cnt[k] = b - a
    return out, cnt
