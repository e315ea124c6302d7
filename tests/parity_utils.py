"""Shared helpers of the GPU parity tests: build the same cloud on both sides, compare hit lists with the
documented exclusion (exact-depth ties / entries at the epsilon cull / grazing entries)."""
import numpy as np
import torch

from oracle import oracle as O
from volprim_balance_b200 import _cabi, synthetic
from volprim_balance_b200.accel import EllipsoidAccel

# tolerances stated by BASELINE.json north_star
RGB_ATOL, RGB_RTOL = 1e-4, 1e-3
GRAD_RTOL = 1e-3


def make_params(integrator=0, kernel=0, max_depth=128, srgb=True, hide_emitters=False, env=(1.0, 1.0, 1.0), image=None):
    p = _cabi.vp_params()
    p.integrator, p.kernel = integrator, kernel
    p.max_depth = 0xFFFFFFFF if max_depth == -1 else max_depth
    p.srgb_primitives, p.hide_emitters = int(srgb), int(hide_emitters)
    p.t_cutoff, p.eps_advance = 0.01, 1e-4
    p.env[0], p.env[1], p.env[2] = env
    p.image_width, p.image_height = image if image else (0, 0)
    op = O.Params(integrator=integrator, kernel=kernel, max_depth=max_depth, srgb_primitives=srgb,
                  hide_emitters=hide_emitters, env=tuple(env))
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


def compare_forward(res_gpu, res_orc, cap, max_fragile_frac=5e-3):
    """Returns dict of stats; asserts the parity contract.

    Contract: for every ray whose hit list the oracle reports as robust (no two entries closer than 1e-5
    relative, no entry within 1e-6 of the epsilon cull, no |discriminant| below 1e-4), the ID list must be
    IDENTICAL and radiance / transmittance within tolerance.  Rays outside that set may differ and are counted;
    their fraction must stay below `max_fragile_frac`."""
    ids_g = res_gpu.hit_ids.t().contiguous().cpu().numpy()[:, :cap]
    nh_g = res_gpu.nhits.cpu().numpy().astype(np.int64)
    rgb_g, beta_g = res_gpu.rgb.cpu().numpy(), res_gpu.beta.cpu().numpy()
    ids_o, nh_o = res_orc.hit_ids[:, :cap], res_orc.nhits.astype(np.int64)
    same = (ids_g == ids_o).all(axis=1) & (nh_g == nh_o)
    frag = res_orc.fragility
    scale = np.maximum(1.0, np.nan_to_num(np.where(np.isfinite(res_orc.hit_t), res_orc.hit_t, 0.0).max(axis=1)))
    robust = (frag[:, 0] > 1e-5 * scale) & (frag[:, 1] > 2e-6 * scale) & (frag[:, 2] > 1e-4)
    bad = robust & ~same
    assert not bad.any(), f"{bad.sum()} robust rays have different hit lists, e.g. ray {np.flatnonzero(bad)[:5]}"
    n_diff = int((~same).sum())
    assert n_diff <= max_fragile_frac * len(same) + 1, f"{n_diff} of {len(same)} rays differ (all fragile) -- too many"
    ok = np.abs(rgb_g - res_orc.rgb) <= RGB_ATOL + RGB_RTOL * np.abs(res_orc.rgb)
    assert ok[same].all(), f"radiance mismatch: max abs diff {np.abs(rgb_g - res_orc.rgb)[same].max()}"
    okb = np.abs(beta_g - res_orc.beta) <= RGB_ATOL + RGB_RTOL * np.abs(res_orc.beta)
    assert okb[same].all(), f"transmittance mismatch: max abs diff {np.abs(beta_g - res_orc.beta)[same].max()}"
    return {"rays": len(same), "identical": int(same.sum()), "fragile_diff": n_diff,
            "robust": int(robust.sum()), "max_rgb_diff": float(np.abs(rgb_g - res_orc.rgb)[same].max(initial=0.0)),
            "mean_hits": float(nh_o.mean())}


def grad_close(g_gpu, g_orc, rtol=GRAD_RTOL, what=""):
    """Relative L2 / max-norm agreement of a gradient block (atomics reorder fp32 sums, so compare with a
    norm-relative tolerance as well as elementwise where the value is not tiny)."""
    g_gpu = np.asarray(g_gpu, np.float64).reshape(-1)
    g_orc = np.asarray(g_orc, np.float64).reshape(-1)
    scale = np.abs(g_orc).max()
    if scale == 0:
        assert np.abs(g_gpu).max() == 0, what
        return 0.0
    err = np.abs(g_gpu - g_orc).max() / scale
    assert err <= rtol, f"{what}: max |diff| / max |ref| = {err:.3e} > {rtol}"
    return err
