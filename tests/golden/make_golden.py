"""Generates tests/golden/*.npz by EXECUTING the reference's own source
(/root/reference/volprim/integrators/{common,volprim_rf,volprim_tomography}.py, unmodified) over the torch
stand-in for drjit / mitsuba in refshim.py.  Run in the authoring container (needs /root/reference):

    python tests/golden/make_golden.py

Fixtures (float64 so that formula differences, not rounding, are what a comparison sees):
  kernels.npz          GaussianKernel.eval / EpanechnikovKernel.eval / density_integral (full range) /
                       ray_ellipsoid_intersection on random ray-ellipsoid pairs
  sample_rf_*.npz      VolumetricPrimitiveRadianceFieldIntegrator.sample: Primal (L, hit sequence) and Backward
                       (gradients of primitives.data / opacities / sh_coeffs by autograd through the reference code)
  sample_tomo_*.npz    same for VolumetricPrimitiveTomographyIntegrator (sigma_t)
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
import refshim as R  # noqa: E402
from volprim_balance_b200 import synthetic  # noqa: E402

dr, mi, common, rf_mod, tomo_mod = R.load_reference()
T = lambda a: torch.as_tensor(np.asarray(a), dtype=R.DTYPE)


def vec(a):
    a = T(a)
    return R.Vec([R.Arr(a[:, i]) for i in range(a.shape[1])], n=a.shape[1])


def make_ellipsoid(rec, extent=3.0):
    rec = T(rec)
    q = vec(rec[:, 6:10])
    return common.Ellipsoid(vec(rec[:, 0:3]), vec(rec[:, 3:6]), q, dr.quat_to_matrix(q, size=3),
                            R.Float(torch.full((rec.shape[0],), extent, dtype=R.DTYPE)))


def gen_kernels():
    rng = np.random.default_rng(0)
    n = 256
    rec = np.concatenate([rng.normal(0, 0.3, (n, 3)), np.exp(rng.normal(np.log(0.1), 0.5, (n, 3))),
                          rng.normal(size=(n, 4))], 1)
    rec[:, 6:10] /= np.linalg.norm(rec[:, 6:10], axis=1, keepdims=True)
    rec[::7, 6:10] *= 1.004  # un-normalised quaternions (quirk Q6)
    o = rng.normal(0, 0.4, (n, 3)) + np.array([0, 0, -2.0])
    d = rec[:, 0:3] + rng.normal(0, 0.15, (n, 3)) - o
    d /= np.linalg.norm(d, axis=1, keepdims=True)
    p = rec[:, 0:3] + rng.normal(0, 0.15, (n, 3))
    out = {'rec': rec, 'o': o, 'd': d, 'p': p}
    ray = mi.Ray3f(vec(o), vec(d), R.Float(torch.full((n,), 1e30, dtype=R.DTYPE)))
    act = R.Bool(torch.ones(n, dtype=torch.bool))
    for extent in (3.0, 1.0):
        e = make_ellipsoid(rec, extent)
        valid, tn, tf = common.ray_ellipsoid_intersection(ray, e, act)
        out[f'isect_valid_e{extent:g}'] = valid.t.numpy()
        out[f'isect_near_e{extent:g}'] = tn.t.numpy()
        out[f'isect_far_e{extent:g}'] = tf.t.numpy()
        for name in ('gaussian', 'epanechnikov'):
            k = common.Kernel.factory({'kernel_type': name, 'kernel_full_range': True, 'kernel_normalized': False})
            out[f'{name}_density_integral_e{extent:g}'] = k.density_integral(ray, e, None, None, act).t.numpy()
    e = make_ellipsoid(rec, 3.0)
    for name in ('gaussian', 'epanechnikov'):
        k = common.Kernel.factory({'kernel_type': name})
        out[f'{name}_eval'] = k.eval(vec(p), e, act).t.numpy()
    np.savez_compressed(os.path.join(HERE, 'kernels.npz'), **out)
    print('kernels.npz', {k: v.shape for k, v in out.items() if k.startswith('gauss')})


def run_sample(kind, kernel, n=160, sigma=0.13, deg=3, W=16, H=12, view=1, max_depth=64, srgb=True, extent=3.0,
               seed=5, tag='', hide_emitters=False, finite_maxt=False, unnormalised_quats=False):
    cloud = synthetic.make_cloud(n, sigma, seed=seed, sh_degree=deg)
    if unnormalised_quats:      # quirk Q6: the maths never normalises (smoke.ply stores norms up to 1.0047)
        cloud.data[::3, 6:10] *= 1.0047
        cloud.data[1::5, 6:10] *= 0.996
    cam = synthetic.ring_camera(view, 8, W, H)
    o, d, mt = synthetic.camera_rays(cam)
    o, d, mt = o.astype(np.float64), d.astype(np.float64), mt.astype(np.float64)
    rng = np.random.default_rng(seed + 100)
    if finite_maxt:             # every third ray stops inside the cloud (maxt applies to the re-based origin too)
        mt[::3] = rng.uniform(0.6, 1.6, mt[::3].shape)
    data = torch.tensor(cloud.data.astype(np.float64), requires_grad=True)
    attrs = {}
    if kind == 'rf':
        attrs['opacities'] = torch.tensor(cloud.opacities.astype(np.float64)[:, None], requires_grad=True)
        attrs['sh_coeffs'] = torch.tensor(cloud.sh_coeffs.astype(np.float64), requires_grad=True)
        props = R.Properties({'max_depth': max_depth, 'rr_depth': max_depth, 'kernel_type': kernel, 'srgb_primitives': srgb})
        integ = rf_mod.VolumetricPrimitiveRadianceFieldIntegrator(props)
        attr_name = 'opacities'
    else:
        sig = rng.uniform(0.002, 0.05, n)
        attrs['sigma_t'] = torch.tensor(sig[:, None], requires_grad=True)
        props = R.Properties({'max_depth': max_depth, 'kernel_type': kernel, 'hide_emitters': hide_emitters})
        integ = tomo_mod.VolumetricPrimitiveTomographyIntegrator(props)
        attr_name = 'sigma_t'
    shape = R.RefShape(data, attrs, extent)
    env = (1.0, 0.6, 0.3)
    scene = R.RefScene(shape, env)
    Rn = o.shape[0]

    def ray():
        return mi.Ray3f(vec(o), vec(d), R.Float(T(mt)))
    active = R.Bool(torch.ones(Rn, dtype=torch.bool))
    with torch.no_grad():
        L, _, _, state = integ.sample(dr.ADMode.Primal, scene, R._Sampler(), ray(), None, 0.0, active)
    hits = torch.stack(scene.hit_log, 1).numpy()               # [R, iterations], -1 when invalid / inactive
    L_np = np.stack([c.t.numpy() for c in L.c], -1)
    dL = rng.normal(size=(Rn, 3))
    dL[::9] = 0.0                                              # rays with zero gradient are skipped (rf:111-112)
    scene.hit_log = []
    with torch.no_grad():
        integ.sample(dr.ADMode.Backward, scene, R._Sampler(), ray(), vec(dL), state, R.Bool(torch.ones(Rn, dtype=torch.bool)))
    out = {'data': cloud.data.astype(np.float64), 'attr': attrs[attr_name].detach().numpy()[:, 0], 'o': o, 'd': d,
           'maxt': mt, 'L': L_np, 'hits': hits, 'dL': dL, 'extent': extent, 'max_depth': max_depth, 'srgb': srgb,
           'env': np.array(env), 'kernel': kernel, 'hide_emitters': hide_emitters,
           'g_data': data.grad.numpy() if data.grad is not None else np.zeros_like(cloud.data, dtype=np.float64),
           'g_attr': attrs[attr_name].grad.numpy()[:, 0] if attrs[attr_name].grad is not None else np.zeros(n)}
    if kind == 'rf':
        out['sh'] = cloud.sh_coeffs.astype(np.float64)
        g = attrs['sh_coeffs'].grad
        out['g_sh'] = g.numpy() if g is not None else np.zeros_like(out['sh'])
    name = f'sample_{kind}_{kernel}{tag}.npz'
    np.savez_compressed(os.path.join(HERE, name), **out)
    nh = (hits >= 0).sum(1)
    print(name, 'rays', Rn, 'mean hits', nh.mean(), 'max', nh.max(), '|g_data|', np.abs(out['g_data']).sum())


if __name__ == '__main__':
    gen_kernels()
    run_sample('rf', 'gaussian')
    run_sample('rf', 'epanechnikov')
    run_sample('rf', 'gaussian', deg=1, srgb=False, max_depth=5, tag='_deg1_depth5')
    run_sample('tomo', 'gaussian', max_depth=-1)
    run_sample('tomo', 'epanechnikov', extent=1.0, max_depth=-1, tag='_extent1')
    run_sample('tomo', 'epanechnikov', extent=3.0, max_depth=7, tag='_extent3_depth7')
    run_sample('tomo', 'gaussian', max_depth=-1, hide_emitters=True, finite_maxt=True, seed=8, tag='_hide_maxt')
    run_sample('rf', 'gaussian', deg=2, unnormalised_quats=True, finite_maxt=True, seed=9, tag='_deg2_unnorm_maxt')
    run_sample('rf', 'epanechnikov', deg=0, max_depth=3, seed=10, tag='_deg0_depth3')
