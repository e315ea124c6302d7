#!/usr/bin/env python
"""Fit a 3-D grid of volumetric primitives to views of a density volume with `volprim_tomography` -- counterpart of the
reference's examples/optimize_volume.py (setup :124-194, loop :232-249, pruning :255-262).

The reference renders its target views with Mitsuba's `prbvolpath` from a `.vol` grid (a large blob missing from the
repository); here the targets are `volprim_tomography` renders of a primitive cloud -- `--ply FILE` (e.g. the reference's
resources/smoke.ply) or, by default, a synthetic puff of 835 ellipsoids -- so the loop runs without Mitsuba.  Everything
else follows the reference: batch sensor of `cam_count` views, grid of volprim_count^3 Gaussians with the reference's
initial values, BoundedAdam with its learning rates and bounds, l1 loss, psnr, pruning by sigma_t and scale.

    python examples/optimize_volume.py --output /tmp/out --iterations 64
"""
import argparse
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import volprim_balance_b200 as volprim  # noqa: E402
from volprim_balance_b200.integrators.common import Ellipsoid, EllipsoidsFactory  # noqa: E402

ap = argparse.ArgumentParser(description='Optimize volumetric primitives from views of a volume')
ap.add_argument('--output', type=str, default=None, help='Path to the result output folder')
ap.add_argument('--ply', type=str, default=None, help='Ellipsoid PLY (sigma_t attribute) that defines the target volume')
ap.add_argument('--cam_count', type=int, default=8)
ap.add_argument('--cam_res', type=int, default=256)
ap.add_argument('--ref_spp', type=int, default=32)
ap.add_argument('--opt_spp', type=int, default=1)
ap.add_argument('--grad_spp', type=int, default=1)
ap.add_argument('--max_depth', type=int, default=-1)
ap.add_argument('--kernel', type=str, default='gaussian')
ap.add_argument('--iterations', type=int, default=64)
ap.add_argument('--volprim_count', type=int, default=16)
ap.add_argument('--init_albedo', type=float, default=0.9)
ap.add_argument('--init_sigmat', type=float, default=0.0001)
ap.add_argument('--no_prune', action='store_true')
ap.add_argument('--global_lr', type=float, default=1.0)
ap.add_argument('--centers_lr', type=float, default=0.015)
ap.add_argument('--scales_lr', type=float, default=0.0001)
ap.add_argument('--quats_lr', type=float, default=0.0001)
ap.add_argument('--sigmat_lr', type=float, default=0.0001)
args = ap.parse_args()

# ---- cameras on a ring around the volume (reference :70-87) ----------------------------------------------------------
rng = np.random.default_rng(0)
cams = {}
for i in range(args.cam_count):
    angle = 2.0 * np.pi * i / args.cam_count
    origin = [4.0 * np.sin(angle), rng.uniform(-1.0, 1.0), 4.0 * np.cos(angle)]
    cams[f'cam_{i:04d}'] = {'type': 'perspective', 'fov': 40.0, 'fov_axis': 'x',
                            'to_world': volprim.Transform4f().look_at(origin, [0, 0, 0], [0, 1, 0]),
                            'film': {'type': 'hdrfilm', 'width': args.cam_res, 'height': args.cam_res,
                                     'rfilter': {'type': 'tent'}, 'pixel_format': 'rgb'}}
batch = volprim.load_dict({'type': 'batch', 'film': {'type': 'hdrfilm', 'width': args.cam_res * args.cam_count,
                                                      'height': args.cam_res, 'filter': {'type': 'tent'}}, **cams})
integrator = {'type': 'volprim_tomography', 'max_depth': args.max_depth, 'kernel_type': args.kernel}

# ---- target views ------------------------------------------------------------------------------------------------------
if args.ply:
    target_prims = {'type': 'ellipsoidsmesh', 'filename': args.ply, 'extent': 3.0}
else:
    m = 835
    c = rng.normal(0.0, 0.35, (m, 3)) * np.array([1.0, 1.4, 1.0])
    target_prims = {'type': 'ellipsoidsmesh', 'centers': c.astype(np.float32),
                    'scales': np.exp(rng.normal(-3.6, 0.2, (m, 3))).astype(np.float32),
                    'quaternions': np.tile(np.array([0, 0, 0, 1.0], np.float32), (m, 1)),
                    'sigma_t': rng.uniform(2e-4, 8e-4, (m, 1)).astype(np.float32), 'extent': 3.0}
ref_scene = volprim.load_dict({'type': 'scene', 'integrator': integrator, 'primitives': target_prims,
                               'environment': {'type': 'constant'}})
with torch.no_grad():
    ref_image = volprim.render(ref_scene, sensor=batch, spp=args.ref_spp, seed=12345)
del ref_scene

# ---- the grid of primitives to optimise (reference :128-160) --------------------------------------------------------------
factory = EllipsoidsFactory()
delta = 1.0 / args.volprim_count
for x in range(args.volprim_count):
    for y in range(args.volprim_count):
        for z in range(args.volprim_count):
            factory.add(mean=2.0 * delta * np.array([x, y, z], np.float32) - 1, scale=delta / 2, sigmat=args.init_sigmat,
                        albedo=args.init_albedo)
centers, scales, quaternions, sigmats, albedos = factory.build()
scene = volprim.load_dict({'type': 'scene', 'integrator': integrator,
                           'primitives': {'type': 'ellipsoidsmesh', 'centers': centers, 'scales': scales,
                                          'quaternions': quaternions, 'sigma_t': sigmats, 'albedo': albedos, 'extent': 3.0},
                           'environment': {'type': 'constant'}, **cams})      # the cameras travel with the exported asset
params = volprim.traverse(scene)
key_data, key_sigmat = 'primitives.data', 'primitives.sigma_t'
opt = volprim.optimizers.BoundedAdam()
e = Ellipsoid.unravel(params[key_data])
opt['centers'], opt['scales'], opt['quats'], opt['sigmat'] = e.center, e.scale, e.quat, params[key_sigmat]
opt.set_learning_rate({'centers': args.global_lr * args.centers_lr, 'scales': args.global_lr * args.scales_lr,
                       'quats': args.global_lr * args.quats_lr, 'sigmat': args.global_lr * args.sigmat_lr})
opt.set_bounds('scales', lower=1e-6)
opt.set_bounds('sigmat', lower=1e-8, upper=1e-3)


def update_params():
    params[key_data] = Ellipsoid.ravel(opt['centers'], opt['scales'], opt['quats'])
    params[key_sigmat] = opt['sigmat']
    params.update()


update_params()
losses, psnrs = [], []
for it in range(args.iterations):
    opt.zero_grad()
    # the same jitter for the primal and the gradient pass: the backward pass replays the primal's hit records
    image = volprim.render(scene, params, sensor=batch, spp=args.opt_spp, spp_grad=args.grad_spp, seed=it, seed_grad=it)
    loss = volprim.optimizers.l1(ref_image, image)
    psnr = volprim.optimizers.psnr(ref_image, image.detach())
    loss.backward()
    opt.step()
    update_params()
    losses.append(float(loss))
    psnrs.append(float(psnr))
    print(f'-- step {it + 1} / {args.iterations} | psnr={psnrs[-1]:.04f} | loss={losses[-1]:.04f}', flush=True)
print('Done with optimization')

if not args.no_prune:       # reference :255-262
    valid = (opt['sigmat'].detach().reshape(-1) > 1e-6) & (opt['scales'].detach() > 1e-4).all(dim=1)
    print(f"Pruning {int((~valid).sum())} volumetric primitives out of {valid.numel()}")
    params[key_data] = Ellipsoid.ravel(opt['centers'].detach()[valid], opt['scales'].detach()[valid], opt['quats'].detach()[valid])
    params[key_sigmat] = opt['sigmat'].detach().reshape(-1)[valid]
    if 'primitives.albedo' in params:
        params['primitives.albedo'] = params['primitives.albedo'].reshape(valid.numel(), -1)[valid].reshape(-1)
    params.update()
if args.output:
    os.makedirs(args.output, exist_ok=True)
    with torch.no_grad():
        final = volprim.render(scene, sensor=batch, spp=args.ref_spp, seed=12345)
    np.save(os.path.join(args.output, 'final.npy'), final.cpu().numpy())
    np.save(os.path.join(args.output, 'reference.npy'), ref_image.cpu().numpy())
    volprim.io.dict_to_asset(volprim.io.object_to_dict(scene), os.path.join(args.output, 'optimized_asset'))
    print(f"psnr {psnrs[0]:.2f} -> {psnrs[-1]:.2f}; images and asset written to {args.output}")
