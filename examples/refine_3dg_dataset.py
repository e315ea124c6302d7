#!/usr/bin/env python
"""Multi-view refinement of a 3DG-style primitive cloud with `volprim_rf` (counterpart of the reference's
examples/refine_3dg_dataset.py, cfg 4 of BASELINE.md): forward + PRB adjoint per view, L1 loss, BoundedAdam with the
reference's learning rates and bounds, LBVH refit/rebuild after every step.

Single GPU:      python examples/refine_3dg_dataset.py --iterations 20
Several GPUs:    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 \
                     examples/refine_3dg_dataset.py --iterations 20
Views are sharded over the ranks, primitives replicated, ONE packed gradient all-reduce per step, then the identical
optimiser step on every rank.  Input: --ply/--cameras (3DGS PLY + cameras.json) or, by default, a synthetic cloud whose
target images come from a perturbed copy (there are no datasets in this environment).
"""
import argparse
import os
import sys
import time

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import volprim_balance_b200 as volprim  # noqa: E402
from volprim_balance_b200 import parallel, synthetic  # noqa: E402
from volprim_balance_b200.integrators.common import Ellipsoid  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument('--ply'); ap.add_argument('--cameras')
ap.add_argument('--targets', help='.npy file [views, H, W, 3] with the target images of the --ply dataset (linear RGB)')
ap.add_argument('--fused_step', action='store_true', help='drive the kernels through volprim.training.RefineStep: gradient all-reduce '
                'cut into primitive ranges and overlapped with the last view\'s accumulation (same gradients as loss.backward())')
ap.add_argument('--primitives', type=int, default=200_000)
ap.add_argument('--cam_count', type=int, default=8)
ap.add_argument('--width', type=int, default=640); ap.add_argument('--height', type=int, default=360)
ap.add_argument('--iterations', type=int, default=20)
ap.add_argument('--max_depth', type=int, default=128); ap.add_argument('--rr_depth', type=int, default=128)
ap.add_argument('--kernel', default='gaussian')
ap.add_argument('--global_lr', type=float, default=1.0)
ap.add_argument('--centers_lr', type=float, default=1e-4); ap.add_argument('--scales_lr', type=float, default=1e-4)
ap.add_argument('--quats_lr', type=float, default=1e-4); ap.add_argument('--opacities_lr', type=float, default=1e-2)
ap.add_argument('--sh_coeffs_lr', type=float, default=1e-3)
ap.add_argument('--refit', action='store_true', help='keep the LBVH topology between steps (vp_refit) instead of rebuilding')
ap.add_argument('--output', default=None)
args = ap.parse_args()

world = int(os.environ.get('WORLD_SIZE', '1'))
rank = int(os.environ.get('RANK', '0'))
local = int(os.environ.get('LOCAL_RANK', '0'))
torch.cuda.set_device(local)
if world > 1:
    dist.init_process_group('nccl', device_id=torch.device('cuda', local))

# ---- scene ---------------------------------------------------------------------------------------------------------
if args.ply:
    prim = {'type': 'ellipsoidsmesh', 'filename': args.ply}
    specs = volprim.cameras.JSONCameraSpecsIO.load(args.cameras)
    idx = list(range(0, len(specs), max(1, len(specs) // args.cam_count)))[:args.cam_count]
    sensors = [specs[i].to_dict() for i in idx]
    target_prim = None
else:
    n = args.primitives
    cloud = synthetic.make_cloud(n, synthetic.sigma0_for_hits(n, 50.0), seed=3)
    rng = np.random.default_rng(7)
    start = cloud.data.copy()
    start[:, :3] += rng.normal(0, 2e-3, (n, 3)).astype(np.float32)
    prim = {'type': 'ellipsoidsmesh', 'centers': start[:, :3], 'scales': start[:, 3:6], 'quaternions': start[:, 6:],
            'opacities': np.clip(cloud.opacities * 0.8, 1e-4, 1 - 1e-4)[:, None],
            'sh_coeffs': cloud.sh_coeffs * 0.9, 'extent': 3.0}
    target_prim = {'type': 'ellipsoidsmesh', 'centers': cloud.data[:, :3], 'scales': cloud.data[:, 3:6],
                   'quaternions': cloud.data[:, 6:], 'opacities': cloud.opacities[:, None], 'sh_coeffs': cloud.sh_coeffs,
                   'extent': 3.0}
    sensors = []
    for i in range(args.cam_count):
        c = synthetic.ring_camera(i, args.cam_count, args.width, args.height)
        sensors.append({'type': 'perspective', 'fov': c.fov_x_deg, 'fov_axis': 'x', 'to_world': volprim.Transform4f(c.to_world),
                        'near_clip': c.near_clip, 'far_clip': c.far_clip,
                        'film': {'type': 'hdrfilm', 'width': c.width, 'height': c.height, 'rfilter': {'type': 'box'}}})

integrator = {'type': 'volprim_rf', 'max_depth': args.max_depth, 'rr_depth': args.rr_depth, 'kernel_type': args.kernel}
scene = volprim.load_dict({'type': 'scene', 'integrator': integrator, 'primitives': prim})
sensor_objs = [volprim.load_dict(s) for s in sensors]
mine = parallel.shard_views(len(sensor_objs), rank, world)
if target_prim is not None:
    ref_scene = volprim.load_dict({'type': 'scene', 'integrator': integrator, 'primitives': target_prim})
    refs = {i: volprim.render(ref_scene, sensor=sensor_objs[i], spp=1, jitter=False) for i in mine}
    del ref_scene
else:
    if not args.targets:
        raise SystemExit('--ply needs --targets FILE.npy ([views, H, W, 3], the images of the selected cameras)')
    stack = np.load(args.targets)
    refs = {i: torch.from_numpy(np.ascontiguousarray(stack[i], dtype=np.float32)).cuda() for i in mine}
if args.refit:
    scene.ellipsoids().rebuild_policy = 'refit'

# ---- optimiser (reference examples/refine_3dg_dataset.py:131-159) --------------------------------------------------
params = volprim.traverse(scene)
key_data, key_op, key_sh = 'primitives.data', 'primitives.opacities', 'primitives.sh_coeffs'
opt = volprim.optimizers.BoundedAdam()
e = Ellipsoid.unravel(params[key_data])
opt['centers'], opt['scales'], opt['quats'] = e.center, e.scale, e.quat
opt['opacities'], opt['sh_coeffs'] = params[key_op], params[key_sh]
opt.set_learning_rate({'centers': args.global_lr * args.centers_lr, 'scales': args.global_lr * args.scales_lr,
                       'quats': args.global_lr * args.quats_lr, 'opacities': args.global_lr * args.opacities_lr,
                       'sh_coeffs': args.global_lr * args.sh_coeffs_lr})
opt.set_bounds('scales', lower=1e-6)
opt.set_bounds('opacities', lower=1e-6, upper=1.0 - 1e-6)


def update_params():
    params[key_data] = Ellipsoid.ravel(opt['centers'], opt['scales'], opt['quats'])
    params[key_op], params[key_sh] = opt['opacities'], opt['sh_coeffs']
    params.update()


update_params()
n_pix = len(sensor_objs) * sensor_objs[0].width * sensor_objs[0].height * 3
if args.fused_step:
    from volprim_balance_b200 import training
    step = training.RefineStep(scene, sensor_objs, refs, opt, rebuild='refit' if args.refit else 'rebuild')
    for it in range(args.iterations):
        t0 = time.time()
        loss, sq = step.step()
        if rank == 0:
            psnr = 10 * np.log10(1.0 / max(float(sq), 1e-20))
            print(f'-- step {it + 1} / {args.iterations} | psnr={psnr:.04f} | loss={float(loss):.06f} | '
                  f'{1e3 * (time.time() - t0):.1f} ms (exposed all-reduce {step.timing["exposed_allreduce_ms"]:.2f} ms)', flush=True)
    args.iterations = 0
for it in range(args.iterations):
    t0 = time.time()
    opt.zero_grad()
    loss_local = torch.zeros((), device='cuda')
    sq_local = torch.zeros((), device='cuda')
    for i in mine:
        img = volprim.render(scene, params, sensor=sensor_objs[i], spp=1, seed=it, jitter=False)
        l = (refs[i] - img).abs().sum() / n_pix          # this view's share of l1 over the batch film
        l.backward()
        loss_local += l.detach()
        sq_local += ((refs[i] - img.detach()) ** 2).sum() / n_pix
    # primitive gradients: ONE all-reduce over the packed buffer; params[...] grads chain into opt[...] via ravel
    grads = {k: (v.grad if v.grad is not None else torch.zeros_like(v)) for k, v in opt.items()}
    grads = parallel.allreduce_gradients(grads)
    for k, v in opt.items():
        v.grad = grads[k]
    if world > 1:
        stats = torch.stack([loss_local, sq_local])
        dist.all_reduce(stats)
        loss_local, sq_local = stats[0], stats[1]
    opt.step()
    update_params()
    torch.cuda.synchronize()
    if rank == 0:
        psnr = 10 * np.log10(1.0 / max(float(sq_local), 1e-20))
        print(f'-- step {it + 1} / {args.iterations} | psnr={psnr:.04f} | loss={float(loss_local):.06f} | {1e3 * (time.time() - t0):.1f} ms', flush=True)

if rank == 0 and args.output:
    volprim.io.dict_to_asset(volprim.io.object_to_dict(scene), os.path.join(args.output, 'optimized_asset'))
if world > 1:
    dist.destroy_process_group()
