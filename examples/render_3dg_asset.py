#!/usr/bin/env python
"""Render the views of a 3DG asset with `volprim_rf` and pull every image to the host (counterpart of the reference's
examples/render_3dg_asset.py).

    python examples/render_3dg_asset.py --ply point_cloud.ply --cameras cameras.json --output out --all
    python examples/render_3dg_asset.py                      # synthetic asset, written and re-read as PLY + JSON

Same scene description as the reference: `volprim_rf` integrator (max_depth, rr_depth, kernel_type), one
`ellipsoidsmesh` shape loaded from the PLY, one `perspective` sensor per camera of the 3DGS `cameras.json`.  The views go
through `render_to_host`, so the device-to-host copy of one image overlaps the trace of the next; images are written as
little-endian PFM (no OpenEXR writer in this environment).  Without --ply a synthetic cloud is exported with
`ellipsoid_dict_to_ply` / `JSONCameraSpecsIO.write` first and then loaded back, so the file formats are on the path.
"""
import argparse
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import volprim_balance_b200 as volprim  # noqa: E402
from volprim_balance_b200 import synthetic  # noqa: E402

ap = argparse.ArgumentParser(description='Render 3DG asset')
ap.add_argument('--ply', type=str, default=None, help='3DG PLY file (default: a synthetic asset)')
ap.add_argument('--cameras', type=str, default=None, help='3DGS cameras.json')
ap.add_argument('--output', type=str, default='output', help='output folder')
ap.add_argument('--cam_index', type=int, default=0, help='camera to render')
ap.add_argument('--all', action='store_true', help='render every camera instead of --cam_index')
ap.add_argument('--cam_scale', type=float, default=1.0, help='scale of the camera resolution')
ap.add_argument('--spp', type=int, default=2)
ap.add_argument('--max_depth', type=int, default=128)
ap.add_argument('--rr_depth', type=int, default=128)
ap.add_argument('--kernel', type=str, default='gaussian')
ap.add_argument('--primitives', type=int, default=200_000, help='size of the synthetic asset')
args = ap.parse_args()
os.makedirs(args.output, exist_ok=True)

if args.ply is None:
    n = args.primitives
    cloud = synthetic.make_cloud(n, synthetic.sigma0_for_hits(n, 40.0), seed=11)
    args.ply = os.path.join(args.output, 'synthetic.ply')
    args.cameras = os.path.join(args.output, 'cameras.json')
    volprim.io.ellipsoid_dict_to_ply({'centers': cloud.data[:, :3], 'scales': cloud.data[:, 3:6],
                                      'quaternions': cloud.data[:, 6:], 'opacities': cloud.opacities[:, None],
                                      'sh_coeffs': cloud.sh_coeffs}, ['opacities', 'sh_coeffs'], args.ply)
    specs = []
    for i in range(8):
        cam = synthetic.ring_camera(i, 8, 960, 540)
        specs.append(volprim.cameras.CameraSpecs(f'view_{i:02d}', cam.width, cam.height, volprim.Transform4f(cam.to_world),
                                                 fov=cam.fov_x_deg))
    volprim.cameras.JSONCameraSpecsIO.write(specs, args.cameras)
    print(f'synthetic asset: {n} primitives -> {args.ply}, {len(specs)} cameras -> {args.cameras}')

scene_dict = {
    'type': 'scene',
    'integrator': {'type': 'volprim_rf', 'max_depth': args.max_depth, 'rr_depth': args.rr_depth, 'kernel_type': args.kernel},
    'primitives': {'type': 'ellipsoidsmesh', 'filename': args.ply},
}
cam_specs = volprim.cameras.JSONCameraSpecsIO.load(args.cameras)
for spec in cam_specs:
    scene_dict[spec.name] = spec.to_dict(args.cam_scale)
t0 = time.perf_counter()
scene = volprim.load_dict(scene_dict)
torch.cuda.synchronize()
print(f'scene loaded and LBVH built in {1e3 * (time.perf_counter() - t0):.1f} ms')


def write_pfm(path, img):
    h, w, _ = img.shape
    with open(path, 'wb') as fh:
        fh.write(f'PF\n{w} {h}\n-1.0\n'.encode())
        fh.write(np.ascontiguousarray(img[::-1], dtype='<f4').tobytes())


views = list(range(len(cam_specs))) if args.all else [args.cam_index]
written = []


def on_image(i, host):
    path = os.path.join(args.output, f'{cam_specs[views[i]].name}.pfm')
    write_pfm(path, host.numpy())
    written.append(path)


volprim.render_to_host(scene, sensors=views[:1], spp=args.spp)          # warm-up (first launch, allocator)
t0 = time.perf_counter()
volprim.render_to_host(scene, sensors=views, spp=args.spp, on_image=on_image)
dt = time.perf_counter() - t0
rays = sum(int(cam_specs[v].width * args.cam_scale) * int(cam_specs[v].height * args.cam_scale) * args.spp for v in views)
print(f'rendered {len(views)} view(s), {rays / 1e6:.2f} Mrays in {1e3 * dt:.1f} ms ({rays / dt / 1e6:.1f} Mrays/s incl. host '
      f'copies and file writes); wrote {", ".join(written[:3])}{" ..." if len(written) > 3 else ""}')
