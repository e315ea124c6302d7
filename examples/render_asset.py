#!/usr/bin/env python
"""Render the sensors of a saved Python asset -- counterpart of the reference's examples/render_asset.py
(asset_to_dict -> scale_films -> load_dict -> render per sensor); images go to <output>/<sensor>.npy.

    python examples/render_asset.py --asset /tmp/out/optimized_asset --output /tmp/renders --spp 16
"""
import argparse
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import volprim_balance_b200 as volprim  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument('--asset', required=True, help='folder written by volprim.io.dict_to_asset')
ap.add_argument('--output', required=True)
ap.add_argument('--spp', type=int, default=16)
ap.add_argument('--scale', type=float, default=1.0, help='film resolution scale (io.scale_films)')
args = ap.parse_args()

scene_dict = volprim.io.scale_films(volprim.io.asset_to_dict(args.asset), args.scale)
scene = volprim.load_dict(scene_dict)
os.makedirs(args.output, exist_ok=True)
names = [k for k, v in scene_dict.items() if isinstance(v, dict) and v.get('type') == 'perspective']
hosts = volprim.render_to_host(scene, spp=args.spp)          # image i is copied to the host while i + 1 is traced
for name, img in zip(names, hosts):
    np.save(os.path.join(args.output, f'{name}.npy'), img.numpy())
    print(f'{name}: {tuple(img.shape)} mean {float(img.mean()):.4f}')
