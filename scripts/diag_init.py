"""Single-GPU reproduction of step (1) of tests/test_gpu_multi.py for compute-sanitizer:
compute-sanitizer --tool initcheck python scripts/diag_init.py"""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import volprim_balance_b200 as vp
from volprim_balance_b200 import parallel, synthetic, training
from tests.test_gpu_multi import _scene_and_opt
dev = int(os.environ.get("DIAG_DEVICE", "0"))
torch.cuda.set_device(dev)
n, n_views, W, H = 30000, 4, 128, 64
scene, sensors, targets, opt = _scene_and_opt(vp, synthetic, n, n_views, W, H)
step = training.RefineStep(scene, sensors, targets, opt, n_chunks=3)
step.ranges = parallel.chunk_ranges(n, 3, taper=True)
shf = step.shape.attributes['sh_coeffs'].numel() // n
step.bucket = parallel.GradientBucket(n, shf, step.shape.device, ranges=step.ranges)
outs = []
for rep in range(3):
    outs.append(step.accumulate_gradients()[0].flat.clone())
torch.cuda.synchronize()
rms = float(outs[0].pow(2).mean().sqrt())
print("non-finite gradient elements per pass:", [int((~torch.isfinite(o)).sum()) for o in outs], flush=True)
for k in (1, 2):
    rel = (outs[k] - outs[0]).abs() / (outs[0].abs() + rms)
    print("rep", k, "vs 0: max rel %.3e, n off %d" % (float(rel.max()), int((rel > 1e-4).sum())), flush=True)
