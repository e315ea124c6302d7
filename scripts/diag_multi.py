"""Diagnostic for tests/test_gpu_multi.py: torchrun --nproc-per-node 2 scripts/diag_multi.py"""
import os, sys
import numpy as np, torch, torch.distributed as dist
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import volprim_balance_b200 as vp
from volprim_balance_b200 import parallel, synthetic, training
from tests.test_gpu_multi import _scene_and_opt
rank, ws = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(rank)
dist.init_process_group("nccl", device_id=torch.device("cuda", rank))
n, n_views, W, H = 30000, 4, 128, 64
scene, sensors, targets, opt = _scene_and_opt(vp, synthetic, n, n_views, W, H)
step = training.RefineStep(scene, sensors, targets, opt, n_chunks=3)
print(rank, "ranges", step.ranges, flush=True)
for rep in range(3):
    step.views, step.world = list(range(n_views)), 1
    g_single = step.accumulate_gradients()[0].flat.clone()
    per_view = []
    for v in range(n_views):
        step.views = [v]
        per_view.append(step.accumulate_gradients()[0].flat.clone())
    step.views, step.world = parallel.shard_views(n_views, rank, ws), ws
    bucket, works, _ = step.accumulate_gradients()
    g_local_after = None
    for w in works:
        w.wait()
    torch.cuda.synchronize()
    g_multi = bucket.flat.clone()
    rms = float(g_single.pow(2).mean().sqrt())
    rel = (g_multi - g_single).abs() / (g_single.abs() + rms)
    s4 = sum(per_view)
    rel2 = (s4 - g_single).abs() / (g_single.abs() + rms)
    others = [torch.empty_like(g_single) for _ in range(ws)]
    dist.all_gather(others, g_single)
    others_m = [torch.empty_like(g_multi) for _ in range(ws)]
    dist.all_gather(others_m, g_multi)
    print(rank, "rep", rep, "err multi-vs-single %.3e at %d" % (float(rel.max()), int(rel.argmax())), "sum-of-views-vs-single %.3e" % float(rel2.max()),
          "g_single equal across ranks", bool(torch.equal(others[0], others[1])), "max diff %.3e" % float((others[0]-others[1]).abs().max()),
          "g_multi equal across ranks", bool(torch.equal(others_m[0], others_m[1])), "n bad", int((rel > 1e-4).sum()), flush=True)
dist.destroy_process_group()
