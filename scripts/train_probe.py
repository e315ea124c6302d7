#!/usr/bin/env python
"""Training-step probe (bench.py's train_step leg alone) for several all-reduce range counts in one launch:
python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 scripts/train_probe.py 2 3 4 6"""
import json
import os
import sys
import types

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
import volprim_balance_b200 as vp  # noqa: E402
from volprim_balance_b200 import parallel, training  # noqa: E402
from volprim_balance_b200.integrators.common import Ellipsoid  # noqa: E402


def main():
    world, rank, local = (int(os.environ.get(k, d)) for k, d in (("WORLD_SIZE", 1), ("RANK", 0), ("LOCAL_RANK", 0)))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    wl = bench.WORKLOADS["cfg2"]
    cloud = bench.build_cloud(wl)
    scene = bench.make_scene(vp, wl, cloud, dev)
    for chunks in [int(a) for a in sys.argv[1:]] or [4]:
        args = types.SimpleNamespace(train_chunks=chunks, train_rebuild="refit", train_rebuild_every=8, train_steps=8)
        t = bench.run_train_step(args, torch, dist, vp, training, parallel, Ellipsoid, wl, cloud, scene, dev, rank, world)
        if rank == 0:
            print(json.dumps({"n_gpus": world, "ranges": chunks, **{k: t[k] for k in ("ms_per_step", "exposed_allreduce_ms",
                                                                                      "optimizer_and_rebuild_ms")}}), flush=True)
        torch.cuda.empty_cache()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
