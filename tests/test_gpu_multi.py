"""GPU, two or more devices: the training step with view sharding + NCCL all-reduce reproduces the single-GPU gradient,
the replicas stay bit-identical after the optimiser step + LBVH rebuild, and image-tile sharding assembles the
single-GPU image.  Skipped on a one-GPU box (run with `gpurun --gpus 2`)."""
import os
import socket

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _scene_and_opt(vp, synthetic, n, n_views, W, H):
    from volprim_balance_b200.integrators.common import Ellipsoid
    cloud = synthetic.make_cloud(n, synthetic.sigma0_for_hits(n, 30.0), seed=3, sh_degree=3)
    rng = np.random.default_rng(7)
    start = cloud.data.copy()
    start[:, :3] += rng.normal(0, 2e-3, (n, 3)).astype(np.float32)

    def prim(data, op, sh):
        return {"type": "ellipsoidsmesh", "centers": data[:, :3], "scales": data[:, 3:6], "quaternions": data[:, 6:],
                "opacities": op[:, None], "sh_coeffs": sh, "extent": 3.0}

    integ = {"type": "volprim_rf", "max_depth": 64, "rr_depth": 64}
    sensors = []
    for i in range(n_views):
        c = synthetic.ring_camera(i, n_views, W, H)
        sensors.append(vp.load_dict({"type": "perspective", "fov": c.fov_x_deg, "fov_axis": "x", "to_world": vp.Transform4f(c.to_world),
                                     "near_clip": c.near_clip, "far_clip": c.far_clip,
                                     "film": {"type": "hdrfilm", "width": W, "height": H, "rfilter": {"type": "box"}}}))
    ref_scene = vp.load_dict({"type": "scene", "integrator": integ, "primitives": prim(cloud.data, cloud.opacities, cloud.sh_coeffs)})
    targets = {i: vp.render(ref_scene, sensor=s, spp=1, jitter=False) for i, s in enumerate(sensors)}
    scene = vp.load_dict({"type": "scene", "integrator": integ,
                          "primitives": prim(start, np.clip(cloud.opacities * 0.8, 1e-4, 1 - 1e-4), cloud.sh_coeffs * 0.9)})
    params = vp.traverse(scene)
    opt = vp.optimizers.BoundedAdam()
    e = Ellipsoid.unravel(params["primitives.data"])
    opt["centers"], opt["scales"], opt["quats"] = e.center, e.scale, e.quat
    opt["opacities"], opt["sh_coeffs"] = params["primitives.opacities"], params["primitives.sh_coeffs"]
    opt.set_learning_rate({"centers": 1e-4, "scales": 1e-4, "quats": 1e-4, "opacities": 1e-2, "sh_coeffs": 1e-3})
    opt.set_bounds("scales", lower=1e-6)
    opt.set_bounds("opacities", lower=1e-6, upper=1.0 - 1e-6)
    return scene, sensors, targets, opt


def _worker(rank, ws, port, q):
    try:
        os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(ws))
        import torch.distributed as dist
        import volprim_balance_b200 as vp
        from volprim_balance_b200 import parallel, synthetic, training
        torch.cuda.set_device(rank)
        dist.init_process_group("nccl", rank=rank, world_size=ws, device_id=torch.device("cuda", rank))
        n, n_views, W, H = 30000, 4, 128, 64
        scene, sensors, targets, opt = _scene_and_opt(vp, synthetic, n, n_views, W, H)
        step = training.RefineStep(scene, sensors, targets, opt, n_chunks=3)
        # (1) single-GPU gradient of ALL views, computed locally without any collective
        step.views, step.world = list(range(n_views)), 1
        g_single = step.accumulate_gradients()[0].flat.clone()
        # (2) view-sharded gradient + chunked all-reduce overlapped with the last view's accumulation
        step.views, step.world = parallel.shard_views(n_views, rank, ws), ws
        bucket, works, _ = step.accumulate_gradients()
        for w in works:
            w.wait()
        torch.cuda.synchronize()
        g_multi = bucket.flat.clone()
        rms = float(g_single.pow(2).mean().sqrt())
        rel = (g_multi - g_single).abs() / (g_single.abs() + rms)
        err = float(rel.max())
        where = int(rel.argmax())
        both_s, both_m = [torch.empty_like(g_single) for _ in range(ws)], [torch.empty_like(g_multi) for _ in range(ws)]
        dist.all_gather(both_s, g_single)
        dist.all_gather(both_m, g_multi)
        bad = torch.nonzero(rel > 1e-4).flatten()
        detail = (f"single-GPU gradients of the ranks differ by {float((both_s[0] - both_s[1]).abs().max()):.3e}, reduced gradients by "
                  f"{float((both_m[0] - both_m[1]).abs().max()):.3e}; off elements span {int(bad.min()) if len(bad) else -1}.."
                  f"{int(bad.max()) if len(bad) else -1}; element {where} of {rel.numel()}: sharded {float(g_multi[where]):.6e} single {float(g_single[where]):.6e}, "
                  f"{int((rel > 1e-4).sum())} elements off, ranges {step.ranges}, record usable {step.record.usable()}, "
                  f"estimate {step.shape.accel().hits_per_ray_estimate:.1f}")
        # (3) full steps: replicas must stay BIT-identical (same reduced gradient, same Adam, deterministic rebuild)
        losses = []
        for _ in range(3):
            loss, sq = step.step()
            losses.append(float(loss))
        flat = torch.cat([opt[k].detach().reshape(-1) for k in ("centers", "scales", "quats", "opacities", "sh_coeffs")])
        others = [torch.empty_like(flat) for _ in range(ws)]
        dist.all_gather(others, flat)
        identical = all(torch.equal(o, flat) for o in others)
        # (4) image-tile sharding of ONE view: bands rendered per rank, gathered on rank 0
        full = vp.render(scene, sensor=sensors[1], spp=1, jitter=False)
        tiled = parallel.render_tiles(scene, sensors[1], vp.render, spp=1, jitter=False)
        tiles_ok = True if rank != 0 else bool(torch.equal(tiled, full))
        q.put((rank, err, identical, losses, tiles_ok, dict(step.timing, detail=detail)))
        dist.destroy_process_group()
    except Exception as e:  # noqa: BLE001
        import traceback
        q.put((rank, "error", traceback.format_exc(), None, None, None))


@pytest.mark.timeout(600)
def test_sharded_training_step_equals_single_gpu_and_replicas_stay_identical():
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    import torch.multiprocessing as mp
    ws = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, ws, port, q)) for r in range(ws)]
    for p in procs:
        p.start()
    res = sorted((q.get(timeout=500) for _ in range(ws)), key=lambda r: r[0])
    for p in procs:
        p.join(timeout=60)
    for r in res:
        assert r[1] != "error", r[2]
    for rank, err, identical, losses, tiles_ok, timing in res:
        assert err < 1e-4, f"rank {rank}: reduced gradient differs from the single-GPU gradient ({err:.2e}; {timing.get('detail')})"
        assert identical, "replicas diverged after optimiser step + rebuild"
        assert tiles_ok
        assert losses[-1] < losses[0]
    assert res[0][3] == res[1][3]
    print("2-GPU step timing", res[0][5], "losses", res[0][3])


def test_training_step_equals_autograd_path_single_gpu():
    """RefineStep (direct kernel calls, chunked finish) == render() + loss.backward() + BoundedAdam, on one GPU."""
    import volprim_balance_b200 as vp
    from volprim_balance_b200 import synthetic, training
    from volprim_balance_b200.integrators.common import Ellipsoid
    n, n_views, W, H = 20000, 3, 96, 64
    scene, sensors, targets, opt = _scene_and_opt(vp, synthetic, n, n_views, W, H)
    step = training.RefineStep(scene, sensors, targets, opt, n_chunks=4)
    bucket, works, (loss, sq, _flags) = step.accumulate_gradients()
    torch.cuda.synchronize()
    # autograd path on an identical copy
    scene2, sensors2, targets2, opt2 = _scene_and_opt(vp, synthetic, n, n_views, W, H)
    params = vp.traverse(scene2)
    params["primitives.data"] = Ellipsoid.ravel(opt2["centers"], opt2["scales"], opt2["quats"])
    params["primitives.opacities"], params["primitives.sh_coeffs"] = opt2["opacities"], opt2["sh_coeffs"]
    params.update()
    n_pix = n_views * W * H * 3
    total = 0.0
    for i, s in enumerate(sensors2):
        img = vp.render(scene2, params, sensor=s, spp=1, jitter=False)
        l = (targets2[i] - img).abs().sum() / n_pix
        l.backward()
        total += float(l)
    assert abs(total - float(loss)) < 2e-6
    g_data, g_attr, g_sh = bucket.gather()
    g = g_data.view(-1, 10)
    from tests.parity_utils import grad_close
    grad_close(g[:, 0:3].cpu().numpy(), opt2["centers"].grad.cpu().numpy(), rtol=1e-4, what="centers")
    grad_close(g[:, 3:6].cpu().numpy(), opt2["scales"].grad.cpu().numpy(), rtol=1e-4, what="scales")
    grad_close(g[:, 6:10].cpu().numpy(), opt2["quats"].grad.cpu().numpy(), rtol=1e-4, what="quats")
    grad_close(g_attr.cpu().numpy(), opt2["opacities"].grad.cpu().numpy(), rtol=1e-4, what="opacities")
    grad_close(g_sh.cpu().numpy(), opt2["sh_coeffs"].grad.cpu().numpy(), rtol=1e-4, what="sh")
    # and a few full steps reduce the loss
    first = float(step.step()[0])
    for _ in range(4):
        last = float(step.step()[0])
    assert last < first
