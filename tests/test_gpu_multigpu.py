"""GPU test (-m gpu, needs >= 2 GPUs, skipped otherwise) of the ray-sharded multi-GPU path WITH THE REAL KERNELS over
NCCL: the N-rank image equals the 1-rank image bit for bit (rays are independent), and the N-rank all-reduced MLP
weight gradients of a NeRFLE training step equal the 1-rank gradients of the whole batch up to fp32 summation order.
(The host-side sharding logic is also covered on CPU with gloo: tests/test_distributed_cpu.py.)"""
import os
import socket

import numpy as np
import pytest

import helpers
import synth

pytestmark = pytest.mark.gpu


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close(); return p


def _build(dev, tprec):
    import torch
    from neural_raytracing_b200 import config, training
    from neural_raytracing_b200.pathtracer.shapes.nerf import NeRFLE
    config.set_train_precision(tprec)
    n = NeRFLE(device=dev)
    synth.fill_module(n, 3)
    with torch.no_grad():
        n.first.out.bias[0] = 0.8
    n.far_jitter = torch.full((1,), 0.37, device=dev)      # the far-plane jitter of nerf.py:178, fixed for the comparison
    return n, training.FlatParameters([n.first, n.second])


def _worker(rank, world, port, q):
    import torch
    import torch.distributed as dist
    from neural_raytracing_b200 import distributed as D, ops
    from neural_raytracing_b200.pathtracer.lights import PointLights
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    try:
        # ---- render: every rank its slice, all_gather, vs the whole frame on one rank ----
        w1, w2 = helpers.nerfle_weights(False)
        m1, m2 = helpers.cuda_mlp(w1, device=dev), helpers.cuda_mlp(w2, device=dev)
        R = 20001                                               # uneven slices
        rays = torch.from_numpy(synth.camera_rays(11, R)).to(dev)
        ts = torch.linspace(0, 2.05, 64, device=dev)
        code = torch.tensor([[0.4, 1.0, 0.3]], device=dev)
        img = D.render_sharded(lambda r: ops.nerfle_render(m1, m2, r, ts, code, prec="f16"), rays)
        whole = ops.nerfle_render(m1, m2, rays, ts, code, prec="f16")
        ok_img = bool(torch.equal(img, whole))
        # ---- the same from a camera (f4): every rank generates the rays of ITS rows on its own device ----
        c2w, focal = synth.nerf_cameras(2, 75, device=dev)

        def rows(x0, n):
            cam = ops.CameraDesc(ops.CAM_NERF, c2w, None, focal=focal, size=75, x0=x0, y0=0, nx=n, ny=75)
            return ops.nerfle_render_camera(m1, m2, cam, ts, code.expand(2, 3).contiguous(), prec="f16")
        frame = D.render_camera_sharded(rows, 75)              # 75 rows over the ranks: uneven
        ok_img = ok_img and bool(torch.equal(frame, rows(0, 75))) and tuple(frame.shape) == (2, 75, 75, 1, 3)
        # ---- training step: flat gradient, all-reduced over the ranks, vs the whole batch on one rank ----
        res = {}
        for tprec in ("f32", "f16"):
            n, flat = _build(dev, tprec)
            Rt = 4096
            all_rays = torch.from_numpy(synth.camera_rays(5, Rt)).to(dev)
            lights = PointLights(device=dev, location=torch.tensor([[0.4, 1.0, 0.3]], device=dev), scale=10)

            def grad_of(lo, hi):
                flat.zero_grad()
                r = all_rays[lo:hi].reshape(1, hi - lo, 1, 1, 6)
                loss = (n(r, lights) - 0.5).square().sum() / (Rt * 3)
                loss.backward()
                return flat.grad.clone(), float(loss.detach())
            g_all, loss_all = grad_of(0, Rt)
            lo, hi = D.shard_range(Rt, rank, world)
            g_loc, loss_loc = grad_of(lo, hi)
            flat.allreduce(average=False)                       # the product's exchange: one NCCL call on flat.grad
            g_sum = flat.grad.clone()
            lt = torch.tensor([loss_loc], device=dev, dtype=torch.float64)
            dist.all_reduce(lt)
            a, b = g_sum.double(), g_all.double()
            res[tprec] = (float((a @ b) / (a.norm() * b.norm())), float((a - b).abs().max() / b.abs().max()),
                          abs(float(lt) - loss_all) / abs(loss_all), float(b.norm()))
        q.put((rank, ok_img, res))
    finally:
        from neural_raytracing_b200 import config
        config.set_train_precision("f32")
        dist.destroy_process_group()


def test_two_rank_image_and_gradients_equal_single_rank():
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs >= 2 GPUs")
    import torch.multiprocessing as mp
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=600) for _ in range(world)]
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    for rank, ok_img, r in res:
        assert ok_img, "rank %d: sharded image differs from the single-GPU image" % rank
        for tprec, (cos, rel, dloss, norm) in r.items():
            assert norm > 0
            # fp32 sums in a different order: per-rank partial sums + NCCL vs one kernel's atomics
            assert cos > 0.999999 and rel < 1e-4 and dloss < 1e-5, (rank, tprec, cos, rel, dloss)
    print("2-rank parity:", res)
