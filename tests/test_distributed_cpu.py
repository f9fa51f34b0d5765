"""world_size-2 gloo tests (CPU) of the multi-GPU host logic: ray sharding, image assembly, flat-bucket
gradient all-reduce, loss-denominator reduction.  The render function is a stand-in (the kernels need a
GPU); what is tested is that N-rank results equal the 1-rank results."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from neural_raytracing_b200 import distributed as D


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close(); return p


def test_shard_range_partitions_exactly():
    for n in (0, 1, 7, 64, 640000, 8294400):
        for world in (1, 2, 3, 4, 8):
            spans = [D.shard_range(n, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(spans[i][1] == spans[i + 1][0] for i in range(world - 1))
            sizes = [hi - lo for lo, hi in spans]
            assert max(sizes) - min(sizes) <= 1


def _fake_render(rays):
    return torch.stack([rays[:, 0] * 2 + rays[:, 3], rays[:, 1] - rays[:, 4], rays[:, 2] * rays[:, 5]], dim=-1)


def _worker(rank, world, port, R, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        g = torch.Generator().manual_seed(0)
        rays = torch.randn(R, 6, generator=g)
        img = D.render_sharded(_fake_render, rays)
        ok_img = torch.equal(img, _fake_render(rays))
        # camera-driven frame: every rank renders its rows of the window, the gathered image equals the whole frame
        nx, ny = (R % 37) + 3, 5

        def rows(x0, n):
            i = torch.arange(x0, x0 + n, dtype=torch.float)[None, :, None, None]
            j = torch.arange(ny, dtype=torch.float)[None, None, :, None]
            v = torch.arange(2, dtype=torch.float)[:, None, None, None]
            return torch.cat([i * 100 + j + 0 * v, v + 0 * i + 0 * j, i * j + v], dim=-1)
        ok_img = ok_img and torch.equal(D.render_camera_sharded(rows, nx), rows(0, nx))
        # gradients: each rank contributes the gradient of ITS slice of a sum-loss; the all-reduced
        # (summed) bucket must equal the single-process gradient over all rays
        lin = torch.nn.Linear(6, 3)
        with torch.no_grad():
            lin.weight.copy_(torch.arange(18.).reshape(3, 6) / 10); lin.bias.fill_(0.1)
        lo, hi = D.shard_range(R, rank, world)
        lin(rays[lo:hi]).square().sum().backward()
        D.allreduce_gradients(lin.parameters(), average=False)
        ref = torch.nn.Linear(6, 3)
        with torch.no_grad():
            ref.weight.copy_(torch.arange(18.).reshape(3, 6) / 10); ref.bias.fill_(0.1)
        ref(rays).square().sum().backward()
        ok_grad = torch.allclose(lin.weight.grad, ref.weight.grad, rtol=1e-5, atol=1e-5) and \
            torch.allclose(lin.bias.grad, ref.bias.grad, rtol=1e-5, atol=1e-5)
        # mean over a per-rank-varying count (eikonal-style)
        mask = rays[:, 0] > 0.3
        num = (rays[lo:hi, 1] ** 2)[mask[lo:hi]].sum()
        cnt = mask[lo:hi].sum()
        mean = D.allreduce_mean(num, cnt)
        ok_mean = torch.allclose(mean, (rays[:, 1] ** 2)[mask].mean(), rtol=1e-5)
        q.put((rank, bool(ok_img), bool(ok_grad), bool(ok_mean)))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("R", [1001, 64])
def test_two_rank_results_equal_single_rank(R):
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, R, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, ok_img, ok_grad, ok_mean in res:
        assert ok_img and ok_grad and ok_mean, (rank, ok_img, ok_grad, ok_mean)
