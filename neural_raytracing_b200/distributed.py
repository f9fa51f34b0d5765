"""Ray sharding across the GPUs of one box.

Rays are independent, so rendering needs no data-path collective: rank g renders a contiguous slice
and the image is assembled with one all_gather (or left rank-local).  Training replicates the small
MLPs and all-reduces ONE flat fp32 gradient bucket per step (0.66 MB for NeRFLE ... 13 MB for the DTU
configuration) over NCCL/NVLink; loss terms that are not plain sums over rays must all-reduce their
numerators/denominators before the non-linearity."""
from typing import Callable, Iterable, Optional, Tuple

import torch
import torch.distributed as dist


def shard_range(n: int, rank: int, world: int) -> Tuple[int, int]:
    """[lo, hi) of the contiguous slice of n units owned by `rank`; sizes differ by at most one."""
    base, rem = divmod(n, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def _world(group=None):
    if dist.is_available() and dist.is_initialized():
        return dist.get_rank(group), dist.get_world_size(group)
    return 0, 1


def render_sharded(render_fn: Callable[[torch.Tensor], torch.Tensor], rays: torch.Tensor, gather: bool = True,
                   group=None) -> torch.Tensor:
    """Every rank holds the same rays [R,6]; rank g renders its slice with render_fn and, if gather,
    all ranks receive the full [R,C] result (uneven slices are padded for the collective)."""
    rank, world = _world(group)
    R = rays.shape[0]
    lo, hi = shard_range(R, rank, world)
    local = render_fn(rays[lo:hi].contiguous())
    if world == 1 or not gather:
        return local
    C = local.shape[-1]
    per = (R + world - 1) // world
    padded = torch.zeros((per, C), dtype=local.dtype, device=local.device)
    padded[: hi - lo] = local
    out = torch.empty((world * per, C), dtype=local.dtype, device=local.device)
    dist.all_gather_into_tensor(out, padded, group=group)
    pieces = []
    for g in range(world):
        glo, ghi = shard_range(R, g, world)
        pieces.append(out[g * per: g * per + (ghi - glo)])
    return torch.cat(pieces, dim=0)


def render_camera_sharded(render_rows_fn: Callable[[int, int], torch.Tensor], nx: int, gather: bool = True,
                          group=None) -> torch.Tensor:
    """Camera-driven multi-GPU frame (f4): rank g renders the pixel ROWS [lo, hi) of the frame's first image axis with
    `render_rows_fn(lo, hi - lo) -> [n_views, hi - lo, ny, C]` (e.g. `ops.nerfle_render_camera` on a window of the camera
    descriptor: every rank generates its own rays on its own device, no ray array exists anywhere) and, if gather, all
    ranks receive the whole [n_views, nx, ny, C] image.  Rays are independent: the result equals the 1-rank frame bit for
    bit when the sample distances are deterministic (the reference's shared `ts`, or stratified / hierarchical sampling with
    jitter_seed = 0); the library's stratified and pixel jitter are keyed by a ray's index inside its CALL, so with a
    non-zero seed an N-rank frame is a different (equally distributed) draw than the 1-rank frame."""
    rank, world = _world(group)
    lo, hi = shard_range(nx, rank, world)
    local = render_rows_fn(lo, hi - lo)
    if world == 1 or not gather:
        return local
    per = (nx + world - 1) // world
    n_views = local.shape[0]
    tail = tuple(local.shape[2:])
    padded = torch.zeros((per, n_views) + tail, dtype=local.dtype, device=local.device)     # rows first: contiguous shards
    padded[: hi - lo] = local.transpose(0, 1)
    out = torch.empty((world * per, n_views) + tail, dtype=local.dtype, device=local.device)
    dist.all_gather_into_tensor(out, padded, group=group)
    pieces = []
    for g in range(world):
        glo, ghi = shard_range(nx, g, world)
        pieces.append(out[g * per: g * per + (ghi - glo)])
    return torch.cat(pieces, dim=0).transpose(0, 1).contiguous()


def allreduce_gradients(params: Iterable[torch.Tensor], average: bool = True, group=None) -> Optional[torch.Tensor]:
    """One flat-bucket all-reduce of every .grad (missing grads count as zero); writes the reduced
    gradients back.  Returns the bucket (for inspection)."""
    rank, world = _world(group)
    ps = [p for p in params if p.requires_grad]
    if not ps:
        return None
    dev = next((p.grad.device for p in ps if p.grad is not None), ps[0].device)
    flat = torch.cat([(p.grad if p.grad is not None else torch.zeros_like(p)).reshape(-1).to(dev, torch.float32)
                      for p in ps])
    if world > 1:
        dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=group)
        if average:
            flat /= world
    off = 0
    for p in ps:
        n = p.numel()
        g = flat[off:off + n].reshape(p.shape).to(p.device, p.dtype)
        if p.grad is None:
            p.grad = g.clone()
        else:
            p.grad.copy_(g)
        off += n
    return flat


def allreduce_mean(numerator: torch.Tensor, count: torch.Tensor, group=None) -> torch.Tensor:
    """Mean over a per-rank-varying number of elements (e.g. eikonal loss over the hits of this rank's
    rays): all-reduce numerator and count, then divide -- equals the single-GPU value."""
    rank, world = _world(group)
    t = torch.stack([numerator.reshape(()).float(), count.reshape(()).float()])
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)
    return t[0] / t[1].clamp(min=1)
