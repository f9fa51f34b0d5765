"""CUDA-graph capture of a training step (the reference's loops -- nerfle.py:104-120,
training_utils.py:211-260 -- are launch-bound at the 4,096-ray batches they use: ~150 small launches per step).

`GraphedStep` captures   zero_grad -> loss_fn() -> loss.backward()   in one CUDA graph and   optimizer.step()   in a
second one; the gradient all-reduce of the multi-GPU path runs between the two (eagerly: one NCCL call on a flat
bucket).  Everything the library enqueues (weight re-packing, the tensor-core forward / dgrad / wgrad kernels,
compositing) goes to the capturing stream, so a replay re-runs it on the current weights.

Rules (the usual ones for whole-network capture): static input tensors (update them in place between replays), an
optimizer constructed with `capturable=True`, no host synchronisation inside `loss_fn` (the NeRFLE path has none; the
SDF path syncs on `out_active.any()` like the reference and cannot be captured).  Python-side randomness inside
`loss_fn` is frozen at capture time: NeRFLE's far-plane jitter is therefore read from the device tensor
`NeRFLE.far_jitter` when it is set (see shapes/nerf.py)."""
from typing import Callable, Iterable, List, Optional, Sequence

import torch
import torch.distributed as dist


class FlatParameters:
    """ONE flat fp32 parameter buffer and ONE flat gradient buffer for a set of SkipConnMLPs (+ any other tensors).

    The reference optimises ~30 small tensors per MLP; a data-parallel step then pays ~30 gradient copies into a
    bucket, the all-reduce, ~30 copies back and a multi-tensor optimizer (12 % of the 8-GPU NeRFLE step).
    Here every nn.Linear of the given MLPs is re-homed into a slice of `flat` IN THE PACKED-F32 LAYOUT OF THE C ABI
    (W^T [K][N] then bias [N], evaluation order: `lin.weight` becomes the transposed view of its slice), and `.grad` of
    every parameter is the matching view of `grad`.  Consequences:
      * the fused kernels read the fp32 masters in place (SkipConnMLP.packed() copies nothing) and ACCUMULATE their
        packed-f32 weight gradients straight into `grad` (no per-tensor scatter);
      * the data-parallel exchange is `dist.all_reduce(self.grad)`: one NCCL call on the buffer the kernels wrote;
      * the optimizer sees ONE parameter (`self.param`): a fused AdamW step is one kernel.
    state_dict()/load_state_dict() of the modules keep working (parameters are ordinary nn.Parameters whose storage
    happens to be shared); `zero_grad()` must be this object's (set_to_none would drop the views)."""

    def __init__(self, mlps: Sequence[torch.nn.Module], others: Iterable[torch.Tensor] = ()):
        self.mlps = list(mlps)
        others = [t for t in others]
        dev = next(self.mlps[0].parameters()).device if self.mlps else others[0].device
        sizes = []
        for m in self.mlps:
            sizes.append(sum(l.weight.numel() + l.bias.numel() for l in m._linears()))
        pad = lambda k: (k + 63) // 64 * 64      # every MLP's slice starts 256-byte aligned (the kernels need 16)
        n = sum(pad(k) for k in sizes) + sum(t.numel() for t in others)
        self.flat = torch.zeros(n, dtype=torch.float32, device=dev)
        self.grad = torch.zeros_like(self.flat)
        self.slices: List[slice] = []
        off = 0
        with torch.no_grad():
            for m, sz in zip(self.mlps, sizes):
                base = off
                for lin in m._linears():
                    N, K = lin.weight.shape
                    w = self.flat[off:off + K * N].view(K, N)
                    w.copy_(lin.weight.detach().t())
                    lin.weight.data = w.t()
                    lin.weight.grad = self.grad[off:off + K * N].view(K, N).t()
                    off += K * N
                    b = self.flat[off:off + N]
                    b.copy_(lin.bias.detach())
                    lin.bias.data = b
                    lin.bias.grad = self.grad[off:off + N]
                    off += N
                assert off - base == sz
                self.slices.append(slice(base, off))
                m._flat_view = self.flat[base:off]
                m._flat_grad = self.grad[base:off]
                m._pack_key = None
                m._packed = None
                off = base + pad(sz)
            for t in others:
                k = t.numel()
                v = self.flat[off:off + k].view(t.shape)
                v.copy_(t.detach())
                t.data = v
                t.grad = self.grad[off:off + k].view(t.shape)
                off += k
        self.param = torch.nn.Parameter(self.flat)     # shares storage with `flat`: what the optimizer steps
        self.param.grad = self.grad

    def zero_grad(self):
        self.grad.zero_()

    def allreduce(self, average: bool = False, group=None):
        """Sum (or mean) of the flat gradient over the ranks: one NCCL call, no staging copies."""
        if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
            dist.all_reduce(self.grad, op=dist.ReduceOp.SUM, group=group)
            if average:
                self.grad.div_(dist.get_world_size(group))


class GraphedStep:
    def __init__(self, loss_fn: Callable[[], torch.Tensor], optimizer: torch.optim.Optimizer,
                 modules: Iterable[torch.nn.Module] = (), allreduce: Optional[Callable[[], None]] = None, warmup: int = 3,
                 flat: Optional[FlatParameters] = None, single_graph: bool = False):
        """flat: the FlatParameters the optimizer steps (its zero_grad replaces optimizer.zero_grad).
        single_graph: capture zero_grad -> forward -> backward -> allreduce -> optimizer.step() in ONE graph (NCCL
        collectives are capturable); default: two graphs with the all-reduce launched eagerly between them."""
        self.optimizer = optimizer
        self.allreduce = allreduce
        self.flat = flat
        self._mlps = []
        for mod in modules:
            self._mlps += [m for m in mod.modules() if hasattr(m, "_pack_key")]

        def zero():
            if flat is not None:
                flat.zero_grad()
            else:
                optimizer.zero_grad(set_to_none=True)
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            for _ in range(warmup):      # allocator / lazy-init warm-up outside the capture
                zero()
                loss_fn().backward()
                if allreduce is not None:
                    allreduce()
                optimizer.step()
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        zero()
        self._invalidate()
        self.graph_opt = None
        self.graph_fb = torch.cuda.CUDAGraph()
        if single_graph:
            with torch.cuda.graph(self.graph_fb):
                if flat is not None:
                    flat.zero_grad()
                self.loss = loss_fn()
                self.loss.backward()
                if allreduce is not None:
                    allreduce()
                optimizer.step()
            self.allreduce = None
        else:
            with torch.cuda.graph(self.graph_fb):
                if flat is not None:
                    flat.zero_grad()
                self.loss = loss_fn()
                self.loss.backward()
            self.graph_opt = torch.cuda.CUDAGraph()
            with torch.cuda.graph(self.graph_opt):
                optimizer.step()
        self._invalidate()

    def _invalidate(self):
        # replays change the weights without bumping tensor versions: drop the packed-parameter caches so that eager
        # calls made after (or between) replays re-pack from the current weights
        for m in self._mlps:
            m._pack_key = None

    def __call__(self) -> torch.Tensor:
        """One training step; returns the (static) loss tensor of this step."""
        self.graph_fb.replay()
        if self.graph_opt is not None:
            if self.allreduce is not None:
                self.allreduce()
            self.graph_opt.replay()
        self._invalidate()
        return self.loss
