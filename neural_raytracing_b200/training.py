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
from typing import Callable, Iterable, Optional

import torch


class GraphedStep:
    def __init__(self, loss_fn: Callable[[], torch.Tensor], optimizer: torch.optim.Optimizer,
                 modules: Iterable[torch.nn.Module] = (), allreduce: Optional[Callable[[], None]] = None, warmup: int = 3):
        self.optimizer = optimizer
        self.allreduce = allreduce
        self._mlps = []
        for mod in modules:
            self._mlps += [m for m in mod.modules() if hasattr(m, "_pack_key")]
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            for _ in range(warmup):      # allocator / lazy-init warm-up outside the capture
                optimizer.zero_grad(set_to_none=True)
                loss_fn().backward()
                if allreduce is not None:
                    allreduce()
                optimizer.step()
        torch.cuda.current_stream().wait_stream(side)
        optimizer.zero_grad(set_to_none=True)
        self._invalidate()
        self.graph_fb = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph_fb):
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
        if self.allreduce is not None:
            self.allreduce()
        self.graph_opt.replay()
        self._invalidate()
        return self.loss
