"""Shading frame and interaction records (pytorch3d/pathtracer/interaction.py)."""
from dataclasses import dataclass

import torch
import torch.nn.functional as F

from .. import ops


def _fused_ok(*ts):
    """CUDA fp32 tensors and no autograd graph needed -> the elementwise CUDA kernels can be used."""
    if not all(t.is_cuda and t.dtype == torch.float32 for t in ts):
        return False
    return not (torch.is_grad_enabled() and any(t.requires_grad for t in ts))


def coordinate_system(n):
    """Branch-free orthonormal frame around n (interaction.py:9-27), returned as columns
    [s, t, n] of a [...,3,3] tensor.  Keeps the reference's 1e-6 / 1e-7 guards and the three
    re-normalisations."""
    if _fused_ok(n):
        return ops.shading_frame(n)
    n = F.normalize(n, eps=1e-7, dim=-1)
    x, y, z = n[..., 0:1], n[..., 1:2], n[..., 2:3]
    sign = torch.where(z >= 0, 1., -1.)
    sz = sign + z
    a = -torch.where(sz.abs() < 1e-6, torch.full_like(sz, 1e-6), sz).reciprocal()
    b = x * y * a
    s = F.normalize(torch.cat([x * x * a * sign + 1, b * sign, x * -sign], dim=-1), eps=1e-7, dim=-1)
    t = F.normalize(torch.cross(s, n, dim=-1), eps=1e-7, dim=-1)
    s = F.normalize(torch.cross(n, t, dim=-1), eps=1e-7, dim=-1)
    return torch.stack([s, t, n], dim=-1)


def partial_frame(n, wi):
    c = F.normalize(torch.cross(n, wi, dim=-1), eps=1e-7, dim=-1)
    return torch.stack([n, wi, c], dim=-1)


def to_local(frame, wo):
    """interaction.py:38-41: normalize(mean over xyz of frame * wo) -- i.e. frame^T wo / 3, renormalised."""
    if _fused_ok(frame, wo) and wo.shape[:-1] == frame.shape[:-2]:
        return ops.to_local(frame, wo)
    w = wo.unsqueeze(-1).expand_as(frame)
    return F.normalize((frame * w).mean(dim=-2), eps=1e-7, dim=-1)


def from_local(frame, v):
    """interaction.py:44-51."""
    out = frame[..., 0] * v[..., 0:1] + frame[..., 1] * v[..., 1:2] + frame[..., 2] * v[..., 2:3]
    return F.normalize(out, eps=1e-7, dim=-1)


@dataclass
class Interaction:
    p: torch.Tensor

    def spawn_rays(self, d):
        return torch.cat([self.p.expand_as(d), d], dim=-1)


@dataclass
class SurfaceInteraction(Interaction):
    uv: torch.Tensor = None
    wi: torch.Tensor = None
    t: torch.Tensor = None
    bsdf: object = None
    obj: object = None
    bidirectional_normals: bool = False
    frame = None
    n: torch.Tensor = None

    def set_normals(self, normals):
        self.n = normals
        self.frame = coordinate_system(normals)

    def to_local(self, wo):
        return to_local(self.frame, wo)

    def from_local(self, v):
        return from_local(self.frame, v)

    def shape(self):
        return self.p.shape

    def device(self):
        return self.p.device

    @classmethod
    def positions(cls, positions):
        return cls(positions)

    @classmethod
    def zeros(cls, shape, device):
        return cls(torch.zeros(shape, dtype=torch.float, device=device))

    @classmethod
    def like(cls, tensor):
        return cls(torch.zeros_like(tensor, dtype=torch.float))


@dataclass
class MixedInteraction(SurfaceInteraction):
    """Surface interaction that also carries the soft-silhouette logit (`throughput`)."""
    throughput: torch.Tensor = None
    with_logits: bool = True
    medium_mask = torch.tensor(False)

    # ---- spatially varying BSDF weights (bsdfs.py:527-536 stores them on the interaction for ALL rays) ----
    # The fused Direct path evaluates the sp_var MLP on the hits only; the misses are completed the first time either
    # attribute is read (colocate.py:104-105 reads normalized_weights of all rays for its extra loss).
    def _complete_weights(self):
        lazy = self.__dict__.pop("_lazy_weights", None)
        if lazy is None:
            return
        bsdf, logits_hit, idx, grad_mode = lazy
        nb = logits_hit.shape[-1]
        with torch.set_grad_enabled(grad_mode):          # the mode the shading ran in (e.g. no_grad renders)
            flat_p = self.p.reshape(-1, 3)
            miss = torch.ones(flat_p.shape[0], dtype=torch.bool, device=flat_p.device)
            miss[idx] = False
            midx = miss.nonzero().squeeze(-1)
            full = torch.zeros(flat_p.shape[0], nb, device=flat_p.device)
            if midx.numel() > 0:
                full = full.index_copy(0, midx, bsdf.sp_var_fn(bsdf.preprocess(flat_p[midx].detach())).reshape(-1, nb))
            full = full.index_copy(0, idx, logits_hit).reshape(self.p.shape[:-1] + (nb,))
            self.__dict__["_nonnormalized_weights"] = full
            self.__dict__["_normalized_weights"] = full.sigmoid()

    @property
    def normalized_weights(self):
        self._complete_weights()
        if "_normalized_weights" not in self.__dict__:
            raise AttributeError("normalized_weights (set by ComposeSpatialVarying.eval_and_pdf; no ray hit)")
        return self.__dict__["_normalized_weights"]

    @normalized_weights.setter
    def normalized_weights(self, v):
        self.__dict__["_normalized_weights"] = v

    @property
    def nonnormalized_weights(self):
        self._complete_weights()
        if "_nonnormalized_weights" not in self.__dict__:
            raise AttributeError("nonnormalized_weights (set by ComposeSpatialVarying.normalized_weights; no ray hit)")
        return self.__dict__["_nonnormalized_weights"]

    @nonnormalized_weights.setter
    def nonnormalized_weights(self, v):
        self.__dict__["_nonnormalized_weights"] = v

    def mark_mediums(self, medium_mask):
        self.medium_mask = medium_mask

    def surface_interactions(self):
        return ~self.medium_mask


@dataclass
class DirectionSample:
    p: torch.Tensor = None
    n: torch.Tensor = 0
    pdf: torch.Tensor = 1
    delta: torch.Tensor = True
    obj: object = None
    d: torch.Tensor = None
    dist: torch.Tensor = None
