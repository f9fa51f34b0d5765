"""Volumetric NeRF models (pytorch3d/pathtracer/shapes/nerf.py): NeRFLE and PlainNeRF."""
import random

import torch
import torch.nn as nn
import torch.nn.functional as F

from ... import config, ops
from ..neural_blocks import SkipConnMLP
from ..utils import dir_to_elev_azim, elev_azim_to_dir


def composite_reference_ops(sigma_raw, rgb, ts):
    """nerf.py:205-213 on sample-major tensors with torch ops (differentiable)."""
    sigma_a = F.relu(sigma_raw)
    shape = (-1,) + (1,) * (sigma_a.dim() - 1)
    alpha = 1 - torch.exp(-sigma_a * ts.reshape(shape))
    cp = torch.cumprod((1 - alpha).clamp(min=1e-10), dim=0)
    cp = torch.roll(cp, 1, 0)
    cp[-1, ...] = 1
    return ((alpha * cp)[..., None] * rgb).sum(dim=0)


class _Composite(torch.autograd.Function):
    """Compositing through the HBM-bound CUDA kernels (forward + backward)."""

    @staticmethod
    def forward(ctx, sigma_raw, rgb, ts):
        s, c, t = sigma_raw.contiguous().float(), rgb.contiguous().float(), ts.contiguous().float()
        ctx.save_for_backward(s, c, t)
        return ops.composite_forward(s, c, t)

    @staticmethod
    def backward(ctx, g):
        s, c, t = ctx.saved_tensors
        gs, gc = ops.composite_backward(s, c, t, g.contiguous().float())
        return gs, gc, None


def composite(sigma_raw, rgb, ts):
    if sigma_raw.is_cuda:
        return _Composite.apply(sigma_raw, rgb, ts)
    return composite_reference_ops(sigma_raw, rgb, ts)


class _NerfleFused(torch.autograd.Function):
    """first MLP -> [latent | r_d | light] -> second MLP -> sigmoid on the tensor-core training kernels, nothing but
    sigma [S,...] and rgb [S,...,3] materialised in fp32 (ops.nerfle_train_forward / _backward).  Gradients reach the
    weights of both MLPs; rays, ts and the light code are treated as constants (as in nerfle.py, where only
    `model.parameters()` are optimised)."""

    @staticmethod
    def forward(ctx, module, rays, ts, code, view, prec, *params):
        p1, p2 = module.first.packed(), module.second.packed()
        sigma, rgb, state = ops.nerfle_train_forward(p1, p2, rays.detach().float(), ts.detach().float(), code.detach().float(), view, prec)
        ctx.p1, ctx.p2, ctx.state, ctx.prec = p1, p2, state, prec
        ctx.fg1, ctx.fg2 = getattr(module.first, "_flat_grad", None), getattr(module.second, "_flat_grad", None)
        if (ctx.fg1 is None) != (ctx.fg2 is None):
            ctx.fg1 = ctx.fg2 = None
        ctx.n1 = len(module.first._flat_params())
        ctx.save_for_backward(rgb)
        lead = tuple(rays.shape[:-1])
        return sigma.reshape((ts.numel(),) + lead), rgb.reshape((ts.numel(),) + lead + (3,))

    @staticmethod
    def backward(ctx, g_sigma, g_rgb):
        rgb, = ctx.saved_tensors
        g1, g2 = ops.nerfle_train_backward(ctx.p1, ctx.p2, rgb, g_sigma.contiguous().float(), g_rgb.contiguous().float(),
                                           ctx.state, ctx.prec, ctx.fg1, ctx.fg2)
        if ctx.fg1 is not None:      # accumulated straight into training.FlatParameters' gradient buffer
            return (None,) * (6 + 2 * (len(ctx.p1.dims) + len(ctx.p2.dims)))
        flat = []
        for pk, g in ((ctx.p1, g1), (ctx.p2, g2)):
            gW, gb = pk.unpack(g)
            for w, b in zip(gW, gb):
                flat += [w, b]
        return (None, None, None, None, None, None) + tuple(flat)


class NeRFLE(nn.Module):
    """NeRF with a point light / environment-light code (nerf.py:153-214)."""

    def __init__(self, envmap=False, bins=4, device="cuda"):
        super().__init__()
        self.latent_size = 64
        self.first = SkipConnMLP(num_layers=5, hidden_size=128, in_size=3, out=1 + self.latent_size,
                                 device=device).to(device)
        self.bins = bins
        self.second = SkipConnMLP(in_size=self.latent_size + (6 if not envmap else 3 + bins * bins * 3), out=3,
                                  device=device).to(device)
        self.envmap = envmap
        # optional device tensor in [0,1): when set, replaces random.random() of the far-plane jitter (nerf.py:178) so
        # that a CUDA-graph-captured step (neural_raytracing_b200.training.GraphedStep) still jitters per replay
        self.far_jitter = None

    def _sample_ts(self, device):
        if getattr(self, "far_jitter", None) is not None:
            return torch.linspace(0, 1, 64, device=device) * (2 + self.far_jitter.reshape(()).to(device) * 0.1)
        return torch.linspace(0, 2 + random.random() * 0.1, 64, device=device)

    def _light_code(self, lights, device):
        """[n_views, 3] light location, or the [n_views, 3*bins^2] environment code (nerf.py:184-197)."""
        if getattr(self, "envmap", False):
            grid = torch.stack(torch.meshgrid(torch.linspace(0, 180, self.bins, device=device),
                                              torch.linspace(0, 45, self.bins, device=device), indexing="ij"),
                               dim=-1).reshape(-1, 2)
            code = lights.envmap(elev_azim_to_dir(grid))
            return code.reshape(code.shape[0], -1)
        return lights.location

    @staticmethod
    def _code_rows(code, n_views):
        """One light-code row per view: a single light broadcasts over the views as it does in nerf.py:199-201 (`expand`)."""
        if code.shape[0] == 1 and n_views > 1:
            return code.expand(n_views, code.shape[-1])
        if code.shape[0] != n_views:
            raise ValueError("light code has %d rows for %d views" % (code.shape[0], n_views))
        return code

    def _view_index(self, rays):
        """[R] int32: which view (light code row) each ray belongs to; cached per batch shape."""
        key = (rays.shape[0], rays[0].numel() // 6, rays.device)
        cache = self.__dict__.setdefault("_view_cache", {})
        if key not in cache:
            cache.clear()
            cache[key] = torch.arange(key[0], device=rays.device, dtype=torch.int32).repeat_interleave(key[1])
        return cache[key]

    def _needs_grad(self, rays):
        return torch.is_grad_enabled() and (rays.requires_grad or any(p.requires_grad for p in self.parameters()))

    def render_camera(self, cam, lights):
        """The frame of `cam` (ops.CameraDesc) in one library call, rays generated on the device inside it (f4):
        what pathtrace's tile loop + sample_positions + forward compute for an inference render (main.py:57-88,
        nerf.py:175-214), with ONE far-plane draw (nerf.py:178) per frame instead of one per tile.
        -> rgb [n_views, nx, ny, bundle, 3]."""
        device = cam.device
        ts = self._sample_ts(device)
        code = self._code_rows(self._light_code(lights, device), cam.n_views)
        prec = config.precision if self.first.precision() != "f32" and self.second.precision() != "f32" else "f32"
        return ops.nerfle_render_camera(self.first.packed(), self.second.packed(), cam, ts, code.detach().float(),
                                        prec=prec)

    def forward(self, rays, lights):
        r_o, r_d = rays.split([3, 3], dim=-1)
        device = r_o.device
        ts = self._sample_ts(device)
        code = self._code_rows(self._light_code(lights, device), rays.shape[0])
        if rays.is_cuda and not self._needs_grad(rays):
            # fused render: rays in, rgb out
            view = self._view_index(rays)
            # the fused tensor-core render instantiates the point-light and the bins = 4 environment net only: any
            # other shape keeps the exact fp32 kernel (SkipConnMLP.precision() knows which shapes are instantiated)
            prec = config.precision if self.first.precision() != "f32" and self.second.precision() != "f32" else "f32"
            return ops.nerfle_render(self.first.packed(), self.second.packed(), rays.detach().float(), ts,
                                     code.detach().float(), view, prec=prec)
        # differentiable path: fused MLP kernels where available + CUDA compositing
        tprec = self.first.train_precision()
        if rays.is_cuda and tprec != "f32" and self.second.train_precision() == tprec and not rays.requires_grad \
                and not code.requires_grad:
            view = self._view_index(rays)
            alpha, rgb = _NerfleFused.apply(self, rays, ts, code, view, tprec, *self.first._flat_params(),
                                            *self.second._flat_params())
            return composite(alpha, rgb, ts)
        pts = r_o.unsqueeze(0) + torch.tensordot(ts, r_d, dims=0)
        first_out = self.first(pts)
        latent, alpha = first_out[..., 1:], first_out[..., 0]
        lead = latent.shape[:-1]
        light = code.reshape((1, code.shape[0]) + (1,) * (len(lead) - 2) + (-1,)).expand(lead + (-1,))
        rgb = self.second(torch.cat([latent, r_d[None, ...].expand(lead + (3,)), light], dim=-1), out_act=ops.OUT_SIGMOID)
        return composite(alpha, rgb, ts)


class PlainNeRF(nn.Module):
    """Per-image-latent NeRF (nerf.py:9-74)."""

    def __init__(self, latent_size: int = 32, intermediate_size: int = 32, steps=32, device="cuda"):
        super().__init__()
        self.latent = None
        self.latent_size = latent_size
        self.steps = steps
        self.first = SkipConnMLP(in_size=3, out=1 + intermediate_size, latent_size=latent_size, num_layers=5,
                                 hidden_size=32, device=device).to(device)
        self.second = SkipConnMLP(in_size=2, out=3, latent_size=latent_size + intermediate_size, num_layers=5,
                                  hidden_size=32, device=device).to(device)

    def assign_latent(self, latent):
        assert latent.shape[-1] == self.latent_size
        assert len(latent.shape) == 2, "expected latent in [B, L]"
        self.latent = latent

    def forward(self, rays, lights):
        assert self.latent is not None
        r_o, r_d = rays.split([3, 3], dim=-1)
        ts = torch.linspace(0.4, 2 + random.random() * 0.1, self.steps, device=r_o.device)
        pts = r_o.unsqueeze(0) + torch.tensordot(ts, r_d, dims=0)
        latent = self.latent[None, :, None, None, None, :].expand(pts.shape[:-1] + (-1,))
        first_out = self.first(pts, latent)
        alpha, inter = first_out[..., 0], first_out[..., 1:]
        view = dir_to_elev_azim(r_d)[None, ...].expand(latent.shape[:-1] + (2,))
        rgb = self.second(view, torch.cat([inter, latent], dim=-1)).tanh()
        alpha = alpha + torch.randn_like(alpha) * 1e-3      # nerf.py:66
        return (composite(alpha, rgb, ts) + 1) / 2
