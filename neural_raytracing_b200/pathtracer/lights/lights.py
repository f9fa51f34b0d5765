"""Emitters (pytorch3d/pathtracer/lights/lights.py): PointLights and the learned LightField."""
from itertools import chain

import torch
import torch.nn as nn
import torch.nn.functional as F

from ..interaction import DirectionSample
from ..neural_blocks import SkipConnMLP


class Light(nn.Module):
    def sample_towards(self, points, sampler):
        raise NotImplementedError()

    def sample_direction(self, it, sampler, active=True):
        raise NotImplementedError()

    def intersect(self, _rays):
        return None, False


class PointLights(Light):
    """Point lights with learned constant / linear / quadratic falloff (lights.py:40-110).
    scale / const / linear / square stay 0-dim CPU leaf tensors as in the reference
    (they are in the optimiser through intensity_parameters())."""

    def __init__(self, intensity=[1., 1., 1.], location=[0, 1, 0], const=1e-8, linear=1e-8, square=1, scale=1e2,
                 device="cuda"):
        super().__init__()
        self.device = device
        self.scale = torch.tensor(scale, dtype=torch.float, requires_grad=True)
        if type(intensity) is torch.Tensor:
            self.intensity = intensity
        else:
            assert type(intensity) is list
            self.intensity = torch.tensor([intensity], device=device, requires_grad=True, dtype=torch.float)
        if type(location) is torch.Tensor:
            self.location = location
        else:
            assert type(location) is list
            self.location = torch.tensor(location, device=device, requires_grad=True, dtype=torch.float)
            if len(self.location.shape) == 1:
                self.location = self.location.unsqueeze(0).detach()
        self.const = torch.tensor(const, dtype=torch.float, requires_grad=True)
        self.linear = torch.tensor(linear, dtype=torch.float, requires_grad=True)
        self.square = torch.tensor(square, dtype=torch.float, requires_grad=True)

    def parameters(self):
        return chain(self.location_parameters(), self.spectrum_parameters())

    def location_parameters(self):
        return [self.location]

    def spectrum_parameters(self):
        return [self.scale, self.intensity, self.const, self.linear, self.square]

    def intensity_parameters(self):
        return [self.scale, self.const, self.linear, self.square]

    def coefficients(self):
        return [self.const, self.linear, self.square]

    def sample_towards(self, points):
        return F.normalize(points - self.location, dim=-1)

    def _falloff(self, dist):
        denom = self.const.clamp(min=1e-6) + self.linear.clamp(min=1e-6) * dist + \
            self.square.clamp(min=1e-6) * dist.square()
        return denom.clamp(min=1e-6)

    def envmap(self, p):
        d = p[None, ...] - self.location[:, None, None, :]
        dist = torch.linalg.norm(d, dim=-1, keepdim=True)
        return self.scale * F.normalize(self.intensity, dim=-1) / self._falloff(dist)

    def sample_direction(self, it, sampler, active=True):
        ds = DirectionSample()
        ds.p = self.location[:, None, None, None, :]
        ds.n, ds.uv, ds.obj, ds.delta = 0, 0, self, True
        d = ds.p - it.p
        ds.dist = torch.linalg.norm(d, dim=-1, keepdim=True)
        ds.d = F.normalize(d, eps=1e-6, dim=-1)
        color = self.intensity[:, None, None, None, :]
        spectrum = self.scale * F.normalize(color, dim=-1) / self._falloff(ds.dist)
        spectrum[~active] = 0
        return ds, spectrum


def identity(x):
    return x


class LightField(nn.Module):
    """5-D light field p -> direction * magnitude with a constant colour (lights.py:155-195)."""

    def __init__(self, device="cuda"):
        super().__init__()
        self.light_field_approx = SkipConnMLP(in_size=3, out=3, num_layers=10, hidden_size=256, device=device).to(device)
        self.color = nn.Parameter(torch.tensor([0., 0., 0.], dtype=torch.float, device=device), requires_grad=True)
        self.device = device
        self.preproc = identity
        self.postproc = identity

    def sample_towards(self, points, sampler):
        raise NotImplementedError()

    def sample_direction(self, it, sampler, active=True):
        pre = getattr(self, "preproc", identity)
        v = self.light_field_approx(pre(it.p[active]))      # compacted to the hits, like the reference
        ds = DirectionSample()
        ds.p, ds.dist, ds.n, ds.uv, ds.obj, ds.delta = None, None, 0, 0, self, True
        ds.pdf = torch.ones(it.p.shape[:-1], device=it.p.device, dtype=torch.float)
        ds.d = torch.zeros_like(it.p)
        ds.d[active] = F.normalize(v, eps=1e-6, dim=-1).clamp(min=1e-6, max=1)   # components clamped positive (quirk)
        magn = torch.linalg.norm(v, ord=2, dim=-1, keepdim=True)
        spectrum = torch.zeros_like(it.p, dtype=torch.float)
        spectrum[active] = magn * self.color.sigmoid()
        return ds, spectrum
