"""Analytic shapes the vis scripts render BSDF bases on (pytorch3d/pathtracer/shapes/shapes.py:9-97): a single sphere with a
closed-form ray intersection.  Elementwise torch (a handful of ops per ray; the shading that follows is the hot part and runs
on the library's MLP kernels through the BSDF classes)."""
import math

import torch
import torch.nn.functional as F

from ..interaction import SurfaceInteraction

EPS = 1e-8


class Sphere:
    """One sphere (centre, radius).  `intersect` follows the surface protocol of the integrators: (SurfaceInteraction,
    hit mask); the hit point is pushed 1e-5 along the normal, misses carry t = inf (shapes.py:44-68)."""

    def __init__(self, center, radius, device="cuda"):
        self.device = torch.device(device)
        self.center = torch.tensor(center, device=device, dtype=torch.float)
        self.radius = float(radius)
        self.sqr_radius = self.radius * self.radius

    def __len__(self):
        return 1

    def _roots(self, rays):
        """Both roots of |o + t d - c|^2 = r^2 as [..., 2] (the non-real ones keep the raw discriminant, exactly as the
        reference's quad_solve leaves them, shapes.py:11-18) and the mask of rays with a real root at t >= EPS."""
        r_o, r_d = torch.split(rays, 3, dim=-1)
        rel = r_o - self.center
        a = (r_d * r_d).sum(dim=-1)
        b = 2 * (r_d * rel).sum(dim=-1)
        c = (rel * rel).sum(dim=-1) - self.sqr_radius
        disc = b * b - 4 * a * c
        real = disc > 0
        root = torch.where(real, disc.clamp(min=0).sqrt(), disc)
        ts = (-b.unsqueeze(-1) + torch.stack([root, -root], dim=-1)) / (2 * a.unsqueeze(-1))
        return r_o, r_d, ts, real & (ts >= EPS).any(dim=-1)

    def intersect(self, rays, active=True, primary=True, **_unused):
        r_o, r_d, ts, hit = self._roots(rays)
        ts = torch.where(ts < EPS, torch.full_like(ts, math.inf), ts)     # behind the origin
        t = ts.min(dim=-1).values
        p = r_o + t.unsqueeze(-1) * r_d
        n = F.normalize(p - self.center, dim=-1)
        si = SurfaceInteraction(p=p + n * 1e-5, t=t, obj=self)
        si.set_normals(n)
        si.wi = si.to_local(-r_d)
        return si, hit

    def intersect_test(self, rays, active=True, **_unused):
        return self._roots(rays)[3]

    def intersect_limits(self, rays, max_t=math.inf, active=True):
        _o, _d, ts, hit = self._roots(rays)
        ts = torch.where(ts < EPS, torch.full_like(ts, math.inf), ts)
        return ts.min(dim=-1).values, ts.max(dim=-1).values, hit
