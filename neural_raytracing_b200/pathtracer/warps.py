"""Sample warps used by `bsdf.sample` (pytorch3d/pathtracer/warps.py:10-52): unit square -> concentric disk ->
cosine-weighted hemisphere.  Elementwise torch (a handful of ops on [..., 2] samples per bounce)."""
import math

import torch


def square_to_uniform_disk_concentric(sample):
    """Shirley-Chiu concentric map as the reference writes it (warps.py:10-30): the radius is the coordinate of larger
    magnitude (sign kept, |r| >= 1e-12), the angle pi/4 * (other / r), mirrored in quadrants 1 and 3; returns
    (r sin(phi), r cos(phi))."""
    v = 2 * sample - 1
    x, y = v[..., 0:1], v[..., 1:2]
    swap = x.abs() < y.abs()
    r = torch.where(swap, y, x)
    rp = torch.where(swap, x, y)
    r = r.sign() * r.abs().clamp(min=1e-12)
    phi = (0.25 * math.pi) * rp / r
    phi = torch.where(swap, 0.5 * math.pi - phi, phi)
    phi = torch.where((v == 0).all(dim=-1, keepdim=True), torch.zeros_like(phi), phi)
    return torch.cat([r * phi.sin(), r * phi.cos()], dim=-1)


def square_to_cos_hemisphere(sample):
    """warps.py:44-49: lift the disk sample onto the hemisphere (z >= sqrt(1e-7))."""
    p = square_to_uniform_disk_concentric(sample)
    z = (1 - (p * p).sum(dim=-1, keepdim=True)).clamp(min=1e-7).sqrt()
    return torch.cat([p, z], dim=-1)


def square_to_cos_hemisphere_pdf(d):
    """warps.py:51-52."""
    return d[..., 2] / math.pi
