"""Elementwise helpers with the semantics of pytorch3d/pathtracer/utils.py (only what the hot
path and its callers use).  All fp32 torch ops: this is differentiable glue, not the hot loop."""
import math
import random

import numpy as np
import torch
import torch.nn.functional as F


def create_fourier_basis2(batch_size, features=3, freq=40, device="cuda"):
    """utils.py:33-36: B = freq * randn(batch_size, features)^T, shape [features, batch_size]."""
    basis = freq * torch.randn(batch_size, features, device=device).T
    return basis, 2 * batch_size + features


def fourier2(x, B):
    """utils.py:37-40: [x, sin(x@B), cos(x@B)]."""
    phase = x @ B
    return torch.cat([x, phase.sin(), phase.cos()], dim=-1)


def nonzero_eps(v, eps: float = 1e-7):
    """utils.py:43-51: magnitudes below eps are replaced by +eps (sign deliberately dropped)."""
    return torch.where(v.abs() < eps, torch.full_like(v, eps), v)


def rotate_vector(v, axis, c, s):
    """Rodrigues rotation with given cosine / sine (utils.py:152-155)."""
    along = (v * axis).sum(dim=-1, keepdim=True)
    return v * c + axis * along * (1 - c) + torch.cross(axis, v, dim=-1) * s


def param_rusin2(wo, wi):
    """utils.py:233-258: (cos phi_d, cos theta_h, cos theta_d) from two local directions,
    including the reference's s = -sqrt(clamp(1 - H_z, 1e-6)) quirk."""
    if wo.is_cuda and wi.is_cuda and wo.shape == wi.shape and wo.dtype == torch.float32 and \
            not (torch.is_grad_enabled() and (wo.requires_grad or wi.requires_grad)):
        from .. import ops
        return ops.param_rusin2(wo, wi)
    wo = F.normalize(wo, dim=-1)
    wi = F.normalize(wi, dim=-1)
    y_axis = torch.tensor([0., 1., 0.], device=wo.device, dtype=wo.dtype).expand_as(wo)
    z_axis = torch.tensor([0., 0., 1.], device=wo.device, dtype=wo.dtype).expand_as(wo)
    half = F.normalize(wo + wi, dim=-1)
    hx, hy, hz = half[..., 0], half[..., 1], half[..., 2]
    r = nonzero_eps(hy).hypot(nonzero_eps(hx)).clamp(min=1e-6)
    tmp = F.normalize(rotate_vector(wi, z_axis, (hx / r).unsqueeze(-1), -(hy / r).unsqueeze(-1)), dim=-1)
    s = -(1 - hz).clamp(min=1e-6).sqrt().unsqueeze(-1)
    diff = F.normalize(rotate_vector(tmp, y_axis, hz.unsqueeze(-1), s), dim=-1)
    cos_phi_d = torch.atan2(nonzero_eps(diff[..., 1]), nonzero_eps(diff[..., 0])).cos()
    return torch.stack([cos_phi_d, hz, diff[..., 2]], dim=-1)


def dir_to_elev_azim(direc):
    """utils.py:490-494."""
    d = F.normalize(direc, dim=-1).clamp(min=-1 + 1e-7, max=1 - 1e-7)
    x, z = d[..., 0:1], d[..., 2:3]
    elev = z.asin()
    azim = torch.atan2(x, (1 - x.square() - z.square()).clamp(min=1e-10).sqrt())
    return torch.cat([elev, azim], dim=-1)


def elev_azim_to_dir(elev_azim):
    """utils.py:479-487 (limit = pi - 1e-7)."""
    limit = math.pi - 1e-7
    ea = elev_azim.clamp(min=-limit, max=limit)
    elev, azim = ea[..., 0:1], ea[..., 1:2]
    return torch.cat([azim.sin() * elev.cos(), azim.cos() * elev.cos(), elev.sin()], dim=-1)


def smooth_min(v, k: float = 32, dim: int = 0):
    """utils.py:385-387."""
    return -torch.exp(-k * v).sum(dim).clamp(min=1e-4).log() / k


def eikonal_loss(grad):
    """utils.py:294-295."""
    return (torch.norm(grad, dim=-1) - 1).square().mean()


def mse2psnr(x):
    return -10 * torch.log10(x)


def count_parameters(params):
    return sum(p.numel() for p in params)


def crop(img, u, v, size):
    return img[u:u + size, v:v + size, ...]


def rand_uv(w: int, h: int, size: int):
    """utils.py:374-375 (python RNG, like the reference)."""
    return random.randint(0, w - size), random.randint(0, h - size)


def rand_uv_mask(mask, size: int):
    """utils.py:378-383: top-left corner of a crop, drawn among the mask's non-zero pixels that keep the crop
    (and half a crop of margin) inside the image; python RNG like the reference."""
    half = int(math.ceil(size / 2))
    valid = mask[half:-half - size, half:-half - size, ...]
    p, q = valid.nonzero(as_tuple=True)[:2]
    idx = random.randint(0, len(p) - 1)
    return p[idx], q[idx]


def load_image(src, resize=None):
    """utils.py:365-369: image file -> float tensor in [0, 1]."""
    from PIL import Image
    img = Image.open(src)
    if resize is not None:
        img = img.resize(resize)
    return torch.from_numpy(np.array(img, dtype=float) / 255).float()


class LossSampler:
    """utils.py:134-147: samples view indices proportionally to the square of their last loss."""

    def __init__(self, N, default=1e5, likelihood_inc=1.00001):
        self.losses = np.array([default] * N)
        self.l_inc = likelihood_inc

    def update(self, idx, loss):
        self.losses *= self.l_inc
        self.losses[idx] = loss + 1

    def sample(self, n=1, replace=False):
        sq = self.losses * self.losses
        return np.random.choice(len(self.losses), replace=replace, size=n, p=sq / sq.sum())

    def update_idxs(self, idxs, loss):
        for i in idxs:
            self.update(i, loss)


def masked_loss(got, exp, throughput, exp_mask, eps: float = 1e-10, trim: int = 0, mask_weight: float = 1,
                with_logits: bool = True, tone_mapping: bool = False, ssim_fn="default"):
    """utils.py:307-359: 10 * (l2 + rmse + l1 - log ssim) on the hit pixels + mask_weight * BCE on the misses.
    The SSIM term of the reference comes from the unpinned third-party `pytorch_msssim`; the default here is the
    in-repo restatement (pathtracer/ssim.py, parity unpinned).  `ssim_fn=None` drops the term
    (what the benchmarks do), any callable `(a, b) -> scalar` replaces it."""
    if ssim_fn == "default":
        from .ssim import ssim as _ssim

        def ssim_fn(a, b):
            return _ssim(a, b, data_range=1, size_average=True)
    active = ((throughput > 0) & (exp_mask == 1)).squeeze(-1)
    misses = ~active
    color_loss = 0
    if active.any():
        ga = got * active[..., None]
        ea = exp * active[..., None]
        if tone_mapping:
            ga, ea = ga / (1 + ga), ea / (1 + ea)
        l1 = F.l1_loss(ga, ea)
        l2 = F.mse_loss(ga, ea)
        color_loss = l2 + l2.clamp(min=1e-10).sqrt() + l1
        if ssim_fn is not None:
            color_loss = color_loss - ssim_fn(ga.permute(0, 3, 1, 2), ea.permute(0, 3, 1, 2)).log()
    mask_loss = 0
    if misses.any():
        fn = F.binary_cross_entropy_with_logits if with_logits else F.binary_cross_entropy
        mask_loss = fn(throughput[misses].reshape(-1, 1), exp_mask[misses].reshape(-1, 1))
    return mask_weight * mask_loss + 10 * color_loss


# ---- helpers of the vis scripts (utils.py:409-445) --------------------------------------------------------------
def sphere_examples(bsdf, device="cuda", size=256, chunk_size=128, scale=100):
    """One Direct-lit render of the unit sphere per basis of a spatially varying BSDF (utils.py:409-431; dtu_vis.py:108,
    nerv_vis.py, visualize.py): analytic sphere, look-at camera at distance 2, one point light at (0, 1, 4)."""
    from . import integrators
    from .main import pathtrace
    from .shapes import Sphere
    from ..renderer import OpenGLPerspectiveCameras, PointLights, look_at_view_transform
    ball = Sphere([0, 0, 0], 1, device=device)
    R, T = look_at_view_transform(dist=2.0, elev=0, azim=0)
    cameras = OpenGLPerspectiveCameras(device=device, R=R, T=T)
    lights = PointLights(device=device, location=[[0.0, 1.0, 4.0]], scale=scale)
    return [pathtrace(ball, cameras=cameras, lights=lights, chunk_size=chunk_size, size=size, bsdf=basis,
                      integrator=integrators.Direct(), device=device, silent=True)[0] for basis in bsdf.bsdfs]


def heightmap(warp, size=256, device="cuda"):
    """pdf of a warp over the unit square (utils.py:434-439)."""
    u, v = torch.meshgrid(torch.linspace(0, 1, size, device=device), torch.linspace(0, 1, size, device=device), indexing="ij")
    return warp.pdf(torch.stack([u, v], dim=-1))


def depth_image(img):
    """[depth | mask] -> grey RGBA with the depth scaled by its maximum (utils.py:441-445; colocate.py, nerfle.py)."""
    depth, mask = img.split(1, dim=-1)
    depth = depth / depth.max()
    return torch.cat([depth, depth, depth, mask], dim=-1)
