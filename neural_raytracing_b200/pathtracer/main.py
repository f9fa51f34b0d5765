"""Render entry points with the reference's signatures (pytorch3d/pathtracer/main.py:13-179)."""
import torch

from .samplers import Sampler
from .utils import rand_uv


def nothing(_):
    return None


def _pixel_grid(x0, y0, nx, ny, device):
    gx, gy = torch.meshgrid(torch.arange(x0, x0 + nx, device=device, dtype=torch.float),
                            torch.arange(y0, y0 + ny, device=device, dtype=torch.float), indexing="ij")
    return torch.stack([gy, gx], dim=-1)      # the reference stacks (y, x)


def _render_tile(shapes, lights, cameras, integrator, bsdf, positions, sampler, bundle_size, size, batch_dims,
                 with_noise, w_isect, background):
    rays = cameras.sample_positions(positions, sampler, bundle_size, size=size, N=batch_dims, with_noise=with_noise)
    values, mask, it = integrator.sample(shapes, rays, bsdf=bsdf, lights=lights, sampler=sampler, w_isect=w_isect)
    valid = mask.any(dim=-1)
    v = torch.mean(values, dim=-2)
    v[~valid] = background
    return v, it


def pathtrace(shapes, lights, cameras, integrator, bsdf=None, size=512, width=None, height=None, chunk_size=32,
              bundle_size=4, background=1, addition=nothing, sampler=None, silent=False, trim=0, device="cuda",
              squeeze_first=True, w_isect=False, with_noise=1e-3):
    """Renders [len(cameras), width, height, integrator.dims()] by tiles of chunk_size (main.py:13-93)."""
    sampler = sampler if sampler is not None else Sampler(device=device)
    batch_dims = len(cameras)
    width = size if width is None else width
    height = size if height is None else height
    out = torch.full([batch_dims, width, height, integrator.dims()], background, device=device, dtype=torch.float)
    assert (size % chunk_size) == 0, \
        f"Can only specify chunk sizes which evenly divide size, {size} % {chunk_size}"
    it = None
    for x0 in range(0, width, chunk_size):
        for y0 in range(0, height, chunk_size):
            pos = _pixel_grid(x0 - trim, y0 - trim, chunk_size + 2 * trim, chunk_size + 2 * trim, device)
            v, it = _render_tile(shapes, lights, cameras, integrator, bsdf, pos, sampler, bundle_size, size, batch_dims,
                                 with_noise, w_isect, background)
            if trim != 0:
                v = v[:, trim:-trim, trim:-trim]
            out[:, x0:x0 + chunk_size, y0:y0 + chunk_size, :] = v
    if squeeze_first and batch_dims == 1:
        out = out.squeeze(0)
    return out, addition(it)


def pathtrace_sample(shapes, lights, cameras, integrator, bsdf=None, size=512, chunk_size=32, bundle_size=4,
                     crop_size=128, uv=None, background=1, sampler=None, addition=nothing, silent=False, mode="crop",
                     device="cuda", squeeze_first=True, w_isect=False, with_noise=1e-2):
    """Renders one crop_size x crop_size window of the size x size image (main.py:97-179)."""
    sampler = sampler if sampler is not None else Sampler(device=device)
    if uv is None:
        uv = rand_uv(size, size, crop_size)
    batch_dims = len(cameras)
    shape = [batch_dims, crop_size, crop_size, integrator.dims()] if mode == "crop" else \
        [batch_dims, size, size, integrator.dims()]
    out = torch.full(shape, background, device=device, dtype=torch.float)
    assert (size % chunk_size) == 0, \
        f"Can only specify chunk sizes which evenly divide size, {size} % {chunk_size}"
    chunk_size = min(chunk_size, crop_size)
    u = max(min(int(uv[0]), size - crop_size), 0)
    v = max(min(int(uv[1]), size - crop_size), 0)
    it = None
    for x0 in range(u, u + crop_size, chunk_size):
        for y0 in range(v, v + crop_size, chunk_size):
            pos = _pixel_grid(x0, y0, chunk_size, chunk_size, device)
            vals, it = _render_tile(shapes, lights, cameras, integrator, bsdf, pos, sampler, bundle_size, size,
                                    batch_dims, with_noise, w_isect, background)
            if mode == "crop":
                out[:, x0 - u:x0 - u + chunk_size, y0 - v:y0 - v + chunk_size] = vals
            else:
                out[:, x0:x0 + chunk_size, y0:y0 + chunk_size] = vals
    if squeeze_first and batch_dims == 1:
        out = out.squeeze(0)
    setattr(it, "crop_uv", uv)
    return out, addition(it)
