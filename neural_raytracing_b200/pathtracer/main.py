"""Render entry points with the reference's signatures (pytorch3d/pathtracer/main.py:13-179)."""
import torch

from .. import ops
from .samplers import Sampler
from .utils import rand_uv


def nothing(_):
    return None


def _pixel_grid(x0, y0, nx, ny, device):
    gx, gy = torch.meshgrid(torch.arange(x0, x0 + nx, device=device, dtype=torch.float),
                            torch.arange(y0, y0 + ny, device=device, dtype=torch.float), indexing="ij")
    return torch.stack([gy, gx], dim=-1)      # the reference stacks (y, x)


def _tile_rays(cameras, window, sampler, bundle_size, size, batch_dims, with_noise, device):
    """Rays [N, nx, ny, bundle, 6] of the pixel window (x0, y0, nx, ny).  Without pixel jitter and with a CUDA camera
    the window goes to the ray-generation kernel as four integers (no position tensor); with jitter the positions are
    built and perturbed with torch's generator exactly as the reference does (main.py:67-80)."""
    x0, y0, nx, ny = window
    if not with_noise and str(device).startswith("cuda") and hasattr(cameras, "device_desc"):
        desc = cameras.device_desc(size, x0=x0, y0=y0, nx=nx, ny=ny, bundle_size=bundle_size)
        if desc is not None:
            return ops.camera_rays(desc)
    positions = _pixel_grid(x0, y0, nx, ny, device)
    return cameras.sample_positions(positions, sampler, bundle_size, size=size, N=batch_dims, with_noise=with_noise)


def _render_tile(shapes, lights, cameras, integrator, bsdf, window, sampler, bundle_size, size, batch_dims,
                 with_noise, w_isect, background, device):
    rays = _tile_rays(cameras, window, sampler, bundle_size, size, batch_dims, with_noise, device)
    values, mask, it = integrator.sample(shapes, rays, bsdf=bsdf, lights=lights, sampler=sampler, w_isect=w_isect)
    valid = mask.any(dim=-1)
    v = torch.mean(values, dim=-2)
    v[~valid] = background
    return v, it


def _camera_frame(shapes, lights, cameras, integrator, size, x0, y0, nx, ny, bundle_size, with_noise, background):
    """f4: a volumetric shape under NeRFReproduce on a CUDA camera renders its whole window in ONE library call, the
    rays generated on the device inside it (shapes.render_camera -> nrt_nerfle_render_camera) -- no tile loop, no ray
    tensor, no per-tile launches.  Returns None when that does not apply (surface integrators, training, CPU cameras).
    Rays are independent, so the image equals the tiled one for any chunk_size / trim; pixel jitter comes from the
    library's counter hash instead of torch's generator (same distribution: (U - 0.5) * with_noise pixels)."""
    from .integrators import NeRFReproduce
    if type(integrator) is not NeRFReproduce or not hasattr(shapes, "render_camera") \
            or not hasattr(cameras, "device_desc"):
        return None
    if torch.is_grad_enabled() and any(p.requires_grad for p in shapes.parameters()):
        return None
    seed = int(torch.randint(1, 2 ** 31 - 1, (1,)).item()) if with_noise else 0      # CPU generator: no device sync
    cam = cameras.device_desc(size, x0=x0, y0=y0, nx=nx, ny=ny, bundle_size=bundle_size,
                              jitter=float(with_noise or 0.0), jitter_seed=seed)
    if cam is None:
        return None
    values = shapes.render_camera(cam, lights)
    return torch.mean(values, dim=-2)       # every pixel is valid under NeRFReproduce (integrators.py:260-267)


def _row_block(width, height, rays_per_pixel, chunk_size, addition, device):
    """Rows per integrator call for a gradient-free GPU frame, or 0 to keep the caller's chunk_size tiles (training, CPU, a
    caller that reads the last tile's interaction through `addition`, or tiles that are already larger)."""
    from .. import config
    budget = config.max_tile_rays
    if budget <= 0 or torch.is_grad_enabled() or addition is not nothing or not str(device).startswith("cuda"):
        return 0
    rows = min(width, budget // max(1, height * rays_per_pixel))
    if rows < 1 or rows * height <= chunk_size * chunk_size:
        return 0                                         # one row is over the budget, or the caller's tile is at least as large
    return rows


def pathtrace(shapes, lights, cameras, integrator, bsdf=None, size=512, width=None, height=None, chunk_size=32,
              bundle_size=4, background=1, addition=nothing, sampler=None, silent=False, trim=0, device="cuda",
              squeeze_first=True, w_isect=False, with_noise=1e-3):
    """Renders [len(cameras), width, height, integrator.dims()] by tiles of chunk_size (main.py:13-93)."""
    sampler = sampler if sampler is not None else Sampler(device=device)
    batch_dims = len(cameras)
    width = size if width is None else width
    height = size if height is None else height
    assert (size % chunk_size) == 0, \
        f"Can only specify chunk sizes which evenly divide size, {size} % {chunk_size}"
    it = None
    if addition is nothing and str(device).startswith("cuda"):
        frame = _camera_frame(shapes, lights, cameras, integrator, size, 0, 0, width, height, bundle_size, with_noise,
                              background)
        if frame is not None:
            return (frame.squeeze(0) if squeeze_first and batch_dims == 1 else frame), None
    out = torch.full([batch_dims, width, height, integrator.dims()], background, device=device, dtype=torch.float)
    rows = _row_block(width, height, batch_dims * bundle_size, chunk_size, addition, device)
    if rows:
        # gradient-free frame on the GPU: row blocks of up to config.max_tile_rays rays instead of chunk_size tiles (the
        # per-ray results do not depend on the tiling; `trim` only widens tiles whose border is then cut off again)
        for x0 in range(0, width, rows):
            nx = min(rows, width - x0)
            v, it = _render_tile(shapes, lights, cameras, integrator, bsdf, (x0, 0, nx, height), sampler, bundle_size, size,
                                 batch_dims, with_noise, w_isect, background, device)
            out[:, x0:x0 + nx, :, :] = v
        if squeeze_first and batch_dims == 1:
            out = out.squeeze(0)
        return out, addition(it)
    for x0 in range(0, width, chunk_size):
        for y0 in range(0, height, chunk_size):
            win = (x0 - trim, y0 - trim, chunk_size + 2 * trim, chunk_size + 2 * trim)
            v, it = _render_tile(shapes, lights, cameras, integrator, bsdf, win, sampler, bundle_size, size, batch_dims,
                                 with_noise, w_isect, background, device)
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
            vals, it = _render_tile(shapes, lights, cameras, integrator, bsdf, (x0, y0, chunk_size, chunk_size), sampler,
                                    bundle_size, size, batch_dims, with_noise, w_isect, background, device)
            if mode == "crop":
                out[:, x0 - u:x0 - u + chunk_size, y0 - v:y0 - v + chunk_size] = vals
            else:
                out[:, x0:x0 + chunk_size, y0:y0 + chunk_size] = vals
    if squeeze_first and batch_dims == 1:
        out = out.squeeze(0)
    setattr(it, "crop_uv", uv)
    return out, addition(it)
