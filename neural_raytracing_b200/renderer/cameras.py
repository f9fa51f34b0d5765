"""The part of `pytorch3d.renderer.cameras` that the path tracer's scripts use:
`look_at_view_transform` (renderer/cameras.py:1363-1422) and the perspective camera with the fork's ray generator
`FoVPerspectiveCameras.sample_positions` (renderer/cameras.py:539-575), as colocate.py:54-56 / nerfle.py:96 /
training_utils.py:198 construct them (`OpenGLPerspectiveCameras(device=, R=, T=)`).

Written from the conventions, not from the upstream class hierarchy: PyTorch3D transforms ROW vectors,
    x_view = x_world @ R + T,        x_clip = [x_view, 1] @ P^T,        x_ndc = x_clip[:3] / x_clip[3],
so a camera here is just the batch of 4x4 matrices `world -> clip`; rays come from its inverse.  The matrices are
a few 4x4 torch products per call; on a CUDA device the per-pixel unprojection is one launch of the library's ray
generator (`nrt_camera_rays`, csrc/nrt_camera.cu), which the camera-driven render also runs inside the library (f4).
"""
import math
from typing import Optional, Sequence, Tuple

import torch
import torch.nn.functional as F

from .. import ops


def _as_batch(x, device, width=None):
    t = x if torch.is_tensor(x) else torch.tensor(x, dtype=torch.float32)
    t = t.to(device=device, dtype=torch.float32)
    if t.dim() == 0:
        t = t.reshape(1)
    if width is not None and t.dim() == 1 and t.shape[0] == width:
        t = t.reshape(1, width)
    return t


def _broadcast(*ts):
    n = max(t.shape[0] for t in ts)
    out = []
    for t in ts:
        if t.shape[0] not in (1, n):
            raise ValueError("cannot broadcast batch sizes %s" % ([int(q.shape[0]) for q in ts],))
        out.append(t.expand((n,) + tuple(t.shape[1:])) if t.shape[0] == 1 else t)
    return out


def camera_position_from_spherical_angles(dist, elev, azim, degrees: bool = True, device="cpu") -> torch.Tensor:
    """Position at distance `dist`, elevation `elev` above the xz-plane and azimuth `azim` from +z (y is up)."""
    dist, elev, azim = _broadcast(_as_batch(dist, device), _as_batch(elev, device), _as_batch(azim, device))
    if degrees:
        elev, azim = elev * (math.pi / 180.0), azim * (math.pi / 180.0)
    x = dist * torch.cos(elev) * torch.sin(azim)
    y = dist * torch.sin(elev)
    z = dist * torch.cos(elev) * torch.cos(azim)
    return torch.stack([x, y, z], dim=-1).reshape(-1, 3)


def look_at_rotation(camera_position, at=((0, 0, 0),), up=((0, 1, 0),), device="cpu") -> torch.Tensor:
    """Rotation whose COLUMNS are the camera's x, y, z axes in world coordinates (z looks at `at`).  An `up` parallel
    to the viewing direction gives a zero x axis, which is replaced by normalize(cross(y, z)) like upstream."""
    c, at, up = _broadcast(_as_batch(camera_position, device, 3), _as_batch(at, device, 3), _as_batch(up, device, 3))
    z = F.normalize(at - c, eps=1e-5)
    x = F.normalize(torch.cross(up, z, dim=1), eps=1e-5)
    y = F.normalize(torch.cross(z, x, dim=1), eps=1e-5)
    degenerate = torch.isclose(x, torch.zeros_like(x), atol=5e-3).all(dim=1, keepdim=True)
    x = torch.where(degenerate, F.normalize(torch.cross(y, z, dim=1), eps=1e-5), x)
    return torch.stack([x, y, z], dim=2)     # == cat of the three axes as rows, transposed


def look_at_view_transform(dist=1.0, elev=0.0, azim=0.0, degrees: bool = True, eye: Optional[Sequence] = None,
                           at=((0, 0, 0),), up=((0, 1, 0),), device="cpu") -> Tuple[torch.Tensor, torch.Tensor]:
    """(R, T) of the world -> view transform of cameras looking at `at` (renderer/cameras.py:1363-1422)."""
    at_t = _as_batch(at, device, 3)
    if eye is not None:
        c = _as_batch(eye, device, 3)
    else:
        c = camera_position_from_spherical_angles(dist, elev, azim, degrees=degrees, device=device)
        c, at_b = _broadcast(c, at_t)
        c = c + at_b
    R = look_at_rotation(c, at_t, up, device=device)
    c = c.expand(R.shape[0], 3)
    T = -torch.bmm(R.transpose(1, 2), c[:, :, None])[:, :, 0]
    return R, T


class FoVPerspectiveCameras:
    """Batch of perspective cameras given by field of view (upstream defaults: znear 1, zfar 100, fov 60 degrees,
    R = identity, T = 0) with the ray generator the path tracer calls."""

    def __init__(self, znear=1.0, zfar=100.0, aspect_ratio=1.0, fov=60.0, degrees: bool = True, R=None, T=None,
                 device="cpu"):
        self.device = torch.device(device)
        R = torch.eye(3)[None] if R is None else R
        T = torch.zeros(1, 3) if T is None else T
        R = _as_batch(R, self.device)
        if R.dim() == 2:
            R = R[None]
        T = _as_batch(T, self.device, 3)
        znear, zfar, aspect, fov = (_as_batch(v, self.device) for v in (znear, zfar, aspect_ratio, fov))
        n = max(R.shape[0], T.shape[0], znear.shape[0], zfar.shape[0], aspect.shape[0], fov.shape[0])
        R = R.expand(n, 3, 3) if R.shape[0] == 1 else R
        T, self.znear, self.zfar, self.aspect_ratio, self.fov = _broadcast(T, znear, zfar, aspect, fov) \
            if n > 1 else (T, znear, zfar, aspect, fov)
        if T.shape[0] != n:
            T = T.expand(n, 3)
        self.R, self.T, self.degrees = R, T, degrees
        self._n = n

    def __len__(self):
        return self._n

    def to(self, device):
        self.device = torch.device(device)
        for k in ("R", "T", "znear", "zfar", "aspect_ratio", "fov"):
            setattr(self, k, getattr(self, k).to(self.device))
        return self

    # ---- matrices (row-vector convention) ----------------------------------------------------------
    def world_to_view_matrix(self) -> torch.Tensor:
        m = torch.zeros(self._n, 4, 4, device=self.device)
        m[:, :3, :3] = self.R
        m[:, 3, :3] = self.T
        m[:, 3, 3] = 1.0
        return m

    def projection_matrix(self) -> torch.Tensor:
        """P^T of the OpenGL-style frustum with z in [0, 1] and +z forward."""
        fov = self.fov * (math.pi / 180.0) if self.degrees else self.fov
        znear, zfar = self.znear.expand(self._n), self.zfar.expand(self._n)
        top = torch.tan(fov / 2).expand(self._n) * znear
        right = top * self.aspect_ratio.expand(self._n)
        k = torch.zeros(self._n, 4, 4, device=self.device)
        k[:, 0, 0] = 2.0 * znear / (2.0 * right)
        k[:, 1, 1] = 2.0 * znear / (2.0 * top)
        k[:, 3, 2] = 1.0
        k[:, 2, 2] = zfar / (zfar - znear)
        k[:, 2, 3] = -(zfar * znear) / (zfar - znear)
        return k.transpose(1, 2)

    def full_projection_matrix(self) -> torch.Tensor:
        return torch.bmm(self.world_to_view_matrix(), self.projection_matrix())

    def get_camera_center(self) -> torch.Tensor:
        return torch.inverse(self.world_to_view_matrix())[:, 3, :3]

    def device_desc(self, size, x0=0, y0=0, nx=0, ny=0, bundle_size=1, positions=None, jitter=0.0, jitter_seed=0):
        """nrt_camera_t of these cameras (ops.CameraDesc): the inverse full projection and the camera centres stay
        device tensors (two small torch ops per call, no host synchronisation).  None when the kernel does not apply
        (CPU cameras, or R / T that require grad)."""
        if self.device.type != "cuda" or (torch.is_grad_enabled() and (self.R.requires_grad or self.T.requires_grad)):
            return None
        inv = torch.inverse(self.full_projection_matrix()).float()
        return ops.CameraDesc(ops.CAM_FOV, inv, self.get_camera_center().float(), size=size, x0=x0, y0=y0, nx=nx, ny=ny,
                              bundle=bundle_size, positions=positions, jitter=jitter, jitter_seed=jitter_seed)

    # ---- ray generator (the fork's addition, renderer/cameras.py:539-575) ----------------------------
    def sample_positions(self, position_samples, sampler, bundle_size=8, size=512, with_noise=False, N=1) -> torch.Tensor:
        """position_samples [W,H,2] (pixels) -> rays [N,W,H,bundle,6].  Kept from the reference: the jitter is
        `d*U - d/2`, pixels map to NDC as `1 - 2 p/size`, and the direction is the normalised UNPROJECTED POINT on
        the far plane (ndc z = 1), not that point minus the camera centre."""
        device = position_samples.device
        p = position_samples.unsqueeze(-2).expand(*position_samples.shape[:-1], bundle_size, 2)
        if with_noise:
            d = with_noise
            p = p + (d * sampler.sample(p.shape, device=device) - d / 2)
        if position_samples.dim() == 3 and position_samples.is_cuda and position_samples.dtype == torch.float32:
            # one kernel for the unprojection (nrt_camera_rays); the jitter above keeps the caller's sampler
            desc = self.device_desc(size, nx=position_samples.shape[0], ny=position_samples.shape[1],
                                    bundle_size=bundle_size, positions=p if with_noise else position_samples)
            if desc is not None:
                return ops.camera_rays(desc)
        p = -2 * (p / size) + 1
        pts = torch.cat([p, torch.ones(p.shape[:-1] + (2,), device=device)], dim=-1)      # homogeneous, ndc z = 1
        inv = torch.inverse(self.full_projection_matrix())                                  # [N,4,4]
        flat = pts.reshape(1, -1, 4).expand(self._n, -1, 4)
        world = torch.bmm(flat, inv)
        world = world[..., :3] / world[..., 3:]
        directions = F.normalize(world.reshape((N,) + tuple(pts.shape[:-1]) + (3,)), dim=-1)
        origins = self.get_camera_center()[:, None, None, None, :].expand_as(directions)
        return torch.cat([origins, directions], dim=-1)


def OpenGLPerspectiveCameras(znear=1.0, zfar=100.0, aspect_ratio=1.0, fov=60.0, degrees: bool = True, R=None, T=None,
                             device="cpu"):
    """Upstream's deprecated alias, which is the name the scripts import."""
    return FoVPerspectiveCameras(znear=znear, zfar=zfar, aspect_ratio=aspect_ratio, fov=fov, degrees=degrees, R=R, T=T,
                                 device=device)
