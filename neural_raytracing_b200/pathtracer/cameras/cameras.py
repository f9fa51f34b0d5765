"""Ray generators specific to the path tracer (pytorch3d/pathtracer/cameras/cameras.py)."""
from dataclasses import dataclass

import torch
import torch.nn.functional as F

from ... import ops


def _on_device(*ts):
    """The ray-generation kernel serves fixed (non-differentiated) fp32 CUDA cameras; anything else keeps the torch
    expression (learned poses need autograd through it)."""
    return all(isinstance(t, torch.Tensor) and t.is_cuda and t.dtype == torch.float32
               and not (torch.is_grad_enabled() and t.requires_grad) for t in ts)


@dataclass
class Camera:
    camera_to_world = None
    world_to_camera = None

    def sample_positions(self, positions, sampler, bundle_size):
        raise NotImplementedError()


@dataclass
class NeRFCamera(Camera):
    """cameras.py:17-54: pixel grid -> rays [N,W,H,1,6] from NeRF-style camera-to-world matrices."""
    cam_to_world: torch.Tensor = None
    focal: float = None
    device: str = "cuda"

    def __len__(self):
        return self.cam_to_world.shape[0]

    def device_desc(self, size, x0=0, y0=0, nx=0, ny=0, bundle_size=1, positions=None, jitter=0.0, jitter_seed=0):
        """nrt_camera_t of this camera for the window / positions (ops.CameraDesc; one ray per pixel, cameras.py:50),
        or None when the kernel does not apply (CPU tensors, a pose that requires grad, a tensor focal)."""
        if not _on_device(self.cam_to_world) or isinstance(self.focal, torch.Tensor):
            return None
        return ops.CameraDesc(ops.CAM_NERF, self.cam_to_world, None, focal=self.focal, size=size, x0=x0, y0=y0, nx=nx,
                              ny=ny, bundle=1, positions=positions, jitter=jitter, jitter_seed=jitter_seed)

    def sample_positions(self, position_samples, sampler, bundle_size=4, size=512, with_noise=False, N=1):
        u, v = position_samples.split(1, dim=-1)
        if with_noise:
            u = u + (torch.rand_like(u) - 0.5) * with_noise
            v = v + (torch.rand_like(v) - 0.5) * with_noise
        if position_samples.dim() == 3 and _on_device(position_samples, self.cam_to_world):
            # one kernel instead of ~12 elementwise launches (nrt_camera_rays); the jitter above keeps torch's generator
            pos = torch.cat([u, v], dim=-1) if with_noise else position_samples
            desc = self.device_desc(size, nx=pos.shape[0], ny=pos.shape[1], positions=pos)
            if desc is not None:
                return ops.camera_rays(desc)
        d = torch.stack([(u - size * 0.5) / self.focal, -(v - size * 0.5) / self.focal, -torch.ones_like(u)], dim=-1)
        r_d = torch.sum(d[..., None, :] * self.cam_to_world[..., :3, :3], dim=-1)
        r_d = F.normalize(r_d, dim=-1).permute(2, 0, 1, 3).unsqueeze(-2)
        r_o = self.cam_to_world[..., :3, -1][:, None, None, None, :].expand_as(r_d)
        return torch.cat([r_o, r_d], dim=-1)


def lift(x, y, z, intrinsics, size):
    """cameras.py:132-147 (IDR's pixel lifting)."""
    shape = x.shape
    fx = intrinsics[..., 0, 0, None].expand(shape)
    fy = intrinsics[..., 1, 1, None].expand(shape)
    cx = intrinsics[..., 0, 2, None].expand(shape)
    cy = intrinsics[..., 1, 2, None].expand(shape)
    sk = intrinsics[..., 0, 1, None].expand(shape)
    x, y, z = x.expand(shape), y.expand(shape), z.expand(shape)
    x_lift = (x - cx + cy * sk / fy - sk * y / fy) / fx * z
    y_lift = (y - cy) / fy * z
    return torch.stack([x_lift, y_lift, z, torch.ones_like(z)], dim=-1)


@dataclass
class DTUCamera(Camera):
    """cameras.py:151-192: DTU / IDR cameras (pose + intrinsics; 1600x1200 normalisation)."""
    pose: torch.Tensor = None
    intrinsic: torch.Tensor = None
    device: str = "cuda"

    def __len__(self):
        return self.pose.shape[0]

    def device_desc(self, size, x0=0, y0=0, nx=0, ny=0, bundle_size=1, positions=None, jitter=0.0, jitter_seed=0):
        """nrt_camera_t of this camera (see NeRFCamera.device_desc).  The reference's DTU generator takes no pixel
        jitter (cameras.py:156-192 never reads with_noise), so none is applied here either."""
        if not _on_device(self.pose, self.intrinsic) or self.pose.dim() != 3 or self.pose.shape[1] == 7 \
                or self.intrinsic.dim() != 3:
            return None
        return ops.CameraDesc(ops.CAM_DTU, self.pose, self.intrinsic, size=size, x0=x0, y0=y0, nx=nx, ny=ny,
                              bundle=bundle_size, positions=positions)

    def sample_positions(self, position_samples, sampler, bundle_size=4, size=512, with_noise=False, N=1):
        pose, intrinsic = self.pose, self.intrinsic
        assert pose.shape[1] != 7, "quaternion poses are not supported (neither in the reference)"
        if position_samples.dim() == 3 and _on_device(position_samples):
            desc = self.device_desc(size, nx=position_samples.shape[0], ny=position_samples.shape[1],
                                    bundle_size=bundle_size, positions=position_samples)
            if desc is not None:
                return ops.camera_rays(desc)
        r_o = pose[:, :3, 3]
        W, H, _ = position_samples.shape
        N = len(self)
        normalize = torch.tensor([1600, 1200], device=self.device, dtype=torch.float) / size
        u, v = (position_samples * normalize).reshape(-1, 2).split(1, dim=-1)
        u = u.reshape(1, -1).expand(N, -1)
        v = v.reshape(1, -1).expand(N, -1)
        pts = lift(u, v, torch.ones_like(u), intrinsics=intrinsic, size=size)
        world = torch.bmm(pose, pts.permute(0, 2, 1)).permute(0, 2, 1)[..., :3]
        r_o = r_o[:, None, :].expand_as(world)
        r_d = F.normalize(world - r_o, dim=-1)
        return torch.cat([r_o, r_d], dim=-1).reshape(N, W, H, 1, 6).expand(N, W, H, bundle_size, 6)
