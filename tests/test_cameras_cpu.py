"""CPU tests (-m "not gpu") of the perspective camera + look-at helpers against fixtures produced by the UNMODIFIED
reference (tests/golden/make_golden.py::gen_cameras; renderer/cameras.py:539-575, 1284-1422).  Ray generation is pure
torch (no kernel), so it runs wherever torch runs."""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))

from neural_raytracing_b200.renderer import (FoVPerspectiveCameras, OpenGLPerspectiveCameras,  # noqa: E402
                                              look_at_view_transform)

sys.path.insert(0, os.path.join(HERE, "golden"))
G = np.load(os.path.join(HERE, "golden", "cameras.npz"))
GR = np.load(os.path.join(HERE, "golden", "camera_rays.npz"))


class _Sampler:
    def sample(self, shape, device="cpu"):
        return torch.rand(shape, device=device)


def test_look_at_view_transform_matches_reference():
    R, T = look_at_view_transform(dist=1.0, elev=torch.from_numpy(G["elev"]), azim=torch.from_numpy(G["azim"]))
    assert np.abs(R.numpy() - G["R"]).max() < 1e-6 and np.abs(T.numpy() - G["T"]).max() < 1e-6
    R2, T2 = look_at_view_transform(dist=2.7, elev=30.0, azim=200.0, at=((0.1, -0.2, 0.05),))
    assert np.abs(R2.numpy() - G["R2"]).max() < 1e-6 and np.abs(T2.numpy() - G["T2"]).max() < 1e-6
    # eye given; the first camera looks along `up` (degenerate x axis is repaired like upstream)
    R3, T3 = look_at_view_transform(eye=((0.0, 1.5, 0.0), (0.4, 0.3, -1.0)), at=((0.0, 0.0, 0.0),))
    assert np.abs(R3.numpy() - G["R3"]).max() < 1e-6 and np.abs(T3.numpy() - G["T3"]).max() < 1e-6


def test_sample_positions_matches_reference():
    cams = OpenGLPerspectiveCameras(device="cpu", R=torch.from_numpy(G["R"]), T=torch.from_numpy(G["T"]))
    assert len(cams) == 5 and isinstance(cams, FoVPerspectiveCameras)
    assert np.abs(cams.get_camera_center().numpy() - G["centers"]).max() < 1e-6
    pos = torch.from_numpy(G["positions"])
    rays = cams.sample_positions(pos, _Sampler(), bundle_size=2, size=16, N=5, with_noise=False)
    assert tuple(rays.shape) == G["rays"].shape
    assert np.abs(rays.numpy() - G["rays"]).max() < 2e-5          # two 4x4 inverses in fp32
    cam1 = OpenGLPerspectiveCameras(device="cpu", R=torch.from_numpy(G["R2"]), T=torch.from_numpy(G["T2"]), fov=45.0,
                                    znear=0.5, zfar=20.0)
    r1 = cam1.sample_positions(pos, _Sampler(), bundle_size=1, size=16, N=1, with_noise=False)
    assert np.abs(r1.numpy() - G["rays_fov45"]).max() < 2e-5
    assert np.abs(cam1.get_camera_center().numpy() - G["center_fov45"]).max() < 1e-6


def test_sample_positions_jitter_is_bounded_and_seeded():
    cams = OpenGLPerspectiveCameras(device="cpu", R=torch.from_numpy(G["R"][:1]), T=torch.from_numpy(G["T"][:1]))
    pos = torch.from_numpy(G["positions"])
    torch.manual_seed(3)
    a = cams.sample_positions(pos, _Sampler(), bundle_size=4, size=16, N=1, with_noise=1e-2)
    torch.manual_seed(3)
    b = cams.sample_positions(pos, _Sampler(), bundle_size=4, size=16, N=1, with_noise=1e-2)
    assert torch.equal(a, b) and tuple(a.shape) == (1, 8, 8, 4, 6)
    clean = cams.sample_positions(pos, _Sampler(), bundle_size=4, size=16, N=1, with_noise=False)
    assert (a[..., 3:] - clean[..., 3:]).abs().max().item() < 2e-3      # +-0.005 pixel of a 16-pixel image
    assert (a[..., :3] - clean[..., :3]).abs().max().item() == 0.0


def test_nerf_and_dtu_cameras_match_reference_rays():
    """The torch expressions of NeRFCamera / DTUCamera (the path CPU tensors and learned poses take) against rays of the
    unmodified reference on a 10 x 6 pixel window (tests/golden/make_golden.py::gen_camera_rays)."""
    import scenes
    import synth
    from neural_raytracing_b200.pathtracer.cameras import DTUCamera, NeRFCamera
    x0, y0, nx, ny = (int(v) for v in GR["window"])
    gx, gy = torch.meshgrid(torch.arange(x0, x0 + nx, dtype=torch.float), torch.arange(y0, y0 + ny, dtype=torch.float),
                            indexing="ij")
    pos = torch.stack([gy, gx], dim=-1)
    c2w, focal = synth.nerf_cameras(3, 16)
    rays = NeRFCamera(cam_to_world=c2w, focal=focal, device="cpu").sample_positions(pos, _Sampler(), bundle_size=1, size=16, N=3)
    assert tuple(rays.shape) == GR["nerf_rays"].shape and np.abs(rays.numpy() - GR["nerf_rays"]).max() < 1e-6
    pose, K = scenes.dtu_cameras(2, device="cpu")
    rays = DTUCamera(pose=pose, intrinsic=K, device="cpu").sample_positions(pos, _Sampler(), bundle_size=2, size=16, N=2)
    assert tuple(rays.shape) == GR["dtu_rays"].shape and np.abs(rays.numpy() - GR["dtu_rays"]).max() < 1e-6
