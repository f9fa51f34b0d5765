"""GPU tests of SURVEY f4 / a21: the ray-generation kernel (nrt_camera_rays) against rays produced by the UNMODIFIED
reference cameras (tests/golden/camera_rays.npz, cameras.npz; pathtracer/cameras/cameras.py:23-54, 132-192,
renderer/cameras.py:539-575), and the camera-driven whole-frame render (nrt_nerfle_render_camera, the fast path of
pathtrace) against the reference's own tiled pathtrace of the same NeRFLE (main.py:13-93)."""
import os
import random
import sys

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
sys.path.insert(0, HERE)
sys.path.insert(0, os.path.join(HERE, "golden"))

import helpers  # noqa: E402
import scenes  # noqa: E402
import synth  # noqa: E402

pytestmark = pytest.mark.gpu

G = np.load(os.path.join(HERE, "golden", "camera_rays.npz"))
GC = np.load(os.path.join(HERE, "golden", "cameras.npz"))


def _window_positions(x0, y0, nx, ny):
    import torch
    gx, gy = torch.meshgrid(torch.arange(x0, x0 + nx, dtype=torch.float), torch.arange(y0, y0 + ny, dtype=torch.float),
                            indexing="ij")
    return torch.stack([gy, gx], dim=-1).cuda()


class _Sampler:
    def sample(self, shape, device="cpu"):
        import torch
        return torch.rand(shape, device=device)


def _nerf_cam(n=3, size=16):
    from neural_raytracing_b200.pathtracer.cameras import NeRFCamera
    c2w, focal = synth.nerf_cameras(n, size, device="cuda")
    return NeRFCamera(cam_to_world=c2w, focal=focal, device="cuda")


def test_nerf_camera_kernel_matches_reference_rays():
    import torch
    from neural_raytracing_b200 import ops
    cam = _nerf_cam()
    x0, y0, nx, ny = (int(v) for v in G["window"])
    win = ops.camera_rays(cam.device_desc(16, x0=x0, y0=y0, nx=nx, ny=ny))
    assert tuple(win.shape) == G["nerf_rays"].shape
    assert np.abs(win.cpu().numpy() - G["nerf_rays"]).max() < 1e-6
    # explicit positions (what sample_positions passes) == the window grid, bit for bit
    via_pos = cam.sample_positions(_window_positions(x0, y0, nx, ny), _Sampler(), bundle_size=1, size=16, N=3)
    assert torch.equal(via_pos, win)
    # a [n,3,4] matrix (row stride 4, view stride 12) reads the same camera
    cam34 = type(cam)(cam_to_world=cam.cam_to_world[:, :3, :].contiguous(), focal=cam.focal, device="cuda")
    assert torch.equal(ops.camera_rays(cam34.device_desc(16, x0=x0, y0=y0, nx=nx, ny=ny)), win)


def test_dtu_camera_kernel_matches_reference_rays():
    import torch
    from neural_raytracing_b200 import ops
    from neural_raytracing_b200.pathtracer.cameras import DTUCamera
    pose, K = scenes.dtu_cameras(2, device="cuda")
    cam = DTUCamera(pose=pose, intrinsic=K, device="cuda")
    x0, y0, nx, ny = (int(v) for v in G["window"])
    got = cam.sample_positions(_window_positions(x0, y0, nx, ny), _Sampler(), bundle_size=2, size=16, N=2)
    assert tuple(got.shape) == G["dtu_rays"].shape
    assert np.abs(got.cpu().numpy() - G["dtu_rays"]).max() < 2e-6
    assert torch.equal(ops.camera_rays(cam.device_desc(16, x0=x0, y0=y0, nx=nx, ny=ny, bundle_size=2)), got)


def test_fov_camera_kernel_matches_reference_rays():
    import torch
    from neural_raytracing_b200.renderer import OpenGLPerspectiveCameras
    cams = OpenGLPerspectiveCameras(device="cuda", R=torch.from_numpy(GC["R"]).cuda(), T=torch.from_numpy(GC["T"]).cuda())
    pos = torch.from_numpy(GC["positions"]).cuda()
    rays = cams.sample_positions(pos, _Sampler(), bundle_size=2, size=16, N=5, with_noise=False)
    assert tuple(rays.shape) == GC["rays"].shape
    assert np.abs(rays.cpu().numpy() - GC["rays"]).max() < 2e-5          # two 4x4 inverses in fp32, as on the CPU
    cam1 = OpenGLPerspectiveCameras(device="cuda", R=torch.from_numpy(GC["R2"]).cuda(), T=torch.from_numpy(GC["T2"]).cuda(),
                                    fov=45.0, znear=0.5, zfar=20.0)
    r1 = cam1.sample_positions(pos, _Sampler(), bundle_size=1, size=16, N=1, with_noise=False)
    assert np.abs(r1.cpu().numpy() - GC["rays_fov45"]).max() < 2e-5
    # jitter through the caller's sampler: seeded, bounded, origins untouched (as the CPU test of the torch expression)
    torch.manual_seed(3)
    a = cams.sample_positions(pos, _Sampler(), bundle_size=4, size=16, N=5, with_noise=1e-2)
    torch.manual_seed(3)
    b = cams.sample_positions(pos, _Sampler(), bundle_size=4, size=16, N=5, with_noise=1e-2)
    clean = cams.sample_positions(pos, _Sampler(), bundle_size=4, size=16, N=5, with_noise=False)
    assert torch.equal(a, b) and tuple(a.shape) == (5, 8, 8, 4, 6)
    assert 0 < (a[..., 3:] - clean[..., 3:]).abs().max().item() < 2e-3
    assert (a[..., :3] - clean[..., :3]).abs().max().item() == 0.0


def test_camera_rays_subrange_view_index_and_hash_jitter():
    import ctypes
    import torch
    from neural_raytracing_b200 import _native as N, ops
    cam = _nerf_cam()
    desc = cam.device_desc(16, x0=1, y0=0, nx=7, ny=9)
    full, view = ops.camera_rays(desc, want_view=True)
    assert torch.equal(view.cpu(), torch.arange(3, dtype=torch.int32).repeat_interleave(63))
    part = torch.empty((50, 6), device="cuda")
    N.check(N.lib().nrt_camera_rays(ctypes.byref(desc.struct), 37, 50, ctypes.c_void_p(part.data_ptr()), None, None))
    torch.cuda.synchronize()
    assert torch.equal(part, full.reshape(-1, 6)[37:87])
    # the library's own pixel jitter: reproducible per seed, different across seeds, within +-jitter/2 pixels
    j1 = ops.camera_rays(cam.device_desc(16, x0=1, y0=0, nx=7, ny=9, jitter=1e-2, jitter_seed=5))
    j2 = ops.camera_rays(cam.device_desc(16, x0=1, y0=0, nx=7, ny=9, jitter=1e-2, jitter_seed=5))
    j3 = ops.camera_rays(cam.device_desc(16, x0=1, y0=0, nx=7, ny=9, jitter=1e-2, jitter_seed=6))
    assert torch.equal(j1, j2) and not torch.equal(j1, j3)
    assert 0 < (j1[..., 3:] - full[..., 3:]).abs().max().item() < 1e-2 / cam.focal
    assert torch.equal(j1[..., :3], full[..., :3])


def test_camera_errors():
    import ctypes
    import torch
    from neural_raytracing_b200 import _native as N, ops
    cam = _nerf_cam()
    desc = cam.device_desc(16, nx=4, ny=4)
    out = torch.empty((16 * 3, 6), device="cuda")
    with pytest.raises(ops.NrtError):       # range outside the block
        N.check(N.lib().nrt_camera_rays(ctypes.byref(desc.struct), 40, 16, ctypes.c_void_p(out.data_ptr()), None, None))
    desc.struct.kind = 7
    with pytest.raises(ops.NrtError):
        N.check(N.lib().nrt_camera_rays(ctypes.byref(desc.struct), 0, 16, ctypes.c_void_p(out.data_ptr()), None, None))
    with pytest.raises(ops.NrtError):       # DTU without intrinsics
        ops.CameraDesc(ops.CAM_DTU, cam.cam_to_world, None, size=16, nx=4, ny=4)
    with pytest.raises(ops.NrtError):       # CPU matrices: there is no CPU path
        ops.CameraDesc(ops.CAM_NERF, cam.cam_to_world.cpu(), None, focal=1.0, size=16, nx=4, ny=4)
    with pytest.raises(ops.NrtError):       # a rotation without the translation column is not a camera-to-world matrix
        ops.camera_rays(ops.CameraDesc(ops.CAM_NERF, cam.cam_to_world[:, :3, :3].contiguous(), None, focal=1.0, size=16,
                                       nx=4, ny=4))
    # empty window
    assert ops.camera_rays(cam.device_desc(16, nx=0, ny=4)).numel() == 0


def _golden_nerfle():
    import torch
    from neural_raytracing_b200.pathtracer.lights import PointLights
    from neural_raytracing_b200.pathtracer.shapes.nerf import NeRFLE
    random.random = lambda: float(G["fixed_random"])
    n = NeRFLE(envmap=False, device="cuda")
    w1, w2 = helpers.nerfle_weights(False)
    for mod, w in ((n.first, w1), (n.second, w2)):
        mod.basis_p = torch.from_numpy(w["basis"]).cuda()
        for lin, W, b in zip([mod.init] + list(mod.layers) + [mod.out], w["W"], w["b"]):
            with torch.no_grad():
                lin.weight.copy_(torch.from_numpy(W)); lin.bias.copy_(torch.from_numpy(b))
    lights = PointLights(device="cuda", location=torch.from_numpy(G["light_loc"]).cuda(), scale=10)
    return n, lights


@pytest.mark.parametrize("prec,tol", [("f32", 1e-4), ("f16", 2e-3)])
def test_pathtrace_camera_frame_matches_reference_pathtrace(prec, tol):
    """pathtrace(NeRFLE, NeRFReproduce) takes the one-call camera path and reproduces the reference's tiled frames."""
    import torch
    import neural_raytracing_b200.pathtracer as P
    from neural_raytracing_b200 import config, ops
    from neural_raytracing_b200.pathtracer.cameras import NeRFCamera
    from neural_raytracing_b200.pathtracer.integrators import NeRFReproduce
    from neural_raytracing_b200.renderer import OpenGLPerspectiveCameras
    real_random = random.random
    try:
        n, lights = _golden_nerfle()
        config.set_precision(prec)
        c2w, focal = synth.nerf_cameras(2, 16, device="cuda")
        cam = NeRFCamera(cam_to_world=c2w, focal=focal, device="cuda")
        fcam = OpenGLPerspectiveCameras(device="cuda", R=torch.from_numpy(G["fov_R"]).cuda(),
                                        T=torch.from_numpy(G["fov_T"]).cuda())
        for cams, bundle, key in ((cam, 1, "frame_nerf"), (fcam, 2, "frame_fov")):
            ops.profile_collect()
            with torch.no_grad():
                img, extra = P.pathtrace(n, size=16, chunk_size=8, bundle_size=bundle, bsdf=None, integrator=NeRFReproduce(),
                                         lights=lights, cameras=cams, device="cuda", silent=True, with_noise=False)
            counts = {k: c for k, (_, c) in ops.profile_collect().items() if c}
            assert extra is None and tuple(img.shape) == G[key].shape
            assert counts.get("camera_rays") == 1, counts          # one frame = one ray-generation launch, no tiles
            assert np.abs(img.cpu().numpy() - G[key]).max() < tol, key
            # the tile loop (forced by an `addition` callback) renders the same image
            with torch.no_grad():
                tiled, _ = P.pathtrace(n, size=16, chunk_size=8, bundle_size=bundle, bsdf=None, integrator=NeRFReproduce(),
                                       lights=lights, cameras=cams, device="cuda", silent=True, with_noise=False,
                                       addition=lambda it: it)
            assert (tiled - img).abs().max().item() < (1e-6 if prec == "f32" else tol)
            # default pixel jitter (1e-3 pixel): a valid image very close to the clean one
            with torch.no_grad():
                noisy, _ = P.pathtrace(n, size=16, chunk_size=8, bundle_size=bundle, bsdf=None, integrator=NeRFReproduce(),
                                       lights=lights, cameras=cams, device="cuda", silent=True)
            assert torch.isfinite(noisy).all() and (noisy - img).abs().max().item() < 5e-3
    finally:
        config.set_precision("f32")
        random.random = real_random


@pytest.mark.parametrize("prec", ["f32", "f16"])
def test_render_camera_equals_render_of_generated_rays(prec):
    """The in-library ray generation changes nothing: render_camera == render(camera_rays), bit for bit, including
    hierarchical sampling, several views (light code row = view) and more rays than one 262,144-ray chunk."""
    import torch
    from neural_raytracing_b200 import ops
    w1, w2 = helpers.nerfle_weights(False)
    first, second = helpers.cuda_mlp(w1), helpers.cuda_mlp(w2)
    cam = _nerf_cam(n=2, size=400)
    code = torch.tensor([[0.4, 1.0, 0.3], [-0.8, 0.5, 0.6]], device="cuda")
    cases = [dict(nx=24, ny=20, kw=dict(n_coarse=16, n_fine=32, t_near=0.1, t_far=2.0, jitter_seed=9))]
    if prec != "f32":
        cases.append(dict(nx=400, ny=400, kw=dict(n_coarse=8, t_near=0.1, t_far=2.0)))       # 320,000 rays: two chunks
    for c in cases:
        desc = cam.device_desc(400, x0=3, y0=5, nx=c["nx"], ny=c["ny"])
        rays, view = ops.camera_rays(desc, want_view=True)
        a = ops.nerfle_render_camera(first, second, desc, None, code, prec=prec, **c["kw"])
        b = ops.nerfle_render(first, second, rays, None, code, view, prec=prec, **c["kw"])
        assert tuple(a.shape) == (2, c["nx"], c["ny"], 1, 3)
        assert torch.equal(a, b)
        assert (a[0] - a[1]).abs().max().item() > 0          # the two views differ (pose and light)
    # the camera-fed kernels (rays computed in the MLP kernels' prologues, no ray array): same image bit for bit, and no
    # ray-generation launch
    if prec != "f32":
        desc = cam.device_desc(400, x0=3, y0=5, nx=24, ny=20)
        kw = dict(n_coarse=16, n_fine=32, t_near=0.1, t_far=2.0, jitter_seed=9)
        a = ops.nerfle_render_camera(first, second, desc, None, code, prec=prec, **kw)
        try:
            ops.set_camera_rays_mode(True)
            ops.profile_collect()
            b = ops.nerfle_render_camera(first, second, desc, None, code, prec=prec, **kw)
            counts = {k: c for k, (_, c) in ops.profile_collect().items() if c}
        finally:
            ops.set_camera_rays_mode(False)
        assert torch.equal(a, b)
        assert "camera_rays" not in counts and counts["mlp_tc_nerf_first"] == 2, counts
    # host-image variant: same pixels in pinned memory
    desc = cam.device_desc(400, x0=3, y0=5, nx=24, ny=20)
    host = torch.empty((2, 24, 20, 1, 3), dtype=torch.float32).pin_memory()
    ts = torch.linspace(0.1, 2.0, 32)
    ops.nerfle_render_camera_host(first, second, desc, ts.pin_memory(), code, host, prec=prec)
    dev = ops.nerfle_render_camera(first, second, desc, ts.cuda(), code, prec=prec)
    assert torch.equal(host, dev.cpu())


def test_single_light_broadcasts_over_views():
    """nerf.py:199-201 expands the light code over the batch: ONE light for several views is legal in the reference (`expand`).
    The fused paths index the code by view, so a single row is broadcast first -- from rays and from a camera."""
    import torch
    from neural_raytracing_b200.pathtracer.lights import PointLights
    real_random = random.random
    try:
        n, _two_lights = _golden_nerfle()
        cam = _nerf_cam(n=3, size=16)
        one = PointLights(device="cuda", location=torch.tensor([[0.4, 1.0, 0.3]], device="cuda"), scale=10)
        three = PointLights(device="cuda", location=torch.tensor([[0.4, 1.0, 0.3]] * 3, device="cuda"), scale=10)
        rays = cam.sample_positions(_window_positions(0, 0, 16, 16), _Sampler(), size=16, N=3)
        with torch.no_grad():
            a, b = n(rays, one), n(rays, three)
            desc = cam.device_desc(16, nx=16, ny=16)
            c, d = n.render_camera(desc, one), n.render_camera(desc, three)
        assert tuple(a.shape) == (3, 16, 16, 1, 3) and torch.equal(a, b) and torch.equal(c, d) and torch.equal(a, c)
        assert (a[0] - a[1]).abs().max().item() > 0
        two = PointLights(device="cuda", location=torch.tensor([[0.4, 1.0, 0.3]] * 2, device="cuda"), scale=10)
        with pytest.raises(ValueError):
            n(rays, two)
    finally:
        random.random = real_random
