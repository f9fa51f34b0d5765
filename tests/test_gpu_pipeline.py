"""End-to-end drop-in tests (-m gpu): the mirror package renders the same scenes as the UNMODIFIED
reference (tests/golden/pipeline.npz, produced by tests/golden/make_golden.py::gen_pipeline with the
scene code of tests/golden/scenes.py) through the reference's own entry points
pathtrace / pathtrace_sample (main.py:13-179).

Tolerance: everything is fp32; per-pixel differences come from fp32 summation order (1e-5) except
where a ray sits on a hard threshold of the reference (hit eps 1e-3, shadow test, conductor lobe
0.94): such pixels may flip, so the image test allows a small fraction of outliers and the PSNR bar
is the north_star's 50 dB."""
import random

import numpy as np
import pytest

import helpers
import scenes
import synth

pytestmark = pytest.mark.gpu


def _setup():
    import torch
    import neural_raytracing_b200.pathtracer as P
    from neural_raytracing_b200 import config
    config.set_precision("f32")
    g = helpers.golden("pipeline")
    random.random = lambda: float(g["fixed_random"])     # far-plane jitter of sdfs.py:236
    return torch, P, g


def test_colocate_style_pathtrace_matches_reference():
    torch, P, g = _setup()
    from neural_raytracing_b200.pathtracer.cameras import NeRFCamera
    size = 16
    shape, sphere, bsdf, lights, integrator, w_isect = scenes.build_pipeline(P, "colocate", device="cuda")
    c2w, focal = synth.nerf_cameras(1, size, device="cuda")
    cam = NeRFCamera(cam_to_world=c2w, focal=focal, device="cuda")
    with torch.no_grad():
        img, mi = P.pathtrace(shape, size=size, chunk_size=size, bundle_size=1, bsdf=bsdf, integrator=integrator,
                              lights=lights, cameras=cam, device="cuda", silent=True, background=0, w_isect=w_isect,
                              with_noise=False, addition=lambda it: it)
    img = img.cpu().numpy()
    ref = g["colocate_img"]
    assert img.shape == ref.shape
    err = np.abs(img - ref).max(axis=-1)
    assert (err < 1e-3).mean() >= 0.97, (err.max(), (err < 1e-3).mean())
    assert helpers.psnr(img, ref) > 50
    thr = mi.throughput.reshape(-1).cpu().numpy()
    tref = g["colocate_throughput"]
    close = np.abs(thr - tref) < 2e-2 * np.maximum(1.0, np.abs(tref))      # logit scale 1000
    assert close.mean() > 0.97
    w = mi.normalized_weights.reshape(-1, 4).cpu().numpy()
    hit_rows = np.abs(w - g["colocate_weights"]).max(axis=-1) < 2e-3
    assert hit_rows.mean() > 0.97
    assert abs(mi.raw_normals.shape[0] - g["colocate_raw_normals"].shape[0]) <= 2


def test_dtu_style_training_step_matches_reference():
    """pathtrace_sample + loss + backward: image, throughput, loss value and parameter gradients."""
    torch, P, g = _setup()
    from neural_raytracing_b200.pathtracer.cameras import NeRFCamera
    from neural_raytracing_b200.pathtracer.utils import eikonal_loss
    size = 16
    shape, sphere, bsdf, lights, integrator, w_isect = scenes.build_pipeline(P, "dtu", device="cuda")
    c2w, focal = synth.nerf_cameras(2, size, device="cuda")
    cam = NeRFCamera(cam_to_world=c2w, focal=focal, device="cuda")
    got, mi = P.pathtrace_sample(shape, size=size, chunk_size=size, bundle_size=1, crop_size=8, uv=(3, 5), bsdf=bsdf,
                                 integrator=integrator, lights=lights, cameras=cam, device="cuda", silent=True,
                                 background=0, w_isect=w_isect, with_noise=False, addition=lambda it: it,
                                 squeeze_first=False)
    ref = g["dtu_img"]
    assert tuple(got.shape) == ref.shape
    err = np.abs(got.detach().cpu().numpy() - ref).max(axis=-1)
    assert (err < 1e-3).mean() >= 0.97, (err.max(), (err < 1e-3).mean())
    loss = (got[..., :3] - 0.5).square().mean() + 0.1 * eikonal_loss(mi.raw_normals) + \
        torch.nn.functional.binary_cross_entropy_with_logits(
            mi.throughput.reshape(-1), torch.ones(mi.throughput.numel(), device="cuda") * 0.5)
    loss.backward()
    assert abs(loss.item() - float(g["dtu_loss"])) < 2e-3 * abs(float(g["dtu_loss"]))

    def check(name, t, rtol=2e-2):
        a, b = t.detach().cpu().numpy().ravel(), g[name].ravel()
        cos = float(a @ b / (np.linalg.norm(a) * np.linalg.norm(b) + 1e-30))
        assert cos > 0.999, (name, cos)
        assert np.abs(a - b).max() <= rtol * np.abs(b).max() + 1e-7, (name, np.abs(a - b).max(), np.abs(b).max())

    check("dtu_g_sdf_out_w", sphere.shift.out.weight.grad)
    check("dtu_g_sdf_l3_w", sphere.shift.layers[3].weight.grad)
    check("dtu_g_centers", sphere.centers.grad)
    check("dtu_g_bsdf0_init_w", bsdf.bsdfs[0].mlp.init.weight.grad)
    check("dtu_g_spvar_out_w", bsdf.sp_var_fn.out.weight.grad)
    check("dtu_g_light_out_w", lights.light_field_approx.out.weight.grad)
    check("dtu_g_light_color", lights.color.grad)
    check("dtu_g_reflectance", bsdf.bsdfs[2].reflectance.grad)


def test_nerfle_through_pathtrace_sample():
    """nerfle.py-style call: NeRFReproduce integrator, fused fp32 render vs the reference's NeRFLE output."""
    torch, P, g0 = _setup()
    from neural_raytracing_b200.pathtracer.shapes.nerf import NeRFLE
    from neural_raytracing_b200.pathtracer.lights import PointLights
    g = helpers.golden("nerfle")
    random.random = lambda: float(g["fixed_random"])
    n = NeRFLE(device="cuda")
    w1, w2 = helpers.nerfle_weights(False)

    def load(mod, w):
        mod.basis_p = torch.from_numpy(w["basis"]).cuda()
        for lin, W, b in zip([mod.init] + list(mod.layers) + [mod.out], w["W"], w["b"]):
            with torch.no_grad():
                lin.weight.copy_(torch.from_numpy(W)); lin.bias.copy_(torch.from_numpy(b))
    load(n.first, w1); load(n.second, w2)
    rays = torch.from_numpy(g["pt_rays"]).cuda()
    lights = PointLights(device="cuda", location=torch.from_numpy(g["pt_light_loc"]).cuda(), scale=10)
    with torch.no_grad():
        rgb = n(rays, lights)
    assert tuple(rgb.shape) == g["pt_rgb"].shape
    assert np.abs(rgb.cpu().numpy() - g["pt_rgb"]).max() < 1e-4
    # differentiable path (MLPs + CUDA compositing kernels): same values, gradients flow to both MLPs
    rgb2 = n(rays, lights)
    assert np.abs(rgb2.detach().cpu().numpy() - g["pt_rgb"]).max() < 1e-4
    rgb2.square().mean().backward()
    assert n.first.init.weight.grad.abs().sum() > 0 and n.second.out.weight.grad.abs().sum() > 0


def test_colocate_style_pathtrace_tensor_core_precision():
    """The same colocate-style render with config.set_precision("f16"): march, shadow march, min scan, NeuralBSDF and
    occlusion MLPs on the tcgen05 kernels.  north_star gate for the 16-bit path: >= 50 dB PSNR on the rendered image
    against the UNMODIFIED reference's output, radiance within 1e-3 on (nearly) all pixels."""
    torch, P, g = _setup()
    from neural_raytracing_b200 import config
    from neural_raytracing_b200.pathtracer.cameras import NeRFCamera
    size = 16
    shape, sphere, bsdf, lights, integrator, w_isect = scenes.build_pipeline(P, "colocate", device="cuda")
    c2w, focal = synth.nerf_cameras(1, size, device="cuda")
    cam = NeRFCamera(cam_to_world=c2w, focal=focal, device="cuda")
    try:
        config.set_precision("f16")
        with torch.no_grad():
            img, mi = P.pathtrace(shape, size=size, chunk_size=size, bundle_size=1, bsdf=bsdf, integrator=integrator,
                                  lights=lights, cameras=cam, device="cuda", silent=True, background=0, w_isect=w_isect,
                                  with_noise=False, addition=lambda it: it)
    finally:
        config.set_precision("f32")
    img = img.cpu().numpy()
    ref = g["colocate_img"]
    err = np.abs(img - ref).max(axis=-1)
    assert (err < 1e-3).mean() >= 0.95, (err.max(), (err < 1e-3).mean())
    assert helpers.psnr(img, ref) > 50, helpers.psnr(img, ref)
    assert abs(mi.raw_normals.shape[0] - g["colocate_raw_normals"].shape[0]) <= 2


def test_nerfle_module_tensor_core_precision():
    torch, P, g0 = _setup()
    from neural_raytracing_b200 import config
    from neural_raytracing_b200.pathtracer.shapes.nerf import NeRFLE
    from neural_raytracing_b200.pathtracer.lights import PointLights
    g = helpers.golden("nerfle")
    random.random = lambda: float(g["fixed_random"])
    n = NeRFLE(device="cuda")
    w1, w2 = helpers.nerfle_weights(False)
    for mod, w in ((n.first, w1), (n.second, w2)):
        mod.basis_p = torch.from_numpy(w["basis"]).cuda()
        for lin, W, b in zip([mod.init] + list(mod.layers) + [mod.out], w["W"], w["b"]):
            with torch.no_grad():
                lin.weight.copy_(torch.from_numpy(W)); lin.bias.copy_(torch.from_numpy(b))
    rays = torch.from_numpy(g["pt_rays"]).cuda()
    lights = PointLights(device="cuda", location=torch.from_numpy(g["pt_light_loc"]).cuda(), scale=10)
    try:
        config.set_precision("f16")
        config.set_train_precision("f16")
        with torch.no_grad():
            rgb = n(rays, lights)
        rgb2 = n(rays, lights)          # differentiable: tensor-core training forward + CUDA compositing
        rgb2.square().mean().backward()
    finally:
        config.set_precision("f32")
        config.set_train_precision("f32")
    ref = g["pt_rgb"]
    assert np.abs(rgb.cpu().numpy() - ref).max() < 1e-3 and helpers.psnr(rgb.cpu().numpy(), ref) > 60
    assert np.abs(rgb2.detach().cpu().numpy() - ref).max() < 1e-3
    assert torch.isfinite(n.first.init.weight.grad).all() and n.first.init.weight.grad.abs().sum() > 0
    assert n.second.out.weight.grad.abs().sum() > 0


def test_gradient_free_frames_render_in_row_blocks():
    """pathtrace under no_grad ignores a small chunk_size and renders row blocks of up to config.max_tile_rays rays: the same
    image as the caller's tiles (rays are independent), one march launch instead of sixteen."""
    torch, P, g = _setup()
    from neural_raytracing_b200 import config, ops
    from neural_raytracing_b200.pathtracer.cameras import NeRFCamera
    size = 16
    shape, sphere, bsdf, lights, integrator, w_isect = scenes.build_pipeline(P, "colocate", device="cuda")
    c2w, focal = synth.nerf_cameras(1, size, device="cuda")
    cam = NeRFCamera(cam_to_world=c2w, focal=focal, device="cuda")

    def frame(chunk, trim=0):
        ops.profile_collect()
        with torch.no_grad():
            img, _ = P.pathtrace(shape, size=size, chunk_size=chunk, bundle_size=1, bsdf=bsdf, integrator=integrator,
                                 lights=lights, cameras=cam, device="cuda", silent=True, background=0, w_isect=w_isect,
                                 with_noise=False, trim=trim)
        counts = {k: c for k, (_, c) in ops.profile_collect().items() if c}
        return img, counts.get("sdf_march_f32", 0) + counts.get("sdf_march_tc", 0)
    whole, n_whole = frame(16)
    blocks, n_blocks = frame(4)
    assert n_whole == 1 and n_blocks == 1
    assert torch.equal(blocks, whole)
    assert torch.equal(frame(4, trim=2)[0], whole)
    prev = config.max_tile_rays
    try:
        config.set_max_tile_rays(0)                       # the caller's tiles, as in the reference
        tiles, n_tiles = frame(4)
        config.set_max_tile_rays(100)                     # 6 rows of 16 pixels per call: uneven last block
        uneven, n_uneven = frame(4)
    finally:
        config.set_max_tile_rays(prev)
    assert n_tiles == 16 and n_uneven == 3
    assert (tiles - whole).abs().max().item() < 1e-6 and (uneven - whole).abs().max().item() < 1e-6
    ref = g["colocate_img"]
    assert helpers.psnr(whole.cpu().numpy(), ref) > 50
