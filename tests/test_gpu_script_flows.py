"""GPU test: the flow of scripts/nerf_synthetic.py (:56-140) line by line on this package, at toy size and with a miniature
dataset on disk -- dataset loader, `SDF(sdf=torch.jit.load(path, device))`, the script's BSDF / light / optimizer set-up, a few
`train_nerf` iterations with its eikonal extra loss and a validation render, `torch.jit.save(density_field.sdf)` +
`torch.save(bsdf / lights)`, then `test_nerf` on the reloaded models.  What it pins: every call and attribute the script touches
exists with the script's signature, the tensor-core path is the one that runs (f16), training changes the archive the script
saves, and the saved models reload to the same renders."""
import os
import random
import sys

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
sys.path.insert(0, HERE)
sys.path.insert(0, os.path.join(HERE, "golden"))

import tiny_datasets  # noqa: E402

pytestmark = pytest.mark.gpu


def test_nerf_synthetic_script_flow(tmp_path):
    import torch
    import torch.nn as nn
    from neural_raytracing_b200 import config, ops
    from neural_raytracing_b200.pathtracer import checkpoint
    from neural_raytracing_b200.pathtracer.bsdf import ComposeSpatialVarying, NeuralBSDF
    from neural_raytracing_b200.pathtracer.integrators import Direct
    from neural_raytracing_b200.pathtracer.lights import LightField
    from neural_raytracing_b200.pathtracer.shapes.sdfs import SDF, SphereSDF
    from neural_raytracing_b200.pathtracer.training_utils import test_nerf, test_nerf_resources, train_nerf
    from neural_raytracing_b200.pathtracer.utils import eikonal_loss, rand_uv
    device, SIZE = "cuda", 16
    random.seed(0); np.random.seed(0); torch.manual_seed(0)
    DIR = tiny_datasets.write_nerf_synthetic(str(tmp_path / "lego"), n_frames=3) + os.sep
    os.makedirs(tmp_path / "models"); os.makedirs(tmp_path / "outputs")
    sdf_path = str(tmp_path / "models" / "lego_sdf_f.pt")
    start = SphereSDF(n=2 << 6, device="cpu")                 # the archive a previous run of the script would have left
    with torch.no_grad():
        start.radii.abs_().add_(0.05)
    checkpoint.save_sdf_archive(start, sdf_path)
    try:
        config.set_precision("f16"); config.set_train_precision("f16")
        # ---- nerf_synthetic.py:56-75
        cam_to_worlds, focal, exp_imgs, exp_masks = test_nerf_resources(DIR, SIZE, device=device)
        integrator = Direct()
        shape = torch.jit.load(sdf_path, device)
        density_field = SDF(sdf=shape)
        density_field.max_steps = 64
        learned_bsdf = ComposeSpatialVarying([NeuralBSDF(activation=nn.Softplus()) for _ in range(8)])
        lights = LightField()
        opt = torch.optim.AdamW([
            {"params": density_field.parameters(), "lr": 8e-5},
            {"params": learned_bsdf.parameters(), "lr": 8e-4},
            {"params": lights.parameters(), "lr": 8e-5},
        ], lr=8e-5, weight_decay=0)

        def extra_loss(mi, got, exp, mask):
            raw_n = getattr(mi, "raw_normals", None)
            if raw_n is None:
                return 0
            return eikonal_loss(raw_n)
        before = {k: v.detach().clone() for k, v in shape.state_dict().items()}
        ops.profile_collect()
        # ---- :86-110
        losses = train_nerf(density_field, bsdf=learned_bsdf, integrator=integrator, lights=lights, focal=focal,
                            cam_to_worlds=cam_to_worlds, exp_imgs=exp_imgs, exp_masks=exp_masks, opt=opt, size=SIZE,
                            crop_size=8, save_freq=5000, valid_freq=2, max_valid_size=SIZE, iters=3, N=2,
                            extra_loss=extra_loss, name_fn=lambda i: str(tmp_path / "outputs" / ("train_%06d.png" % i)),
                            valid_name_fn=lambda i: str(tmp_path / "outputs" / ("valid_%06d.png" % i)), silent=True,
                            uv_select=lambda _, crop_size: rand_uv(SIZE, SIZE, crop_size))
        counts = {k: c for k, (_, c) in ops.profile_collect().items() if c}
        assert len(losses) == 3 and all(np.isfinite(losses))
        assert counts.get("sdf_march_tc", 0) >= 3 and counts.get("sdf_min_scan_tc", 0) >= 3, counts     # fused tensor-core path
        assert "sdf_march_f32" not in counts and not [k_ for k_ in counts if k_.endswith("_f32") and k_ != "sdf_value_grad_f32"]
        changed = [k for k, v in shape.state_dict().items() if not torch.equal(v, before[k])]
        grads = {k: (None if p.grad is None else float(p.grad.abs().sum())) for k, p in shape.named_parameters()}
        assert "centers" in changed and any(k.startswith("shift.") for k in changed), (changed, grads)  # the ARCHIVE trained
        assert os.path.exists(tmp_path / "outputs" / "valid_000000.png")
        # ---- :112-115
        torch.jit.save(density_field.sdf, sdf_path)
        torch.save(learned_bsdf, str(tmp_path / "models" / "lego_bsdf_f.pt"))
        torch.save(lights, str(tmp_path / "models" / "lego_light_f.pt"))
        # ---- :117-135
        density_field.max_steps = 256
        cam_to_worlds, focal, exp_imgs, exp_masks = test_nerf_resources(DIR, SIZE, device=device)
        name_fn = lambda i: str(tmp_path / "outputs" / ("test_%03d.png" % i))        # noqa: E731
        stats = test_nerf(density_field, integrator=integrator, bsdf=learned_bsdf, lights=lights, cam_to_worlds=cam_to_worlds,
                          focal=focal, exp_imgs=exp_imgs, size=SIZE, name_fn=name_fn)
        # a later run of the script: everything reloaded from the files just written gives the same numbers
        field2 = SDF(sdf=torch.jit.load(sdf_path, device))
        field2.max_steps = 256
        bsdf2 = torch.load(str(tmp_path / "models" / "lego_bsdf_f.pt"), weights_only=False)
        lights2 = torch.load(str(tmp_path / "models" / "lego_light_f.pt"), weights_only=False)
        stats2 = test_nerf(field2, integrator=Direct(), bsdf=bsdf2, lights=lights2, cam_to_worlds=cam_to_worlds, focal=focal,
                           exp_imgs=exp_imgs, size=SIZE, name_fn=name_fn)
        assert isinstance(stats, dict) and stats.keys() == stats2.keys()
        for k in stats:
            a, b = np.asarray(stats[k], np.float64), np.asarray(stats2[k], np.float64)
            # (test_nerf renders with pathtrace's default 1e-3-pixel jitter from torch's generator: not the same draw twice)
            assert np.allclose(a, b, rtol=1e-2, atol=1e-4), (k, a, b)
    finally:
        config.set_precision("f32"); config.set_train_precision("f32")


def test_colocate_script_flow(tmp_path):
    """scripts/colocate.py:26-175 on this package at toy size: the cbox relighting layout from disk, `SDF(sdf=torch.jit.load(...))`,
    the script's 2 NeuralBSDF + Diffuse(Softplus) + Conductor(Softplus) BSDF, PointLights(scale=5) with `intensity_parameters()` in
    the optimizer, the learned occlusion MLP passed as `w_isect`, `light_update` moving the light with the camera, `train_sample`
    with the eikonal + weight-spread extra loss, `torch.jit.save` / `torch.save`, then `test(..., w_isect=True)` (shadow rays)."""
    import torch
    import torch.nn as nn
    from neural_raytracing_b200 import config, ops
    from neural_raytracing_b200.pathtracer import checkpoint
    from neural_raytracing_b200.pathtracer.bsdf import ComposeSpatialVarying, Conductor, Diffuse, NeuralBSDF
    from neural_raytracing_b200.pathtracer.integrators import Direct
    from neural_raytracing_b200.pathtracer.lights import PointLights
    from neural_raytracing_b200.pathtracer.neural_blocks import SkipConnMLP
    from neural_raytracing_b200.pathtracer.shapes.sdfs import SDF, SphereSDF
    from neural_raytracing_b200.pathtracer.training_utils import test, test_colocate_resources, train_sample
    from neural_raytracing_b200.pathtracer.utils import eikonal_loss, rand_uv
    device, SIZE, DIST, k = "cuda", 16, 1.0, "bunny"
    random.seed(1); np.random.seed(1); torch.manual_seed(1)
    root = tiny_datasets.write_colocate(str(tmp_path / "mitsuba_scenes" / "cbox_relight"), k)
    os.makedirs(tmp_path / "models"); os.makedirs(tmp_path / "outputs")
    sdf_path = str(tmp_path / "models" / ("col_%s_sdf.pt" % k))
    # the archive an earlier run left (colocate.py:141 saves from the GPU, :65 loads it back without a device argument);
    # written here the way the script's other branch builds its shape: torch.jit.script(SphereSDF(n=2<<5)) (:63)
    start = SphereSDF(n=2 << 5, device=device)
    with torch.no_grad():
        start.radii.abs_().add_(0.05)
    torch.jit.save(torch.jit.script(start), sdf_path)
    try:
        config.set_precision("f16"); config.set_train_precision("f16")
        Rs, Ts, exp_imgs, exp_masks, xyzs = test_colocate_resources(k, SIZE, dist=DIST, device=device, root=root)
        Rs, Ts, exp_imgs, exp_masks = Rs[:6], Ts[:6], exp_imgs[:6], exp_masks[:6]        # six of the 144 views are enough here
        sdf = torch.jit.load(sdf_path)
        density_field = SDF(sdf=sdf)
        density_field.max_steps = 64
        learned_bsdf = ComposeSpatialVarying([
            *[NeuralBSDF() for _ in range(2)],
            Diffuse(preprocess=nn.Softplus()).random(),
            Conductor(activation=nn.Softplus(), device=device).random(),
        ])
        integrator = Direct()
        lights = PointLights(device=device, scale=5)
        occ_mlp = SkipConnMLP(in_size=5, out=1, device=device).to(device)
        opt = torch.optim.AdamW([
            {"params": density_field.parameters(), "lr": 8e-5},
            {"params": learned_bsdf.parameters(), "lr": 8e-5},
            {"params": lights.intensity_parameters(), "lr": 8e-5},
            {"params": occ_mlp.parameters(), "lr": 8e-5},
        ], lr=8e-5, weight_decay=0)

        seen_hits = []

        def extra_loss(mi, got, exp, mask):
            raw_n = getattr(mi, "raw_normals", None)
            loss = 0
            if raw_n is not None:
                seen_hits.append(raw_n.shape[0])
                loss = loss + eikonal_loss(raw_n)
            raw_w = getattr(mi, "normalized_weights", None)
            if raw_w is not None:
                loss = loss + 1e-2 * raw_w.std(dim=-1).mean()
            return loss

        def light_update(cam, light):
            light.location = cam.get_camera_center() * 1.05
        ops.profile_collect()
        losses = train_sample(density_field, bsdf=learned_bsdf, integrator=integrator, lights=lights, Rs=Rs, Ts=Ts,
                              exp_imgs=exp_imgs, exp_masks=exp_masks, opt=opt, size=SIZE, crop_size=SIZE, save_freq=7500,
                              valid_freq=2, max_valid_size=SIZE, iters=3, N=2, extra_loss=extra_loss,
                              uv_select=lambda _, crop_size: rand_uv(SIZE, SIZE, crop_size), light_update=light_update,
                              name_fn=lambda i: str(tmp_path / "outputs" / ("train_%06d.png" % i)),
                              valid_name_fn=lambda i: str(tmp_path / "outputs" / ("valid_%06d.png" % i)),
                              silent=True, really_silent=True, w_isect=occ_mlp)
        counts = {k2: c for k2, (_, c) in ops.profile_collect().items() if c}
        assert len(losses) >= 1 and all(np.isfinite(losses))
        assert counts.get("sdf_march_tc", 0) >= 3 and "sdf_march_f32" not in counts and not [k_ for k_ in counts if k_.endswith("_f32") and k_ != "sdf_value_grad_f32"], counts
        assert seen_hits and min(seen_hits) > 0, seen_hits
        # the learned occlusion is in the graph (it only scales the light where the shadow ray is blocked, scene.py:301-318: with
        # the light next to the camera that is almost nowhere, so its gradient is a tensor of zeros rather than a change)
        assert occ_mlp.out.weight.grad is not None and torch.isfinite(occ_mlp.out.weight.grad).all()
        assert counts.get("sdf_shadow_tc", 0) >= 1, counts
        torch.jit.save(density_field.sdf, sdf_path)
        torch.save(learned_bsdf, str(tmp_path / "models" / ("col_%s_bsdf.pt" % k)))
        ops.profile_collect()
        stats = test(density_field, integrator=integrator, bsdf=learned_bsdf, lights=lights, Rs=Rs, Ts=Ts, exp_imgs=exp_imgs,
                     size=SIZE, light_update=light_update,
                     name_fn=lambda i: str(tmp_path / "outputs" / ("col_final_%03d.png" % i)), w_isect=True)
        counts = {k2: c for k2, (_, c) in ops.profile_collect().items() if c}
        assert isinstance(stats, dict) and all(np.isfinite(np.asarray(v, np.float64)).all() for v in stats.values())
        assert counts.get("sdf_march_tc", 0) == len(Rs), counts                       # one row block per view
        reloaded = SDF(sdf=torch.jit.load(sdf_path))
        assert isinstance(reloaded._impl, SphereSDF)
        for (ka, a), (kb, b) in zip(sdf.state_dict().items(), reloaded.sdf.state_dict().items()):
            assert ka == kb and torch.equal(a.cpu(), b.cpu())
    finally:
        config.set_precision("f32"); config.set_train_precision("f32")


def test_dtu_script_flow(tmp_path):
    """scripts/dtu.py:88-200 on this package at toy size: `SDF(sdf=torch.jit.load(path, device))`, the BSDF and the lights read
    back with `torch.load` and `setattr(bsdf, "act", nn.Sigmoid())` on the children, the three-group AdamW, `train_dtu` with the
    eikonal extra loss and a validation render, the three save calls, `test_dtu`."""
    import torch
    import torch.nn as nn
    import scenes
    from neural_raytracing_b200 import config, ops
    from neural_raytracing_b200.pathtracer.bsdf import ComposeSpatialVarying, Diffuse, NeuralBSDF
    from neural_raytracing_b200.pathtracer.integrators import Direct
    from neural_raytracing_b200.pathtracer.lights import LightField
    from neural_raytracing_b200.pathtracer.shapes.sdfs import SDF, SphereSDF
    from neural_raytracing_b200.pathtracer.training_utils import test_dtu, train_dtu
    from neural_raytracing_b200.pathtracer.utils import eikonal_loss, rand_uv
    device, SIZE, dataset = "cuda", 16, 65
    random.seed(2); np.random.seed(2); torch.manual_seed(2)
    os.makedirs(tmp_path / "models"); os.makedirs(tmp_path / "outputs")
    m = lambda name: str(tmp_path / "models" / ("dtu_%d_%s.pt" % (dataset, name)))      # noqa: E731
    # what an earlier run left behind (dtu.py:159-170), written with the same calls
    start = SphereSDF(n=2 << 5, device=device)
    with torch.no_grad():
        start.radii.abs_().add_(0.05)
    torch.jit.save(torch.jit.script(start), m("sdf"))
    torch.save(ComposeSpatialVarying([NeuralBSDF() for _ in range(10)] +
                                     [Diffuse(preprocess=torch.sigmoid).random() for _ in range(6)]), m("bsdf"))
    torch.save(LightField(), m("lights"))
    poses, intrinsics = scenes.dtu_cameras(4, device=device)
    poses = [p for p in poses]
    g = torch.Generator().manual_seed(0)
    exp_imgs = [torch.rand(SIZE, SIZE, 3, generator=g).to(device) for _ in range(4)]
    exp_masks = [(torch.rand(SIZE, SIZE, generator=g) > 0.4).float().to(device) for _ in range(4)]
    try:
        config.set_precision("f16"); config.set_train_precision("f16")
        integrator = Direct()
        shape = torch.jit.load(m("sdf"), device)
        density_field = SDF(sdf=shape)
        density_field.max_steps = 64
        learned_bsdf = torch.load(m("bsdf"), weights_only=False)
        for bsdf in learned_bsdf.bsdfs:
            setattr(bsdf, "act", nn.Sigmoid())
        lights = torch.load(m("lights"), weights_only=False)
        torch.jit.save(density_field.sdf, str(tmp_path / "models" / "tmp.pt"))
        torch.save(learned_bsdf, str(tmp_path / "models" / "tmp.pt"))
        torch.save(lights, str(tmp_path / "models" / "tmp.pt"))
        opt = torch.optim.AdamW([
            {"params": density_field.parameters(), "lr": 8e-5},
            {"params": learned_bsdf.parameters(), "lr": 8e-5},
            {"params": lights.parameters(), "lr": 8e-5},
        ], lr=8e-5, weight_decay=0)

        def extra_loss(mi, got, exp, mask):
            raw_n = getattr(mi, "raw_normals", None)
            loss = 0
            if raw_n is not None:
                loss = loss + eikonal_loss(raw_n)
            return loss
        before = {k: v.detach().clone() for k, v in shape.state_dict().items()}
        ops.profile_collect()
        losses = train_dtu(density_field, bsdf=learned_bsdf, integrator=integrator, lights=lights, poses=poses,
                           intrinsics=intrinsics, exp_imgs=exp_imgs, exp_masks=exp_masks, opt=opt, size=SIZE, crop_size=SIZE,
                           save_freq=5000, valid_freq=2, max_valid_size=SIZE, N=2, iters=3, extra_loss=extra_loss,
                           uv_select=lambda _, crop_size: rand_uv(SIZE, SIZE, crop_size), silent=True,
                           name_fn=lambda i: str(tmp_path / "outputs" / ("train_dtu_%06d.png" % i)),
                           valid_name_fn=lambda i: str(tmp_path / "outputs" / ("valid_dtu_%06d.png" % i)))
        counts = {k: c for k, (_, c) in ops.profile_collect().items() if c}
        assert len(losses) == 3 and all(np.isfinite(losses))
        assert counts.get("sdf_march_tc", 0) >= 3 and counts.get("mlp_tc_wgrad", 0) >= 3 and "sdf_march_f32" not in counts and not [k_ for k_ in counts if k_.endswith("_f32") and k_ != "sdf_value_grad_f32"], counts
        assert any(not torch.equal(v, before[k]) for k, v in shape.state_dict().items())
        torch.jit.save(density_field.sdf, m("sdf"))
        torch.save(learned_bsdf, m("bsdf"))
        torch.save(lights, m("lights"))
        stats = test_dtu(density_field, integrator=integrator, bsdf=learned_bsdf, lights=lights, poses=poses[:2],
                         intrinsics=intrinsics[:2], exp_imgs=exp_imgs[:2], exp_masks=exp_masks[:2], size=SIZE,
                         name_fn=lambda i: str(tmp_path / "outputs" / ("dtu_test_%03d.png" % i)))
        assert isinstance(stats, dict) and all(np.isfinite(np.asarray(v, np.float64)).all() for v in stats.values())
        back = torch.load(m("bsdf"), weights_only=False)
        assert isinstance(back.bsdfs[0].act, nn.Sigmoid) and len(back.bsdfs) == 16
    finally:
        config.set_precision("f32"); config.set_train_precision("f32")


@pytest.mark.parametrize("envmap", [False, True])
def test_nerfle_script_flow(tmp_path, envmap):
    """scripts/nerfle.py:36-200 on this package at toy size: NeRFLE(envmap=...) under NeRFReproduce, the script's own training
    loop (pathtrace_sample on a crop + F.mse_loss + AdamW, the light following the camera), its validation render (one library
    call per frame, rays generated on the device), `torch.save(nerfle, ...)` of the whole module, `test(...)`."""
    import torch
    import torch.nn.functional as F
    import neural_raytracing_b200.pathtracer as pt
    from neural_raytracing_b200 import config, ops
    from neural_raytracing_b200.pathtracer.integrators import NeRFReproduce
    from neural_raytracing_b200.pathtracer.lights import PointLights
    from neural_raytracing_b200.pathtracer.shapes.nerf import NeRFLE
    from neural_raytracing_b200.pathtracer.training_utils import test
    from neural_raytracing_b200.pathtracer.utils import LossSampler, rand_uv
    from neural_raytracing_b200.renderer import OpenGLPerspectiveCameras, look_at_view_transform
    device, SIZE, DIST, crop_size, N = "cuda", 16, 1.0, 8, 2
    random.seed(3); np.random.seed(3); torch.manual_seed(3)
    os.makedirs(tmp_path / "models"); os.makedirs(tmp_path / "outputs")
    Rs, Ts, exp_imgs = [], [], []
    g = torch.Generator().manual_seed(1)
    for elev in torch.linspace(0, 45, 2, device=device):
        for azim in torch.linspace(-90, 90, 2, device=device):
            R, T = look_at_view_transform(dist=DIST, elev=elev, azim=azim, device=device)
            Rs.append(R); Ts.append(T)
            exp_imgs.append(torch.rand(SIZE, SIZE, 3, generator=g).to(device))
    try:
        config.set_precision("f16"); config.set_train_precision("f16")
        nerfle = NeRFLE(envmap=envmap, device=device)
        with torch.no_grad():
            nerfle.first.out.bias[0] = 0.8      # positive density at the start (a fresh model can be born with relu(sigma) = 0
            #                                     everywhere: a black image and no gradient, in the reference as well)
        integrator = NeRFReproduce()
        lights = PointLights(device=device, scale=10)
        opt = torch.optim.AdamW([{"params": nerfle.parameters(), "lr": 8e-5}], lr=8e-5, weight_decay=0)

        def light_update(cam, light):
            light.location = cam.get_camera_center() * 1.05
        before = nerfle.first.init.weight.detach().clone()
        selector = LossSampler(len(exp_imgs))
        losses = []
        ops.profile_collect()
        for i in range(3):                                                  # nerfle.py:88-118
            idxs = selector.sample(n=N)
            R = torch.cat([Rs[j] for j in idxs], dim=0)
            T = torch.cat([Ts[j] for j in idxs], dim=0)
            exp = torch.stack([exp_imgs[j] for j in idxs])
            cameras = OpenGLPerspectiveCameras(device=device, R=R, T=T)
            light_update(cameras, lights)
            opt.zero_grad()
            (u, v) = rand_uv(SIZE, SIZE, crop_size)
            got, mi = pt.pathtrace_sample(nerfle, size=SIZE, chunk_size=SIZE, bundle_size=1, crop_size=crop_size, bsdf=None,
                                          integrator=integrator, cameras=cameras, lights=lights, device=device, uv=(u, v),
                                          addition=lambda mi: mi, squeeze_first=False, silent=True)
            exp = exp[:, u:u + crop_size, v:v + crop_size]
            loss = F.mse_loss(got, exp)
            assert not loss.isnan()
            loss.backward()
            opt.step()
            losses.append(loss.detach().item())
        counts = {k: c for k, (_, c) in ops.profile_collect().items() if c}
        assert all(np.isfinite(losses)) and not torch.equal(nerfle.first.init.weight.detach(), before)
        assert counts.get("mlp_tc_train_fwd", 0) >= 3 and counts.get("mlp_tc_wgrad", 0) >= 3, counts     # tensor-core training kernels
        with torch.no_grad():                                               # nerfle.py:125-136
            cameras = OpenGLPerspectiveCameras(device=device, R=Rs[0], T=Ts[0])
            light_update(cameras, lights)
            ops.profile_collect()
            validate, _ = pt.pathtrace(nerfle, size=SIZE, chunk_size=min(SIZE, 8), bundle_size=1, bsdf=None,
                                       integrator=integrator, cameras=cameras, lights=lights, device=device, silent=True)
            counts = {k: c for k, (_, c) in ops.profile_collect().items() if c}
        assert tuple(validate.shape) == (SIZE, SIZE, 3) and torch.isfinite(validate).all()
        assert counts.get("camera_rays") == 1 and counts.get("mlp_tc_nerf_first") == 1, counts            # one call per frame
        path = str(tmp_path / "models" / ("nerfle_%s.pt" % ("envmap" if envmap else "ptl")))
        torch.save(nerfle, path)                                            # nerfle.py:160-161
        again = torch.load(path, weights_only=False)
        assert again.envmap == envmap
        with torch.no_grad():
            random.seed(7)                       # each forward draws its far plane from `random` (nerf.py:178): same draw for both
            v2, _ = pt.pathtrace(again, size=SIZE, chunk_size=SIZE, bundle_size=1, bsdf=None, integrator=integrator,
                                 cameras=cameras, lights=lights, device=device, silent=True, with_noise=False)
            random.seed(7)
            v1, _ = pt.pathtrace(nerfle, size=SIZE, chunk_size=SIZE, bundle_size=1, bsdf=None, integrator=integrator,
                                 cameras=cameras, lights=lights, device=device, silent=True, with_noise=False)
        assert torch.equal(v1, v2)
        stats = test(nerfle, integrator=integrator, bsdf=None, lights=lights, Rs=Rs, Ts=Ts, exp_imgs=exp_imgs, size=SIZE,
                     light_update=light_update, name_fn=lambda i: str(tmp_path / "outputs" / ("test_%03d.png" % i)))
        assert isinstance(stats, dict) and all(np.isfinite(np.asarray(v, np.float64)).all() for v in stats.values())
    finally:
        config.set_precision("f32"); config.set_train_precision("f32")
