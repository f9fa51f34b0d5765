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
        assert "sdf_march_f32" not in counts
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
        assert counts.get("sdf_march_tc", 0) >= 3 and "sdf_march_f32" not in counts, counts
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
