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
