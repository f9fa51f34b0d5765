"""GPU tests (-m gpu) of the training / evaluation loops (pathtracer/training_utils.py)
against the UNMODIFIED reference's train_nerf (training_utils.py:211-300) run on CPU for three iterations
(tests/golden/make_golden.py::gen_train_loop -> train_loop.npz).  The reference and the mirror draw the same views and
crops (numpy / python RNG, same seeds); the sub-pixel camera jitter comes from torch's device RNG and differs, and the
silhouette logits are 1000 x the SDF, so losses are compared to 2e-3 relative."""
import os
import random
import sys

import numpy as np
import pytest

import helpers
import synth

pytestmark = pytest.mark.gpu
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden"))


def _case(train_nerf, P, device):
    import torch
    import scenes
    shape, sphere, bsdf, lights, _integ, _w = scenes.build_pipeline(P, "dtu", device=device)
    size, crop = 16, 4
    c2w, focal = synth.nerf_cameras(3, size, device=device)
    gx, gy = np.meshgrid(np.linspace(0, 1, size), np.linspace(0, 1, size), indexing="ij")
    imgs, masks = [], []
    for i in range(3):
        img = np.stack([0.3 + 0.4 * gx, 0.5 + 0.0 * gy, 0.6 - 0.3 * gy], axis=-1) * (0.8 + 0.1 * i)
        m = ((gx - 0.5) ** 2 + (gy - 0.5) ** 2 < 0.2).astype(np.float32)
        imgs.append(torch.tensor(img, dtype=torch.float, device=device))
        masks.append(torch.tensor(m, dtype=torch.float, device=device))
    params = list(sphere.parameters()) + list(bsdf.parameters()) + list(lights.parameters())
    opt = torch.optim.AdamW(params, lr=8e-5, weight_decay=0)
    random.seed(5); np.random.seed(5); torch.manual_seed(5)
    seen = []
    losses = train_nerf(shape, bsdf, P.integrators.Direct(), lights, [c for c in c2w], focal, imgs, masks, opt, size, crop,
                        N=2, iters=3, num_ckpts=1, save_freq=10 ** 6, valid_freq=10 ** 6, silent=True,
                        step_hook=lambda i, l: seen.append((i, l)))
    return losses, seen, sphere, bsdf, (shape, lights, c2w, focal, imgs, masks, size)


def test_train_nerf_matches_reference_loop(capsys):
    import torch
    import neural_raytracing_b200.pathtracer as P
    from neural_raytracing_b200 import config
    from neural_raytracing_b200.pathtracer import training_utils as TU
    g = helpers.golden("train_loop")
    random.random = lambda: float(g["fixed_random"])
    config.set_precision("f32")
    losses, seen, sphere, bsdf, _ = _case(TU.train_nerf, P, "cuda")
    ref = g["losses"]
    assert len(losses) == len(ref) == 3 and [i for i, _ in seen] == [0, 1, 2]
    for a, b in zip(losses, ref):
        assert abs(a - b) <= 2e-3 * abs(b), (losses, ref.tolist())
    # the optimizer moved the same parameters the same way
    w = sphere.shift.out.weight.detach().cpu().numpy()
    assert np.abs(w - g["sdf_out_w_after"]).max() < 5e-5
    assert np.abs(bsdf.sp_var_fn.out.bias.detach().cpu().numpy() - g["spvar_out_b_after"]).max() < 5e-5
    out = capsys.readouterr().out
    assert "000000:" in out and "000002:" in out            # silent=True prints one line per iteration


def test_evaluation_loops_report_metrics(tmp_path):
    import torch
    import neural_raytracing_b200.pathtracer as P
    from neural_raytracing_b200 import config
    from neural_raytracing_b200.pathtracer import training_utils as TU
    from neural_raytracing_b200.renderer import look_at_view_transform
    random.random = lambda: 0.37
    config.set_precision("f32")
    import scenes
    shape, sphere, bsdf, lights, integ, _w = scenes.build_pipeline(P, "dtu", device="cuda")
    size = 16
    c2w, focal = synth.nerf_cameras(2, size, device="cuda")
    exp = [torch.full((size, size, 4), 0.5, device="cuda") for _ in range(2)]
    names = []
    stats = TU.test_nerf(shape, integ, bsdf, lights, [c for c in c2w], focal, exp, size,
                         name_fn=lambda i: names.append(i) or str(tmp_path / ("t%d.png" % i)))
    assert set(stats) == {"l1", "l2", "psnr", "ssim"} and all(np.isfinite(v) for v in stats.values())
    assert abs(stats["psnr"] + 10 * np.log10(stats["l2"])) < 0.5 and names == [0, 1]
    assert (tmp_path / "t0.png").exists()
    # colocate-style test(): look-at cameras, the light follows the camera
    shape, sphere, bsdf, lights, integ, w_isect = scenes.build_pipeline(P, "colocate", device="cuda")
    Rs, Ts = [], []
    for az in (-40.0, 10.0, 60.0):
        R, T = look_at_view_transform(dist=1.0, elev=20.0, azim=az, device="cuda")
        Rs.append(R); Ts.append(T)

    def follow(cameras, lts):
        lts.location = cameras.get_camera_center() * 1.05

    exp3 = [torch.full((size, size, 3), 0.5, device="cuda") for _ in range(3)]
    stats = TU.test(shape, integ, bsdf, lights, Rs, Ts, exp3, size, max_chunk_size=16, light_update=follow,
                    name_fn=lambda i: str(tmp_path / ("c%d.png" % i)), w_isect=w_isect)
    assert all(np.isfinite(v) for v in stats.values()) and 0 < stats["l1"] < 1


def test_train_sample_colocate_style_steps():
    """train_sample with the camera factory the reference forgot to define: runs, updates the weights, skips nothing."""
    import torch
    import neural_raytracing_b200.pathtracer as P
    from neural_raytracing_b200 import config
    from neural_raytracing_b200.pathtracer import training_utils as TU
    from neural_raytracing_b200.renderer import look_at_view_transform
    import scenes
    random.random = lambda: 0.37
    config.set_precision("f32")
    shape, sphere, bsdf, lights, integ, w_isect = scenes.build_pipeline(P, "colocate", device="cuda")
    size, crop = 16, 4
    Rs, Ts = [], []
    for az in (-40.0, 10.0, 60.0):
        R, T = look_at_view_transform(dist=1.0, elev=20.0, azim=az, device="cuda")
        Rs.append(R); Ts.append(T)
    imgs = [torch.full((size, size, 3), 0.4, device="cuda") for _ in range(3)]
    masks = [torch.ones(size, size, device="cuda") for _ in range(3)]
    params = list(sphere.parameters()) + list(bsdf.parameters()) + list(w_isect.parameters())
    opt = torch.optim.AdamW(params, lr=1e-4, weight_decay=0)
    before = sphere.shift.out.weight.detach().clone()
    random.seed(1); np.random.seed(1); torch.manual_seed(1)

    def follow(cameras, lts):
        lts.location = cameras.get_camera_center() * 1.05

    losses = TU.train_sample(shape, bsdf, P.integrators.NeRFIntegrator(integ), lights, Rs, Ts, imgs, masks, opt, size, crop,
                             N=2, iters=2, num_ckpts=1, save_freq=10 ** 6, valid_freq=10 ** 6, silent=True,
                             light_update=follow, w_isect=w_isect)
    assert len(losses) == 2 and all(np.isfinite(l) for l in losses)
    assert (sphere.shift.out.weight.detach() - before).abs().max().item() > 0
