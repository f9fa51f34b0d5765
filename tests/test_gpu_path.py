"""GPU test (-m gpu) of the multi-bounce Path integrator and `bsdf.sample` against the
unmodified reference's renders (tests/golden/path.npz: integrators.py:274-354 with one and two bounces, deterministic
sampler, torch.multinomial replaced by argmax on both sides)."""
import random

import numpy as np
import pytest

import helpers
import synth

pytestmark = pytest.mark.gpu


def _render(depth, prec="f32"):
    import torch
    import neural_raytracing_b200.pathtracer as P
    from neural_raytracing_b200 import config
    from neural_raytracing_b200.pathtracer.cameras import NeRFCamera
    import scenes
    config.set_precision(prec)
    shape, sphere, bsdf, lights, _integ, _w = scenes.build_pipeline(P, "dtu", device="cuda")
    size = 16
    c2w, focal = synth.nerf_cameras(1, size, device="cuda")
    cam = NeRFCamera(cam_to_world=c2w, focal=focal, device="cuda")
    real = torch.multinomial
    torch.multinomial = lambda k, num_samples=1, **kw: k.argmax(dim=-1, keepdim=True)
    try:
        with torch.no_grad():
            img, _ = P.pathtrace(shape, size=size, chunk_size=size, bundle_size=1, bsdf=bsdf, integrator=P.Path(max_depth=depth),
                                 lights=lights, cameras=cam, device="cuda", silent=True, background=0, with_noise=False,
                                 sampler=synth.GoldenRatioSampler())
    finally:
        torch.multinomial = real
        config.set_precision("f32")
    return img.cpu().numpy()


@pytest.mark.parametrize("depth", [1, 2])
def test_path_integrator_matches_reference(depth):
    g = helpers.golden("path")
    random.random = lambda: float(g["fixed_random"])
    ref = g["img_depth%d" % depth]
    img = _render(depth)
    assert img.shape == ref.shape
    err = np.abs(img - ref).max(axis=-1)
    assert (err < 1e-3).mean() >= 0.97, (err.max(), (err < 1e-3).mean())
    assert helpers.psnr(img, ref) > 50
    if depth == 2:
        # the second bounce is really there: the pixels it lights in the reference are lit here as well
        extra = np.abs(g["img_depth2"] - g["img_depth1"]).max(axis=-1) > 1e-4
        assert extra.sum() >= 1
        mine = np.abs(img - _render(1)).max(axis=-1) > 1e-4
        assert (mine & extra).sum() >= 1


def test_path_integrator_tensor_core_precision():
    g = helpers.golden("path")
    random.random = lambda: float(g["fixed_random"])
    img = _render(2, prec="f16")
    assert helpers.psnr(img, g["img_depth2"]) > 45 and np.isfinite(img).all()


def test_conductor_sample_is_the_local_mirror():
    import torch
    import neural_raytracing_b200.pathtracer as P
    from neural_raytracing_b200.pathtracer.interaction import MixedInteraction
    c = P.bsdf.Conductor(device="cuda")
    p = torch.zeros(5, 3, device="cuda")
    it = MixedInteraction(p=p, t=torch.ones(5, device="cuda"), obj=None, throughput=0)
    it.wi = torch.nn.functional.normalize(torch.tensor([[0.3, -0.2, 0.9], [0.1, 0.1, -0.5], [0.0, 0.0, 1.0], [0.6, 0.6, 0.2],
                                                        [-0.4, 0.2, 0.7]], device="cuda"), dim=-1)
    bs, spec = c.sample(it, synth.GoldenRatioSampler(), active=torch.ones(5, dtype=torch.bool, device="cuda"))
    assert torch.allclose(bs.wo, it.wi * torch.tensor([-1.0, -1.0, 1.0], device="cuda"))
    assert (spec[1] == 0).all() and (spec[0] > 0).all()          # back-facing point gets no energy
