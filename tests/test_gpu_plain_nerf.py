"""GPU test (-m gpu) of PlainNeRF (shapes/nerf.py:9-74: per-image latent, (elev, azim) view encoding, tanh colour)
against the unmodified reference's output (tests/golden/plain_nerf.npz; density noise switched off on both sides)."""
import random

import numpy as np
import pytest

import helpers
import synth

pytestmark = pytest.mark.gpu


def test_plain_nerf_matches_reference():
    import torch
    from neural_raytracing_b200.pathtracer.shapes.nerf import PlainNeRF
    g = helpers.golden("plain_nerf")
    random.random = lambda: float(g["fixed_random"])
    n = PlainNeRF(device="cuda")
    synth.fill_module(n, 81)
    with torch.no_grad():
        n.first.out.bias[0] = 0.8
    n.assign_latent(torch.from_numpy(g["latent"]).cuda())
    rays = torch.from_numpy(g["rays"]).cuda()
    real = torch.randn_like
    torch.randn_like = lambda t, **k: torch.zeros_like(t)
    try:
        with torch.no_grad():
            rgb = n(rays, None)
        # differentiable path: same values, gradients reach both MLPs
        rgb2 = n(rays, None)
    finally:
        torch.randn_like = real
    assert tuple(rgb.shape) == g["rgb"].shape
    assert np.abs(rgb.cpu().numpy() - g["rgb"]).max() < 2e-5
    assert np.abs(rgb2.detach().cpu().numpy() - g["rgb"]).max() < 2e-5
    assert g["rgb"].std() > 1e-3                       # the fixture is not the degenerate all-0.5 image
    rgb2.square().mean().backward()
    assert n.first.init.weight.grad.abs().sum() > 0 and n.second.out.weight.grad.abs().sum() > 0


def test_plain_nerf_on_the_tensor_cores_matches_reference():
    """set_precision("f16"): both 5x32 latent networks run on k_mlp_tc (no fp32 fallback: the launch counter of the
    tensor-core kernel moves), the image stays within north_star's 1e-3 of the reference's."""
    import torch
    from neural_raytracing_b200 import config, ops
    from neural_raytracing_b200.pathtracer.shapes.nerf import PlainNeRF
    g = helpers.golden("plain_nerf")
    real_random, real_randn = random.random, torch.randn_like
    random.random = lambda: float(g["fixed_random"])
    n = PlainNeRF(device="cuda")
    synth.fill_module(n, 81)
    with torch.no_grad():
        n.first.out.bias[0] = 0.8
    n.assign_latent(torch.from_numpy(g["latent"]).cuda())
    rays = torch.from_numpy(g["rays"]).cuda()
    torch.randn_like = lambda t, **k: torch.zeros_like(t)
    try:
        config.set_precision("f16")
        assert n.first.precision() == "f16" and n.second.precision() == "f16"
        ops.profile_collect()
        with torch.no_grad():
            rgb = n(rays, None)
        counts = {k: c for k, (_, c) in ops.profile_collect().items() if c}
    finally:
        config.set_precision("f32")
        torch.randn_like = real_randn
        random.random = real_random
    assert counts.get("mlp_tc_generic", 0) >= 2 and "mlp_fwd_f32" not in counts, counts
    assert np.abs(rgb.cpu().numpy() - g["rgb"]).max() < 1e-3
