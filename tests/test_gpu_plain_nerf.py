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
