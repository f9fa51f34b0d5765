"""The vis scripts' helpers on top of the path (utils.sphere_examples / depth_image, shapes.Sphere, renderer.PointLights)
against outputs of the UNMODIFIED reference (tests/golden/sphere_examples.npz; pytorch3d/pathtracer/utils.py:409-445,
shapes/shapes.py:9-97, renderer/lighting.py:220-305).  The analytic sphere and depth_image are elementwise torch and are
checked on the CPU; the per-basis renders run the BSDF networks and are checked on the CPU (torch expressions) and on
cuda:0 (library kernels, fp32 and f16)."""
import os
import sys

import numpy as np
import pytest
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
sys.path.insert(0, HERE)
sys.path.insert(0, os.path.join(HERE, "golden"))

import scenes  # noqa: E402

G = np.load(os.path.join(HERE, "golden", "sphere_examples.npz"))


def test_analytic_sphere_matches_reference():
    from neural_raytracing_b200.pathtracer.shapes import Sphere
    ball = Sphere([0.1, -0.2, 0.05], 0.8, device="cpu")
    rays = torch.from_numpy(G["rays"])
    si, hit = ball.intersect(rays)
    assert np.array_equal(hit.numpy(), G["hit"]) and 0 < hit.sum() < len(hit)
    h = G["hit"]
    assert np.array_equal(np.isinf(si.t.numpy()), np.isinf(G["t"]))
    assert np.abs(si.t.numpy()[h] - G["t"][h]).max() < 1e-6
    for key, got in (("p", si.p), ("n", si.n), ("wi", si.wi)):
        assert np.abs(got.numpy()[h] - G[key][h]).max() < 2e-6, key
    assert np.array_equal(ball.intersect_test(rays).numpy(), G["hit"])
    lo, hi, m = ball.intersect_limits(rays)
    assert np.array_equal(m.numpy(), G["hit"])
    assert np.abs(lo.numpy()[h] - G["lo"][h]).max() < 1e-6
    both = h & np.isfinite(G["hi"])
    assert np.abs(hi.numpy()[both] - G["hi"][both]).max() < 1e-6


def test_depth_image_matches_reference():
    from neural_raytracing_b200.pathtracer.utils import depth_image
    out = depth_image(torch.from_numpy(G["depth_in"]))
    assert np.abs(out.numpy() - G["depth_out"]).max() < 1e-7


def _bases(device):
    import neural_raytracing_b200.pathtracer as P
    from neural_raytracing_b200.pathtracer.utils import sphere_examples
    _shape, _sphere, bsdf, _lights, _integ, _w = scenes.build_pipeline(P, "dtu", device=device)
    with torch.no_grad():
        imgs = sphere_examples(bsdf, device=device, size=24, chunk_size=12, scale=100)
    return np.stack([i.cpu().numpy() for i in imgs])


def _check(imgs, tol):
    ref = G["bases"]
    assert imgs.shape == ref.shape
    # the default 1e-3-pixel jitter of pathtrace is on (torch's generator, different per device): silhouette pixels may
    # flip, everything else moves by ~1e-5
    err = np.abs(imgs - ref).max(axis=-1)
    assert (err < tol * ref.max()).mean() > 0.99, (err.max(), (err < tol * ref.max()).mean())
    assert np.abs(imgs.mean(axis=(1, 2, 3)) - ref.mean(axis=(1, 2, 3))).max() < 5e-3
    assert ref.std() > 0.1                                  # lit sphere on a white background, not a constant image


def test_sphere_examples_cpu_matches_reference():
    _check(_bases("cpu"), 1e-3)


@pytest.mark.gpu
@pytest.mark.parametrize("prec", ["f32", "f16"])
def test_sphere_examples_gpu_matches_reference(prec):
    from neural_raytracing_b200 import config
    try:
        config.set_precision(prec)
        _check(_bases("cuda"), 1e-3 if prec == "f32" else 3e-3)
    finally:
        config.set_precision("f32")
