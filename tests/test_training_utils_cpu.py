"""CPU tests (-m "not gpu") of the host-side pieces around the training loops: the in-repo SSIM restatement
(pathtracer/ssim.py; the reference's `pytorch_msssim` dependency is absent and unpinned, so these are property and
hand-computed checks, not parity), LossSampler (utils.py:134-147), rand_uv_mask (utils.py:378-383) and masked_loss
(utils.py:307-359)."""
import os
import random
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))

from neural_raytracing_b200.pathtracer.ssim import ssim  # noqa: E402
from neural_raytracing_b200.pathtracer.utils import LossSampler, masked_loss, rand_uv_mask  # noqa: E402


def test_ssim_identity_symmetry_and_range():
    torch.manual_seed(0)
    x = torch.rand(2, 3, 32, 40)
    y = (x + 0.1 * torch.randn_like(x)).clamp(0, 1)
    assert abs(ssim(x, x, data_range=1).item() - 1.0) < 1e-6
    a, b = ssim(x, y, data_range=1).item(), ssim(y, x, data_range=1).item()
    assert abs(a - b) < 1e-6 and 0.0 < a < 1.0
    z = torch.rand(2, 3, 32, 40)
    assert ssim(x, z, data_range=1).item() < a            # unrelated noise is less similar than a noisy copy
    per_image = ssim(x, y, data_range=1, size_average=False)
    assert tuple(per_image.shape) == (2,) and abs(per_image.mean().item() - a) < 1e-6


def test_ssim_constant_images_hand_computed():
    # constant images: all variances vanish, the map is (2ab + C1)/(a^2 + b^2 + C1) * (C2/C2) everywhere
    a, b = 0.2, 0.6
    x = torch.full((1, 1, 16, 16), a)
    y = torch.full((1, 1, 16, 16), b)
    c1 = 0.01 ** 2
    expect = (2 * a * b + c1) / (a * a + b * b + c1)
    assert abs(ssim(x, y, data_range=1).item() - expect) < 1e-4      # fp32 cancellation in the variance terms


def test_ssim_valid_window_and_small_images():
    torch.manual_seed(1)
    x, y = torch.rand(1, 3, 11, 11), torch.rand(1, 3, 11, 11)
    # 11x11 image, 11-tap valid window: a single output position per channel = Gaussian-weighted global statistics
    g = torch.exp(-((torch.arange(11.0) - 5) ** 2) / (2 * 1.5 ** 2)); g = g / g.sum()
    w = g[:, None] * g[None, :]
    vals = []
    for c in range(3):
        mx, my = (w * x[0, c]).sum(), (w * y[0, c]).sum()
        sx, sy = (w * x[0, c] ** 2).sum() - mx ** 2, (w * y[0, c] ** 2).sum() - my ** 2
        sxy = (w * x[0, c] * y[0, c]).sum() - mx * my
        c1, c2 = 0.01 ** 2, 0.03 ** 2
        vals.append(((2 * mx * my + c1) / (mx ** 2 + my ** 2 + c1)) * ((2 * sxy + c2) / (sx + sy + c2)))
    assert abs(ssim(x, y, data_range=1).item() - torch.stack(vals).mean().item()) < 1e-5
    # smaller than the window: no blur at all (what the training crops of 4x4 .. 8x8 pixels get)
    s = ssim(x[..., :4, :4], x[..., :4, :4], data_range=1).item()
    assert abs(s - 1.0) < 1e-6
    # gradients flow
    xg = x.clone().requires_grad_()
    (-ssim(xg, y, data_range=1).log()).backward()
    assert torch.isfinite(xg.grad).all() and xg.grad.abs().sum() > 0


def test_loss_sampler_prefers_high_loss_and_ages():
    np.random.seed(0)
    s = LossSampler(4)
    assert sorted(s.sample(n=4).tolist()) == [0, 1, 2, 3]      # without replacement: a permutation
    s.update_idxs([0, 1, 2], 0.0)                               # seen views drop to 1, the unseen one keeps 1e5
    picks = [int(s.sample(n=1)[0]) for _ in range(50)]
    assert picks.count(3) == 50
    before = s.losses.copy()
    s.update(3, 2.0)
    assert s.losses[3] == 3.0 and np.all(s.losses[:3] > before[:3])   # everything else ages by likelihood_inc


def test_rand_uv_mask_stays_inside_and_on_the_mask():
    random.seed(0)
    mask = torch.zeros(32, 32)
    mask[10:20, 12:18] = 1
    for _ in range(20):
        u, v = rand_uv_mask(mask, 8)
        u, v = int(u), int(v)
        # indices are relative to the margin-trimmed view (utils.py:380): offset by half a crop like the reference
        assert mask[4 + u, 4 + v] == 1 and 0 <= u <= 32 - 8 and 0 <= v <= 32 - 8


def test_masked_loss_terms():
    torch.manual_seed(0)
    got = torch.rand(2, 8, 8, 3, requires_grad=True)
    exp = torch.rand(2, 8, 8, 3)
    thr = torch.randn(2, 8, 8) * 3
    mask = (torch.rand(2, 8, 8) > 0.4).float()
    full = masked_loss(got, exp, thr, mask, mask_weight=10)
    no_ssim = masked_loss(got, exp, thr, mask, mask_weight=10, ssim_fn=None)
    active = ((thr > 0) & (mask == 1))
    ga, ea = got * active[..., None], exp * active[..., None]
    l2 = ((ga - ea) ** 2).mean()
    color = l2 + l2.sqrt() + (ga - ea).abs().mean()
    bce = torch.nn.functional.binary_cross_entropy_with_logits(thr[~active].reshape(-1, 1), mask[~active].reshape(-1, 1))
    assert abs(no_ssim.item() - (10 * bce + 10 * color).item()) < 1e-5
    s = ssim(ga.permute(0, 3, 1, 2), ea.permute(0, 3, 1, 2), data_range=1)
    assert abs(full.item() - (no_ssim - 10 * s.log()).item()) < 1e-4
    full.backward()
    assert torch.isfinite(got.grad).all()
