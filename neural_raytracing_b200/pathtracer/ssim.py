"""In-repo SSIM for `masked_loss` and the test loops.

The reference takes `ssim` from the third-party package `pytorch_msssim` (utils.py:7, 328-347; training_utils.py:18,
342, 483, 532), which is neither vendored nor version-pinned by the reference and is not installed here.  This is a
restatement of that package's PUBLISHED algorithm (Wang et al. 2004 as implemented by pytorch_msssim.ssim): an 11-tap
Gaussian window (sigma 1.5) applied separably with VALID padding per channel, K = (0.01, 0.03), the mean of the SSIM
map per image and channel, then the mean over both when `size_average`.  PARITY UNPINNED: no reference test, fixture or
installed copy of the dependency exists to check it against; tests/test_training_utils_cpu.py checks its defining
properties (ssim(x, x) = 1, symmetry, range, a hand-computed constant-image case) instead.
Plain torch (two grouped convolutions per moment): it is a loss on <= 4 x 128 x 128 crops, off the per-ray path."""
import torch
import torch.nn.functional as F


def _gauss_window(size: int, sigma: float, device, dtype):
    x = torch.arange(size, device=device, dtype=dtype) - size // 2
    g = torch.exp(-(x * x) / (2.0 * sigma * sigma))
    return g / g.sum()


def _blur(x, win):
    """Separable valid-mode Gaussian per channel on [N,C,H,W]; a spatial dim shorter than the window is left alone
    (what pytorch_msssim does, with a warning)."""
    c = x.shape[1]
    k = win.numel()
    if x.shape[2] >= k:
        x = F.conv2d(x, win.reshape(1, 1, k, 1).expand(c, 1, k, 1), groups=c)
    if x.shape[3] >= k:
        x = F.conv2d(x, win.reshape(1, 1, 1, k).expand(c, 1, 1, k), groups=c)
    return x


def ssim(X, Y, data_range=255, size_average=True, win_size=11, win_sigma=1.5, K=(0.01, 0.03), nonnegative_ssim=False):
    """Structural similarity of two image batches [N,C,H,W] (same signature subset as pytorch_msssim.ssim)."""
    if X.shape != Y.shape:
        raise ValueError("ssim: inputs must have the same shape, got %s and %s" % (tuple(X.shape), tuple(Y.shape)))
    if X.dim() != 4:
        raise ValueError("ssim: expected [N,C,H,W] inputs, got %d dims" % X.dim())
    if win_size % 2 != 1:
        raise ValueError("ssim: window size must be odd")
    win = _gauss_window(win_size, win_sigma, X.device, X.dtype)
    c1, c2 = (K[0] * data_range) ** 2, (K[1] * data_range) ** 2
    mu1, mu2 = _blur(X, win), _blur(Y, win)
    mu1_sq, mu2_sq, mu12 = mu1 * mu1, mu2 * mu2, mu1 * mu2
    s1 = _blur(X * X, win) - mu1_sq
    s2 = _blur(Y * Y, win) - mu2_sq
    s12 = _blur(X * Y, win) - mu12
    cs = (2 * s12 + c2) / (s1 + s2 + c2)
    smap = ((2 * mu12 + c1) / (mu1_sq + mu2_sq + c1)) * cs
    per_channel = smap.flatten(2).mean(-1)
    if nonnegative_ssim:
        per_channel = torch.relu(per_channel)
    return per_channel.mean() if size_average else per_channel.mean(1)
