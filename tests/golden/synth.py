"""Deterministic synthetic weights / rays shared by tests/golden/make_golden.py (which loads them
into the UNMODIFIED reference modules) and by the tests (which load them into the oracle and
into the CUDA path).  numpy's legacy RandomState stream is frozen across numpy versions, so
fixtures only need to store inputs and reference outputs, never weights."""
import numpy as np


def mlp_weights(seed, in_size, out, num_layers, hidden, freqs, sigma, latent=0, skip=3,
                w_scale=1.0, zero=False):
    """torch-layout weights of a SkipConnMLP (neural_blocks.py:36-55): list order
    [init, layers[0..L-1], out]; W [N,K] ~ U(-1/sqrt(K), 1/sqrt(K)) * w_scale."""
    rs = np.random.RandomState(seed)
    basis = (sigma * rs.standard_normal((freqs, in_size))).astype(np.float32).T.copy()  # [in, freqs]
    dim_p = in_size + 2 * freqs + latent
    shapes = [(hidden, dim_p)]
    for i in range(num_layers):
        sk = (i % skip) == 0 and i != num_layers - 1
        shapes.append((hidden, hidden + (dim_p if sk else 0)))
    shapes.append((out, hidden))
    W, b = [], []
    for (n, k) in shapes:
        bound = w_scale / np.sqrt(k)
        W.append(rs.uniform(-bound, bound, size=(n, k)).astype(np.float32) * (0 if zero else 1))
        b.append(rs.uniform(-bound, bound, size=(n,)).astype(np.float32) * (0 if zero else 1))
    return dict(in_size=in_size, out=out, num_layers=num_layers, hidden=hidden, freqs=freqs,
                latent=latent, skip=skip, basis=basis, W=W, b=b)


def sdf_weights(seed, n=64, hidden=128, num_layers=8, freqs=32, shift_std=0.02):
    """SphereSDF parameters (sdfs.py:17-31): centers 0.3U-0.15, radii 0.2U-0.1, tfs small
    (the constructor uses zeros; a small perturbation exercises the affine warp), and the
    residual MLP `shift` ~ N(0, shift_std) instead of the constructor's zero-init, which would
    hide MLP errors."""
    rs = np.random.RandomState(seed)
    centers = (0.3 * rs.uniform(size=(n, 3)) - 0.15).astype(np.float32)
    radii = (0.2 * rs.uniform(size=(n,)) - 0.1).astype(np.float32)
    tfs = (0.05 * rs.standard_normal((n, 3, 3))).astype(np.float32)
    m = mlp_weights(seed + 1, 3, 1, num_layers, hidden, freqs, 32.0)
    for i in range(len(m["W"])):
        m["W"][i] = (shift_std * rs.standard_normal(m["W"][i].shape)).astype(np.float32)
        m["b"][i] = (shift_std * rs.standard_normal(m["b"][i].shape)).astype(np.float32)
    return dict(n=n, centers=centers, radii=radii, tfs=tfs, shift=m)


def camera_rays(seed, n_rays, dist=1.0, jitter=0.35):
    """Rays from points on a sphere of radius `dist` looking roughly at the origin."""
    rs = np.random.RandomState(seed)
    o = rs.standard_normal((n_rays, 3))
    o = dist * o / np.linalg.norm(o, axis=-1, keepdims=True)
    d = -o + jitter * rs.standard_normal((n_rays, 3))
    d = d / np.linalg.norm(d, axis=-1, keepdims=True)
    return np.concatenate([o, d], -1).astype(np.float32)
