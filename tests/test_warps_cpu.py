"""CPU test (-m "not gpu"): sample warps vs fixtures from the unmodified reference (warps.py:10-52)."""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))

from neural_raytracing_b200.pathtracer.warps import (square_to_cos_hemisphere, square_to_cos_hemisphere_pdf,  # noqa: E402
                                                     square_to_uniform_disk_concentric)

G = np.load(os.path.join(HERE, "golden", "path.npz"))


def test_warps_match_reference():
    u = torch.from_numpy(G["warp_u"])
    assert np.abs(square_to_uniform_disk_concentric(u).numpy() - G["warp_disk"]).max() < 1e-6
    h = square_to_cos_hemisphere(u)
    assert np.abs(h.numpy() - G["warp_hemi"]).max() < 1e-6
    assert (h.norm(dim=-1) - 1).abs().max().item() < 1e-3 and (h[..., 2] > 0).all()
    assert torch.allclose(square_to_cos_hemisphere_pdf(h), h[..., 2] / np.pi)
    assert square_to_uniform_disk_concentric(torch.full((1, 2), 0.5)).abs().max().item() == 0.0
