"""GPU leg of the checkpoint interchange (pathtracer/checkpoint.py): a TorchScript SDF archive in the reference's format, loaded
with load_sdf_archive, drives the FUSED kernels through the drop-in classes (SDF(sdf=...).intersect, dtu.py:93-94) and reproduces
the unmodified reference's hit mask and depths of tests/golden/sdf.npz."""
import random

import numpy as np
import pytest

import helpers
from test_checkpoint_cpu import _sphere_sdf_from

pytestmark = pytest.mark.gpu


def test_loaded_archive_runs_on_the_fused_march(tmp_path):
    import torch
    from neural_raytracing_b200.pathtracer import checkpoint
    from neural_raytracing_b200.pathtracer.shapes.sdfs import SDF, SphereSDF
    g = helpers.golden("sdf")
    path = str(tmp_path / "dtu_sdf.pt")
    checkpoint.save_sdf_archive(_sphere_sdf_from(helpers.golden_sdf_weights()), path)
    sphere = checkpoint.load_sdf_archive(path, device="cuda")
    assert isinstance(sphere, SphereSDF) and sphere.centers.is_cuda and sphere.shift.basis_p.is_cuda
    random.random = lambda: float(g["fixed_random"])
    shape = SDF(sdf=sphere, device="cuda")
    shape.max_steps = 64
    assert shape._fused() is not None                     # the packed parameters of the fused kernels, not the generic march
    rays = torch.from_numpy(g["rays"]).cuda().reshape(1, -1, 1, 1, 6)
    with torch.no_grad():
        it, active = shape.intersect(rays)
    hit = active.reshape(-1).cpu().numpy()
    assert int((hit != g["hit"].astype(bool)).sum()) <= 2            # same gate as the kernel-level test vs the torch reference
    same = hit == g["hit"].astype(bool)
    depth = it.t.reshape(-1).cpu().numpy()
    assert np.abs(depth[same] - g["depth"][same]).max() < 5e-4
