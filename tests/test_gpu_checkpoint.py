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


def test_script_archive_held_by_sdf_runs_on_the_fused_kernels(tmp_path):
    """The scripts' own lines (dtu.py:93-94): `shape = torch.jit.load(path, device); density_field = SDF(sdf=shape)`.  The march,
    the scan and the normals run on the library's kernels (not on the generic loop over a scripted callable), the result equals
    the load_sdf_archive route bit for bit, and gradients of a loss on the normals land on the ARCHIVE's parameters."""
    import torch
    from neural_raytracing_b200 import ops
    from neural_raytracing_b200.pathtracer import checkpoint
    from neural_raytracing_b200.pathtracer.shapes.sdfs import SDF, SphereSDF
    g = helpers.golden("sdf")
    path = str(tmp_path / "dtu_sdf.pt")
    checkpoint.save_sdf_archive(_sphere_sdf_from(helpers.golden_sdf_weights()), path)
    random.random = lambda: float(g["fixed_random"])
    shape = torch.jit.load(path, "cuda")
    field = SDF(sdf=shape, device="cuda")
    field.max_steps = 64
    assert field.sdf is shape and isinstance(field._impl, SphereSDF) and field._fused() is not None
    other = SDF(sdf=checkpoint.load_sdf_archive(path, device="cuda"), device="cuda")
    other.max_steps = 64
    rays = torch.from_numpy(g["rays"]).cuda().reshape(1, -1, 1, 1, 6)
    ops.profile_collect()
    with torch.no_grad():
        it, active = field.intersect(rays)
    counts = {k: c for k, (_, c) in ops.profile_collect().items() if c}
    assert counts.get("sdf_march_f32", 0) == 1 and counts.get("sdf_min_scan_f32", 0) >= 1, counts
    with torch.no_grad():
        it2, active2 = other.intersect(rays)
    assert torch.equal(active, active2) and torch.equal(it.t, it2.t) and torch.equal(it.throughput, it2.throughput)
    hit = active.reshape(-1).cpu().numpy()
    assert int((hit != g["hit"].astype(bool)).sum()) <= 2
    # a differentiable pass: eikonal-style loss on the raw normals -> gradients on the archive's own tensors
    it3, active3 = field.intersect(rays)
    assert active3.any()
    ((it3.raw_normals.norm(dim=-1) - 1) ** 2).mean().backward()
    grads = [p.grad for p in shape.parameters()]
    assert all(gr is not None for gr in grads) and sum(float(gr.abs().sum()) for gr in grads) > 0


def test_cpu_archive_loaded_without_a_device_is_moved_to_the_shape_device(tmp_path):
    """colocate.py:65 `sdf = torch.jit.load(path)` (no device): an archive written from CPU tensors lands on the host; SDF
    (device="cuda" by default) moves it next to the rays before adopting it."""
    import torch
    from neural_raytracing_b200.pathtracer import checkpoint
    from neural_raytracing_b200.pathtracer.shapes.sdfs import SDF, SphereSDF
    g = helpers.golden("sdf")
    path = str(tmp_path / "col_sdf.pt")
    checkpoint.save_sdf_archive(_sphere_sdf_from(helpers.golden_sdf_weights()), path)
    sdf = torch.jit.load(path)
    assert next(sdf.parameters()).device.type == "cpu"
    field = SDF(sdf=sdf)
    field.max_steps = 64
    assert next(sdf.parameters()).is_cuda and isinstance(field._impl, SphereSDF) and field._fused() is not None
    random.random = lambda: float(g["fixed_random"])
    rays = torch.from_numpy(g["rays"]).cuda().reshape(1, -1, 1, 1, 6)
    with torch.no_grad():
        it, active = field.intersect(rays)
    assert int((active.reshape(-1).cpu().numpy() != g["hit"].astype(bool)).sum()) <= 2
