"""GPU tests (-m gpu) of the error conventions of the boundary (TORCH_CHECK-style loud failures, no
silent fallback) and of argument edge cases the reference's callers can produce."""
import numpy as np
import pytest

import helpers
import synth

pytestmark = pytest.mark.gpu


def _t(a):
    import torch
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


def test_wrong_dtype_and_shape_raise():
    import torch
    from neural_raytracing_b200 import ops
    w1, _ = helpers.nerfle_weights(False)
    m1 = helpers.cuda_mlp(w1)
    x = torch.zeros(10, 3, device="cuda", dtype=torch.float64)
    with pytest.raises(ops.NrtError):
        ops.mlp_forward(m1, x)
    with pytest.raises(Exception):
        ops.mlp_forward(m1, torch.zeros(10, 5, device="cuda"))            # in_size is 3
    with pytest.raises(ops.NrtError):
        ops.PackedMLP(3, 0, 16, 128, 5, 3, 65, ops.ACT_LEAKY_RELU, m1.basis, m1.params[:-1])   # truncated blob


def test_non_contiguous_inputs_give_the_same_result():
    import torch
    from neural_raytracing_b200 import ops
    w1, _ = helpers.nerfle_weights(False)
    m1 = helpers.cuda_mlp(w1)
    big = 0.5 * torch.randn(300, 6, device="cuda")
    a = ops.mlp_forward(m1, big[:, :3])                                   # strided view
    b = ops.mlp_forward(m1, big[:, :3].contiguous())
    assert torch.equal(a, b)
    s = helpers.cuda_sdf(helpers.golden_sdf_weights())
    rays = _t(synth.camera_rays(3, 500))
    wide = torch.cat([rays, rays], dim=-1)
    d1, h1 = ops.sphere_trace(s, wide[:, :6], 1e-3, 32, 10.0)
    d2, h2 = ops.sphere_trace(s, rays, 1e-3, 32, 10.0)
    assert torch.equal(d1, d2) and torch.equal(h1, h2)


def test_unsupported_activation_and_arbitrary_callable_sdf():
    """An activation the kernels do not implement is an error naming the supported ones; an SDF that is an arbitrary
    callable (edit_dtu.py:86-102) takes the generic unfused march and matches the reference's loop semantics."""
    import torch
    from neural_raytracing_b200 import ops
    from neural_raytracing_b200.pathtracer import neural_blocks as nb
    from neural_raytracing_b200.pathtracer.shapes import sdfs
    m = nb.SkipConnMLP(device="cuda", in_size=3, out=1, num_layers=2, hidden_size=32, activation=torch.tanh).to("cuda")
    with pytest.raises(ops.NrtError) as e:
        with torch.no_grad():
            m(torch.zeros(4, 3, device="cuda"))
    assert "leaky_relu" in str(e.value)
    shape = sdfs.SDF(sdf=lambda p: p.norm(dim=-1) - 0.5, device="cuda", max_steps=48)
    rays = _t(synth.camera_rays(5, 64).reshape(1, 8, 8, 1, 6))
    it, hit = shape.intersect(rays)
    o, d = rays[..., :3], rays[..., 3:]
    # analytic ray / sphere(0.5) intersection
    b = (o * d).sum(-1)
    disc = b * b - ((o * o).sum(-1) - 0.25)
    expect = disc > 0
    assert (hit == expect).float().mean().item() > 0.95
    t_true = (-b - disc.clamp(min=0).sqrt())[hit & expect]
    assert (it.t.reshape(hit.shape)[hit & expect] - t_true).abs().max().item() < 5e-3
    assert it.raw_normals.shape[0] == int(hit.sum())


def test_zero_steps_and_degenerate_rays():
    import torch
    from neural_raytracing_b200 import ops
    s = helpers.cuda_sdf(helpers.golden_sdf_weights())
    rays = _t(synth.camera_rays(3, 100))
    d, h = ops.sphere_trace(s, rays, 1e-3, 0, 10.0)            # max_steps = 0: nothing marches, nothing hits
    assert not h.any() and torch.equal(d, torch.zeros_like(d))
    zero_dir = rays.clone(); zero_dir[:, 3:] = 0                 # degenerate direction: finite results, no hang
    d, h = ops.sphere_trace(s, zero_dir, 1e-3, 16, 10.0)
    assert torch.isfinite(d).all()
    for prec in ("f32", "f16"):
        nb = ops.shadow_test(s, zero_dir, torch.full((100,), 1.0, device="cuda"), 1e-3, 16, prec=prec)
        assert nb.dtype == torch.bool and nb.numel() == 100
