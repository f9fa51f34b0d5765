"""CPU tests (-m "not gpu") of the registered operator layer (neural_raytracing_b200/torch_ops.py): every operator of
`torch.ops.nrt_b200` exists with a schema, its fake implementation propagates shapes under FakeTensorMode (no kernel
runs), and calling one with real CPU tensors fails loudly (there is no CPU implementation)."""
import os
import sys

import pytest
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))

from neural_raytracing_b200 import ops, torch_ops  # noqa: E402

ARCH = [3, 0, 16, 64, 8, 3, 3, ops.ACT_LEAKY_RELU]


def _nparams(arch):
    return sum(k * n + n for k, n in ops.mlp_layer_dims(*arch[:7]))


def test_every_operator_is_registered_with_a_schema():
    for name in torch_ops.OPERATORS:
        op = getattr(torch.ops.nrt_b200, name)
        assert "nrt_b200::" + name in str(op.default._schema)


def test_fake_implementations_propagate_shapes():
    from torch._subclasses.fake_tensor import FakeTensorMode
    P = _nparams(ARCH)
    with FakeTensorMode():
        x = torch.empty(100, 3, device="cuda")
        params, basis = torch.empty(P, device="cuda"), torch.empty(3, 16, device="cuda")
        out, acts = torch.ops.nrt_b200.mlp_forward(x, None, params, basis, ARCH, 0, ops.PREC_F32)
        assert tuple(out.shape) == (100, 3) and tuple(acts.shape) == (9 * 64, 100)
        gp, gx, gl = torch.ops.nrt_b200.mlp_backward(x, None, out, acts, out, params, basis, ARCH, 0)
        assert tuple(gp.shape) == (P,) and tuple(gx.shape) == (100, 3) and gl.numel() == 0
        sig, rgb, ts = torch.empty(64, 50, device="cuda"), torch.empty(64, 50, 3, device="cuda"), torch.empty(64, device="cuda")
        assert tuple(torch.ops.nrt_b200.composite(sig, rgb, ts).shape) == (50, 3)
        rays = torch.empty(77, 6, device="cuda")
        c, r, t = torch.empty(64, 3, device="cuda"), torch.empty(64, device="cuda"), torch.empty(64, 3, 3, device="cuda")
        sarch = [3, 0, 32, 128, 8, 3, 1, ops.ACT_SOFTPLUS]
        sp, sb = torch.empty(_nparams(sarch), device="cuda"), torch.empty(3, 32, device="cuda")
        d, h = torch.ops.nrt_b200.sdf_sphere_trace(rays, c, r, t, sp, sb, sarch, 1e-3, 64, 10.0, ops.PREC_F32)
        assert tuple(d.shape) == (77,) and h.dtype == torch.bool
        i, pos, mv = torch.ops.nrt_b200.sdf_min_scan(rays, c, r, t, sp, sb, sarch, 0.017, 128, ops.PREC_F32)
        assert i.dtype == torch.int32 and tuple(pos.shape) == (77, 3) and tuple(mv.shape) == (77,)
        c2w = torch.empty(2, 4, 4, device="cuda")
        cr = torch.ops.nrt_b200.camera_rays(ops.CAM_NERF, c2w, None, 20.0, 16.0, 3, 2, 10, 6, 1, None, 0.0, 0)
        assert tuple(cr.shape) == (2, 10, 6, 1, 6)
        a1, a2 = [3, 0, 16, 128, 5, 3, 65, 0], [70, 0, 16, 64, 8, 3, 3, 0]
        img = torch.ops.nrt_b200.nerfle_render_camera(ops.CAM_NERF, c2w, None, 20.0, 16.0, 0, 0, 16, 16, 1, 0.0, 0, ts,
                                                      torch.empty(2, 3, device="cuda"), torch.empty(_nparams(a1), device="cuda"),
                                                      torch.empty(3, 16, device="cuda"), a1, torch.empty(_nparams(a2), device="cuda"),
                                                      torch.empty(70, 16, device="cuda"), a2, ops.PREC_F16)
        assert tuple(img.shape) == (2, 16, 16, 1, 3)
        v, j, a = torch.ops.nrt_b200.mlp_value_jac(torch.empty(9, 3, device="cuda"), sp, sb, sarch)
        assert tuple(v.shape) == (9, 1) and tuple(j.shape) == (9, 1, 3) and tuple(a.shape) == (9 * 128, 36)


def test_cpu_tensors_fail_loudly():
    P = _nparams(ARCH)
    with pytest.raises(Exception) as e:
        torch.ops.nrt_b200.mlp_forward(torch.zeros(4, 3), None, torch.zeros(P), torch.zeros(3, 16), ARCH, 0, ops.PREC_F32)
    assert "CUDA" in str(e.value) or "cuda" in str(e.value) or "libnrt" in str(e.value)
