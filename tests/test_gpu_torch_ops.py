"""GPU tests (-m gpu) of `torch.ops.nrt_b200.*`: the registered forward equals the class layer bit for bit, the
registered backward delivers the same gradients to the nn.Linear weights (through the differentiable pack), and
torch.library.opcheck accepts the registrations (schema, fake implementation, autograd registration)."""
import copy

import numpy as np
import pytest

import helpers
import synth

pytestmark = pytest.mark.gpu


def _mlp(seed=3, **kw):
    import torch
    from neural_raytracing_b200.pathtracer import neural_blocks as nb
    torch.manual_seed(0)
    m = nb.SkipConnMLP(device="cuda", **kw).to("cuda")
    synth.fill_module(m, seed)
    return m


def test_mlp_operator_forward_and_registered_backward():
    import torch
    from neural_raytracing_b200 import ops, torch_ops as T
    m = _mlp(in_size=3, out=3, num_layers=6, hidden_size=96, freqs=64)
    x = (0.5 * torch.randn(300, 3, device="cuda")).requires_grad_()
    go = torch.randn(300, 3, device="cuda")
    y_ref = m(x, out_act=ops.OUT_SIGMOID)                       # class layer: _FusedMLP
    (y_ref * go).sum().backward()
    ref = {k: p.grad.clone() for k, p in m.named_parameters()}
    gx_ref = x.grad.clone()
    m.zero_grad(); x.grad = None
    out, _acts = torch.ops.nrt_b200.mlp_forward(x, None, T.pack_module(m), m.basis_p, T.arch_of(m), ops.OUT_SIGMOID, ops.PREC_F32)
    assert torch.equal(out, y_ref.detach())
    (out * go).sum().backward()
    for k, p in m.named_parameters():
        a, b = p.grad, ref[k]
        assert (a - b).abs().max().item() <= 1e-5 * max(1.0, b.abs().max().item()), k     # fp32 atomics order
    assert (x.grad - gx_ref).abs().max().item() <= 1e-5 * max(1.0, gx_ref.abs().max().item())
    # tensor-core precision through the operator: inference only, same kernel as the class layer
    with torch.no_grad():
        o16, a16 = torch.ops.nrt_b200.mlp_forward(x.detach(), None, T.pack_module(m), m.basis_p, T.arch_of(m), 0, ops.PREC_F16)
    assert a16.numel() == 0 and (o16 - m.forward_reference_ops(x.detach())).abs().max().item() < 5e-3


@pytest.mark.parametrize("kw,out_act,need_x", [(dict(in_size=3, out=3, num_layers=6, hidden_size=96, freqs=64), 1, True),
                                               (dict(in_size=3, out=65, num_layers=5, hidden_size=128, freqs=16), 0, False)])
def test_tensor_core_training_operator_has_forward_and_backward_registered(kw, out_act, need_x):
    """nrt_b200::mlp_forward_train_tc under autograd (tcgen05 forward with saved tiles, streamed dgrad + wgrad) against the
    fp32 operator on the same weights: outputs within 1e-3, every parameter gradient and the input gradient cosine >= 0.999."""
    import torch
    from neural_raytracing_b200 import ops, torch_ops as T
    m = _mlp(**kw)
    g = torch.Generator(device="cuda").manual_seed(5)
    x0 = 0.5 * torch.randn(4096, 3, device="cuda", generator=g)
    go = torch.randn(4096, kw["out"], device="cuda", generator=g)
    res = {}
    for name, prec in (("f32", ops.PREC_F32), ("f16", ops.PREC_F16)):
        m.zero_grad()
        x = x0.clone().requires_grad_(need_x)      # NeRFLE.first: ray samples, no input gradient on the tensor-core path
        if prec == ops.PREC_F32:
            out, _ = torch.ops.nrt_b200.mlp_forward(x, None, T.pack_module(m), m.basis_p, T.arch_of(m), out_act, prec)
        else:
            out, ws = torch.ops.nrt_b200.mlp_forward_train_tc(x, T.pack_module(m), m.basis_p, T.arch_of(m), out_act, prec)
            assert ws.dtype == torch.uint8 and ws.numel() > 0
        (out * go).sum().backward()
        res[name] = (out.detach(), x.grad.clone() if need_x else None, {k: p.grad.clone() for k, p in m.named_parameters()})
    assert (res["f16"][0] - res["f32"][0]).abs().max().item() < 1e-3
    cos = lambda a, b: float((a.double().flatten() @ b.double().flatten()) / (a.double().norm() * b.double().norm() + 1e-300))
    if need_x:
        assert cos(res["f16"][1], res["f32"][1]) > 0.999
    for k in res["f32"][2]:
        assert cos(res["f16"][2][k], res["f32"][2][k]) > 0.999, k


def test_composite_and_value_jac_operators():
    import torch
    import torch.nn.functional as F
    from neural_raytracing_b200 import ops, torch_ops as T
    from neural_raytracing_b200.pathtracer.shapes.nerf import composite_reference_ops
    g = torch.Generator(device="cuda").manual_seed(0)
    sig = (torch.randn(64, 500, device="cuda", generator=g) * 2).requires_grad_()
    rgb = torch.rand(64, 500, 3, device="cuda", generator=g).requires_grad_()
    ts = torch.linspace(0.02, 2.0, 64, device="cuda")
    out = torch.ops.nrt_b200.composite(sig, rgb, ts)
    ref = composite_reference_ops(sig, rgb, ts)
    assert (out - ref).abs().max().item() < 1e-5
    w = torch.randn(500, 3, device="cuda", generator=g)
    gs, gc = torch.autograd.grad((out * w).sum(), [sig, rgb])
    gs_r, gc_r = torch.autograd.grad((ref * w).sum(), [sig, rgb])
    assert (gs - gs_r).abs().max().item() < 1e-4 * max(1.0, gs_r.abs().max().item())
    assert (gc - gc_r).abs().max().item() < 1e-5
    # value + Jacobian of the SDF residual net, gradients into the Linear weights through the registered backward
    m = _mlp(seed=5, in_size=3, out=1, num_layers=8, hidden_size=128, freqs=32, activation=F.softplus)
    m.basis_p = m.basis_p * 0.25
    p = 0.4 * torch.randn(77, 3, device="cuda", generator=g)
    val, jac, _ = torch.ops.nrt_b200.mlp_value_jac(p, T.pack_module(m), m.basis_p, T.arch_of(m))
    ((jac.norm(dim=-1) - 1).square().mean() + val.mean()).backward()
    got = {k: q.grad.clone() for k, q in m.named_parameters()}
    m.zero_grad()
    pr = p.clone().requires_grad_()
    y = m.forward_reference_ops(pr)
    j, = torch.autograd.grad(y.sum(), pr, create_graph=True)
    ((j.reshape(77, 1, 3).norm(dim=-1) - 1).square().mean() + y.mean()).backward()
    for k, q in m.named_parameters():
        scale = max(q.grad.abs().max().item(), 1e-9)
        assert (got[k] - q.grad).abs().max().item() <= 5e-3 * scale, k


def test_scan_operators_equal_the_class_layer():
    import torch
    from neural_raytracing_b200 import ops, torch_ops as T
    from neural_raytracing_b200.pathtracer.shapes import sdfs
    torch.manual_seed(0)
    s = sdfs.SphereSDF(n=64, device="cuda")
    synth.fill_module(s, 61, shift_std=0.02)
    rays = torch.from_numpy(synth.camera_rays(3, 3000)).cuda()
    args = (s.centers.detach(), s.radii.detach(), s.tfs.detach(), T.pack_module(s.shift).detach(), s.shift.basis_p, T.arch_of(s.shift))
    d, h = torch.ops.nrt_b200.sdf_sphere_trace(rays, *args, 1e-3, 64, 10.0, ops.PREC_F32)
    d2, h2 = ops.sphere_trace(s.packed(), rays, 1e-3, 64, 10.0)
    assert torch.equal(d, d2) and torch.equal(h, h2) and 0 < int(h.sum()) < 3000
    nb = torch.ops.nrt_b200.sdf_shadow_test(rays, torch.full((3000,), 2.0, device="cuda"), *args, 1e-3, 64, ops.PREC_F32)
    assert torch.equal(nb, ops.shadow_test(s.packed(), rays, torch.full((3000,), 2.0, device="cuda"), 1e-3, 64))
    i, pos, mv = torch.ops.nrt_b200.sdf_min_scan(rays, *args, 2.2 / 128, 128, ops.PREC_F32)
    i2, pos2, mv2 = ops.min_scan(s.packed(), rays, 2.2 / 128, 128)
    assert torch.equal(i, i2) and torch.equal(pos, pos2) and torch.equal(mv, mv2)
    v = torch.ops.nrt_b200.sdf_eval(rays[:, :3].contiguous(), *args, ops.PREC_F32)
    assert torch.equal(v, ops.sdf_eval(s.packed(), rays[:, :3].contiguous()))


def test_opcheck_accepts_the_registrations():
    import torch
    from neural_raytracing_b200 import ops, torch_ops as T
    m = _mlp(in_size=3, out=3, num_layers=6, hidden_size=96, freqs=64)
    x = (0.5 * torch.randn(64, 3, device="cuda")).requires_grad_()
    params = T.pack_module(m).detach().requires_grad_()
    torch.library.opcheck(torch.ops.nrt_b200.mlp_forward.default,
                          (x, None, params, m.basis_p, T.arch_of(m), 0, ops.PREC_F32),
                          test_utils=("test_schema", "test_faketensor", "test_autograd_registration"))
    sig = torch.randn(16, 40, device="cuda", requires_grad=True)
    rgb = torch.rand(16, 40, 3, device="cuda", requires_grad=True)
    torch.library.opcheck(torch.ops.nrt_b200.composite.default, (sig, rgb, torch.linspace(0.1, 2, 16, device="cuda")),
                          test_utils=("test_schema", "test_faketensor", "test_autograd_registration"))


def test_camera_operators_equal_the_ctypes_layer():
    """torch.ops.nrt_b200.camera_rays / nerfle_render_camera (registered, with fake implementations) == ops.camera_rays /
    ops.nerfle_render_camera on the same camera."""
    import torch
    from neural_raytracing_b200 import ops, torch_ops  # noqa: F401
    w1, w2 = helpers.nerfle_weights(False)
    m1, m2 = helpers.cuda_mlp(w1), helpers.cuda_mlp(w2)
    c2w, focal = synth.nerf_cameras(2, 16, device="cuda")
    desc = ops.CameraDesc(ops.CAM_NERF, c2w, None, focal=focal, size=16, x0=1, y0=2, nx=9, ny=7)
    rays = torch.ops.nrt_b200.camera_rays(ops.CAM_NERF, c2w, None, float(focal), 16.0, 1, 2, 9, 7, 1, None, 0.0, 0)
    assert torch.equal(rays, ops.camera_rays(desc))
    ts = torch.linspace(0, 2.05, 64, device="cuda")
    code = torch.tensor([[0.4, 1.0, 0.3], [-0.8, 0.5, 0.6]], device="cuda")
    a1 = [w1["in_size"], 0, w1["freqs"], w1["hidden"], w1["num_layers"], w1["skip"], w1["out"], ops.ACT_LEAKY_RELU]
    a2 = [w2["in_size"], 0, w2["freqs"], w2["hidden"], w2["num_layers"], w2["skip"], w2["out"], ops.ACT_LEAKY_RELU]
    img = torch.ops.nrt_b200.nerfle_render_camera(ops.CAM_NERF, c2w, None, float(focal), 16.0, 1, 2, 9, 7, 1, 0.0, 0, ts, code,
                                                  m1.params, m1.basis, a1, m2.params, m2.basis, a2, ops.PREC_F16)
    assert torch.equal(img, ops.nerfle_render_camera(m1, m2, desc, ts, code, prec="f16"))
    torch.library.opcheck(torch.ops.nrt_b200.camera_rays.default,
                          (ops.CAM_NERF, c2w, None, float(focal), 16.0, 1, 2, 9, 7, 1, None, 0.0, 0),
                          test_utils=("test_schema", "test_faketensor"))
