"""GPU tests (-m gpu) of the fused fp32 MLP backward kernel against FLOAT64 torch autograd (CPU) of the same op
sequence (the reference's SkipConnMLP.forward, neural_blocks.py:75-86, re-stated in forward_reference_ops).
float64 on the CPU is used as the yardstick because eager fp32 matmuls on the GPU are themselves only accurate to
~1e-3 for some shapes (cuBLAS algorithm choice; see tools/bwd_diag.py).  Tolerance: weight gradients are reduced
with fp32 atomics, so cosine >= 0.99999 and max-abs <= 2e-4 of the gradient's max magnitude."""
import copy

import numpy as np
import pytest

import helpers
import synth

pytestmark = pytest.mark.gpu

CASES = {
    "nerf_first": (dict(in_size=3, out=65, num_layers=5, hidden_size=128, freqs=16), None, 0),
    "nerf_second": (dict(in_size=70, out=3, num_layers=8, hidden_size=64, freqs=16), None, 1),
    "sdf_shift": (dict(in_size=3, out=1, num_layers=8, hidden_size=128, freqs=32), "softplus", 0),
    "neural_bsdf": (dict(in_size=3, out=3, num_layers=6, hidden_size=96, freqs=64), None, 1),
    "latent_small": (dict(in_size=3, out=9, num_layers=5, hidden_size=32, freqs=16, latent_size=8), None, 3),
    "light_field": (dict(in_size=3, out=3, num_layers=10, hidden_size=256, freqs=16), None, 0),
    "one_layer": (dict(in_size=5, out=1, num_layers=1, hidden_size=64, freqs=16), None, 1),
}


def _close(a, b, name, rtol=2e-4, min_cos=0.99999):
    a, b = a.detach().cpu().numpy().ravel().astype(np.float64), b.detach().cpu().numpy().ravel().astype(np.float64)
    scale = max(np.abs(b).max(), 1e-12)
    cos = float(a @ b / (np.linalg.norm(a) * np.linalg.norm(b) + 1e-30))
    assert cos > min_cos, (name, cos)
    assert np.abs(a - b).max() <= rtol * scale + 1e-9, (name, np.abs(a - b).max(), scale)


# NeRFLE.second Fourier-encodes 70 inputs at sigma = 32 (phases of hundreds of radians).  Measured against float64
# (tools/bwd_diag2.py): weight gradients agree to 2e-5 but the INPUT gradient, which multiplies by the basis and
# cancels, is only accurate to 7e-2 of its maximum in fp32 -- for PyTorch's own fp32 CPU path exactly as for the
# fused kernel (7.09e-2 both).  So: tight on the weights, loose on d/dx for that network.
LOOSE = dict(rtol=1e-3, min_cos=0.99999)
LOOSE_X = dict(rtol=0.15, min_cos=0.999)


@pytest.mark.parametrize("name", list(CASES))
@pytest.mark.parametrize("M", [1, 200])
def test_fused_backward_matches_autograd(name, M):
    import torch
    import torch.nn.functional as F
    from neural_raytracing_b200.pathtracer import neural_blocks as nb
    kw, act, out_act = CASES[name]
    kw = dict(kw)
    if act == "softplus":
        kw["activation"] = F.softplus
    torch.manual_seed(0)
    mlp = nb.SkipConnMLP(device="cuda", **kw).to("cuda")
    synth.fill_module(mlp, 7)
    g = torch.Generator("cuda").manual_seed(M)
    x = (0.5 * torch.randn(M, kw["in_size"], device="cuda", generator=g)).requires_grad_()
    lat = None
    if kw.get("latent_size", 0):
        lat = (0.5 * torch.randn(M, kw["latent_size"], device="cuda", generator=g)).requires_grad_()
    go = torch.randn(M, kw["out"], device="cuda", generator=g)

    nb._FUSED_BACKWARD[0] = True
    y = mlp(x, lat, out_act=out_act)
    assert type(y.grad_fn).__name__.startswith("_FusedMLP")     # the fused autograd path is the one under test
    (y * go).sum().backward()
    # float64 reference on the CPU
    m64 = copy.deepcopy(mlp).cpu().double()
    m64.basis_p = mlp.basis_p.detach().cpu().double()
    x64 = x.detach().cpu().double().requires_grad_()
    l64 = lat.detach().cpu().double().requires_grad_() if lat is not None else None
    y64 = m64.forward_reference_ops(x64, l64)
    y64 = [y64, y64.sigmoid(), F.softplus(y64), y64.tanh()][out_act]
    (y64 * go.cpu().double()).sum().backward()
    assert (y.detach().cpu().double() - y64.detach()).abs().max().item() < (5e-4 if name == "nerf_second" else 2e-5)
    tol = LOOSE if name == "nerf_second" else {}
    for (pname, p32), p64 in zip(mlp.named_parameters(), m64.parameters()):
        _close(p32.grad, p64.grad, pname, **tol)
    _close(x.grad, x64.grad, "x", **(LOOSE_X if name == "nerf_second" else {}))
    if lat is not None:
        _close(lat.grad, l64.grad, "latent", **tol)


def test_nerfle_training_step_gradients():
    """nerfle.py-style step: NeRFLE forward under autograd (fused MLP fwd/bwd kernels + CUDA compositing fwd/bwd)
    vs the same step with torch autograd through the reference op sequence."""
    import random
    import torch
    from neural_raytracing_b200.pathtracer import neural_blocks as nb
    from neural_raytracing_b200.pathtracer.shapes.nerf import NeRFLE
    from neural_raytracing_b200.pathtracer.lights import PointLights
    random.random = lambda: 0.37
    torch.manual_seed(0)
    n = NeRFLE(device="cuda")
    synth.fill_module(n, 3)
    with torch.no_grad():
        n.first.out.bias[0] = 0.8
    rays = torch.from_numpy(synth.camera_rays(5, 2 * 8 * 8).reshape(2, 8, 8, 1, 6)).cuda()
    lights = PointLights(device="cuda", location=torch.tensor([[0.4, 1.0, 0.3], [-0.8, 0.5, 0.6]], device="cuda"), scale=10)

    nb._FUSED_BACKWARD[0] = True
    rgb = n(rays, lights)
    loss = torch.nn.functional.mse_loss(rgb, torch.full_like(rgb, 0.5))
    loss.backward()
    # float64 CPU restatement of nerf.py:175-214 with the same weights
    n64 = copy.deepcopy(n).cpu().double()
    n64.first.basis_p = n.first.basis_p.detach().cpu().double()
    n64.second.basis_p = n.second.basis_p.detach().cpu().double()
    r64 = rays.cpu().double()
    ts = torch.linspace(0, 2 + 0.37 * 0.1, 64).double()
    r_o, r_d = r64.split([3, 3], dim=-1)
    pts = r_o.unsqueeze(0) + torch.tensordot(ts, r_d, dims=0)
    f = n64.first.forward_reference_ops(pts)
    lat, alpha = f[..., 1:], f[..., 0]
    lead = lat.shape[:-1]
    light = lights.location.detach().cpu().double()[None, :, None, None, None, :].expand(lead + (3,))
    c = n64.second.forward_reference_ops(torch.cat([lat, r_d[None].expand(lead + (3,)), light], dim=-1)).sigmoid()
    from neural_raytracing_b200.pathtracer.shapes.nerf import composite_reference_ops
    rgb64 = composite_reference_ops(alpha, c, ts)
    loss64 = torch.nn.functional.mse_loss(rgb64, torch.full_like(rgb64, 0.5))
    loss64.backward()
    assert abs(loss.item() - loss64.item()) < 1e-5
    for (pname, p32), p64 in zip(n.named_parameters(), n64.parameters()):
        _close(p32.grad, p64.grad, pname, **(LOOSE_X if pname.startswith("first.") else LOOSE))


@pytest.mark.parametrize("case", ["sdf_softplus_128", "leaky_64", "softplus_32_out2"])
@pytest.mark.parametrize("M", [1, 16, 203])
def test_value_jacobian_reverse_pass_matches_double_backward(case, M):
    """nrt_mlp_value_jac_forward / _backward (forward-mode Jacobian + hand-written reverse pass) vs FLOAT64 torch
    autograd: jac = autograd.grad(y, p, create_graph=True) and a loss on (y, jac) back-propagated into the weights --
    the double backward the reference runs through SDF.autograd_diff (sdfs.py:184-197) for eikonal_loss
    (utils.py:294).  Ragged M (1, 16 = one full tile, 203) covers partial tiles."""
    import torch
    import torch.nn.functional as F
    from neural_raytracing_b200 import ops
    from neural_raytracing_b200.pathtracer import neural_blocks as nb
    kw = {"sdf_softplus_128": dict(in_size=3, out=1, num_layers=8, hidden_size=128, freqs=32, activation=F.softplus),
          "leaky_64": dict(in_size=3, out=1, num_layers=5, hidden_size=64, freqs=16),
          "softplus_32_out2": dict(in_size=3, out=2, num_layers=4, hidden_size=32, freqs=16, sigma=4,
                                   activation=F.softplus)}[case]
    torch.manual_seed(0)
    mlp = nb.SkipConnMLP(device="cuda", **kw).to("cuda")
    synth.fill_module(mlp, 11)
    if kw.get("sigma", 32) == 32:
        mlp.basis_p = mlp.basis_p * 0.25        # keeps the Jacobian O(1..10) so that absolute tolerances mean something
    g = torch.Generator("cuda").manual_seed(M + 1)
    p = 0.5 * torch.randn(M, 3, device="cuda", generator=g)
    gv = torch.randn(M, kw["out"], device="cuda", generator=g)
    gj = torch.randn(M, kw["out"], 3, device="cuda", generator=g)

    pk = mlp.packed()
    val, jac, acts = ops.mlp_value_jac_forward(pk, p, save_acts=True)
    g_params = ops.mlp_value_jac_backward(pk, p, acts, gv, gj)
    gW, gb = pk.unpack(g_params)

    m64 = copy.deepcopy(mlp).cpu().double()
    m64.basis_p = mlp.basis_p.detach().cpu().double()
    p64 = p.detach().cpu().double().requires_grad_()
    y64 = m64.forward_reference_ops(p64)
    rows = []
    for n in range(kw["out"]):
        jn, = torch.autograd.grad(y64[:, n].sum(), p64, create_graph=True)
        rows.append(jn)
    j64 = torch.stack(rows, dim=1)                                        # [M, out, 3]
    scale_v, scale_j = y64.abs().max().item() + 1e-6, j64.abs().max().item() + 1e-6
    assert (val.cpu().double() - y64.detach()).abs().max().item() < 2e-5 * max(1.0, scale_v)
    assert (jac.cpu().double() - j64.detach()).abs().max().item() < 1e-4 * max(1.0, scale_j)
    ((y64 * gv.cpu().double()).sum() + (j64 * gj.cpu().double()).sum()).backward()
    lin64 = [m64.init] + list(m64.layers) + [m64.out]
    for i, (w, b, l64) in enumerate(zip(gW, gb, lin64)):
        _close(w, l64.weight.grad, "W%d" % i, rtol=5e-4, min_cos=0.9999)
        _close(b, l64.bias.grad, "b%d" % i, rtol=5e-4, min_cos=0.9999)


def test_sdf_normals_with_graph_use_the_jacobian_kernels():
    """SDF.autograd_diff with gradients enabled (training): normals equal the no-grad analytic kernel's, and an
    eikonal + normal-dependent loss reaches SphereSDF.shift, centers, radii and tfs like the reference-op autograd."""
    import torch
    from neural_raytracing_b200.pathtracer.shapes import sdfs
    from neural_raytracing_b200.pathtracer.utils import eikonal_loss
    torch.manual_seed(0)
    s = sdfs.SphereSDF(n=64, device="cuda")
    synth.fill_module(s.shift, 5)
    with torch.no_grad():
        for lin in [s.shift.init] + list(s.shift.layers) + [s.shift.out]:
            lin.weight.mul_(0.5)
        s.tfs.add_(0.05 * torch.randn_like(s.tfs))
    shape = sdfs.SDF(sdf=s, device="cuda")
    p = (0.35 * torch.randn(300, 3, device="cuda"))
    with torch.no_grad():
        n0 = shape.autograd_diff(p.clone())
    n1 = shape.autograd_diff(p.clone())
    assert n1.grad_fn is not None
    assert (n0 - n1).abs().max().item() < 1e-4 * max(1.0, n0.abs().max().item())
    w = torch.randn(300, 3, device="cuda")

    def loss_of(n):
        return eikonal_loss(n) + 0.1 * (torch.nn.functional.normalize(n, dim=-1) * w).sum(-1).mean()

    loss_of(n1).backward()
    got = {k: v.grad.detach().clone() for k, v in s.named_parameters()}
    s.zero_grad()
    # reference-op autograd (create_graph through the whole SphereSDF), fp32 on the GPU
    pr = p.clone().requires_grad_()
    out = s.forward_reference_ops(pr)
    nr, = torch.autograd.grad(out, pr, torch.ones_like(out), create_graph=True)
    loss_of(nr).backward()
    for k, v in s.named_parameters():
        if k == "shift.out.bias":       # a constant offset of the SDF: the normals do not depend on it
            assert got[k].abs().max().item() == 0 and (v.grad is None or v.grad.abs().max().item() == 0)
            continue
        assert got[k].abs().max().item() > 0, k
        _close(got[k], v.grad, k, rtol=5e-3, min_cos=0.999)


def test_fused_backward_skips_zero_gradient_tiles_exactly():
    """Tiles whose incoming gradient is exactly zero are skipped by the fused backward (masked rays of a DTU-style crop):
    weight gradients equal those of the compacted batch, input / latent gradients of the skipped samples are zero."""
    import torch
    from neural_raytracing_b200 import ops
    from neural_raytracing_b200.pathtracer import neural_blocks as nb
    torch.manual_seed(0)
    mlp = nb.SkipConnMLP(device="cuda", in_size=3, out=4, num_layers=5, hidden_size=64, freqs=16, latent_size=8).to("cuda")
    synth.fill_module(mlp, 13)
    M = 1000
    g = torch.Generator("cuda").manual_seed(1)
    x = 0.5 * torch.randn(M, 3, device="cuda", generator=g)
    lat = 0.5 * torch.randn(M, 8, device="cuda", generator=g)
    go = torch.randn(M, 4, device="cuda", generator=g)
    live = torch.zeros(M, dtype=torch.bool, device="cuda")
    live[128:200] = True          # straddles tile boundaries (64-sample tiles): partial and full zero tiles around it
    live[777] = True
    go = go * live[:, None]
    pk = mlp.packed()
    out, acts = ops.mlp_forward(pk, x, lat, prec="f32", save_acts=True)
    gp, gx, gl = ops.mlp_backward(pk, x, lat, out, acts, go, need_input_grad=True)
    xs, ls, gs = x[live].contiguous(), lat[live].contiguous(), go[live].contiguous()
    out2, acts2 = ops.mlp_forward(pk, xs, ls, prec="f32", save_acts=True)
    gp2, gx2, gl2 = ops.mlp_backward(pk, xs, ls, out2, acts2, gs, need_input_grad=True)
    _close(gp, gp2, "params", rtol=1e-5)
    assert torch.equal(gx[~live], torch.zeros_like(gx[~live])) and torch.equal(gl[~live], torch.zeros_like(gl[~live]))
    _close(gx[live], gx2, "x", rtol=1e-5)
    _close(gl[live], gl2, "latent", rtol=1e-5)


@pytest.mark.parametrize("name", ["sp_var4", "sp_var16", "light_field", "neural_bsdf", "occ"])
@pytest.mark.parametrize("M", [1, 300])
def test_wide_nets_tensor_core_training_forward(name, M):
    """config.set_train_precision("f16"): the 256-wide nets, NeuralBSDF.mlp and the occlusion MLP run their training
    forward on the tcgen05 kernels, which save the operand-rounded activations in the layout of the fused fp32 backward.  Values within the 16-bit forward's
    tolerance of the exact forward, weight / input gradients cos >= 0.999 against the all-fp32 path."""
    import torch
    from neural_raytracing_b200 import config, ops
    from neural_raytracing_b200.pathtracer import neural_blocks as nb
    kw = {"sp_var4": dict(num_layers=16, hidden_size=256, freqs=128, sigma=2 << 6, in_size=3, out=4, xavier_init=True),
          "sp_var16": dict(num_layers=16, hidden_size=256, freqs=128, sigma=2 << 6, in_size=3, out=16, xavier_init=True),
          "light_field": dict(num_layers=10, hidden_size=256, freqs=16, in_size=3, out=3),
          "neural_bsdf": dict(in_size=3, out=3, num_layers=6, hidden_size=96, freqs=64),
          "occ": dict(in_size=5, out=1)}[name]
    torch.manual_seed(0)
    mlp = nb.SkipConnMLP(device="cuda", **kw).to("cuda")
    synth.fill_module(mlp, 17)
    g = torch.Generator("cuda").manual_seed(M)
    x0 = 0.4 * torch.randn(M, kw["in_size"], device="cuda", generator=g)
    go = torch.randn(M, kw["out"], device="cuda", generator=g)

    def run(tprec):
        config.set_train_precision(tprec)
        try:
            mlp.zero_grad()
            x = x0.clone().requires_grad_()
            y = mlp(x, out_act=ops.OUT_SIGMOID)
            assert type(y.grad_fn).__name__.startswith("_FusedMLP")
            (y * go).sum().backward()
            return y.detach().clone(), x.grad.clone(), {k: p.grad.clone() for k, p in mlp.named_parameters()}
        finally:
            config.set_train_precision("f32")

    y32, gx32, gp32 = run("f32")
    ops.profile_collect(); ops.profile_enable(True)
    y16, gx16, gp16 = run("f16")
    prof = ops.profile_collect(); ops.profile_enable(False)
    tc_tag = "mlp_tc_wide" if kw.get("hidden_size", 64) == 256 else "mlp_tc_generic"
    # the path under test ran: tensor-core forward + the fused fp32 backward, or (networks of config.TRAIN_TC_NETS) the
    # tensor-core training forward / dgrad / wgrad kernels
    if nb.SkipConnMLP._shape_key(mlp) in config.TRAIN_TC_NETS:
        assert all(prof.get(t, (0, 0))[1] >= 1 for t in ("mlp_tc_train_fwd", "mlp_tc_dgrad", "mlp_tc_wgrad")), prof
    else:
        assert prof.get(tc_tag, (0, 0))[1] >= 1 and prof.get("mlp_bwd_f32", (0, 0))[1] >= 1, prof
    assert (y16 - y32).abs().max().item() < 2e-3
    flat32 = torch.cat([v.reshape(-1) for v in gp32.values()]).double()
    flat16 = torch.cat([v.reshape(-1) for v in gp16.values()]).double()
    cos = float(flat32 @ flat16 / (flat32.norm() * flat16.norm() + 1e-30))
    assert cos > 0.999, cos
    cx = float(gx32.double().reshape(-1) @ gx16.double().reshape(-1) / (gx32.double().norm() * gx16.double().norm() + 1e-30))
    assert cx > 0.995, cx


@pytest.mark.parametrize("M", [1, 45, 200])
def test_wide_backward_without_input_gradient_uses_32_sample_tiles(M):
    """256-wide nets whose input carries no gradient (sp_var on it.p, LightField on the hit points) take the
    k_mlp_bwd<256, 32, GX = false> instantiation: same weight gradients as float64 autograd."""
    import torch
    from neural_raytracing_b200.pathtracer import neural_blocks as nb
    torch.manual_seed(0)
    mlp = nb.SkipConnMLP(device="cuda", in_size=3, out=3, num_layers=10, hidden_size=256, freqs=16).to("cuda")
    synth.fill_module(mlp, 7)
    g = torch.Generator("cuda").manual_seed(M)
    x = 0.5 * torch.randn(M, 3, device="cuda", generator=g)            # no requires_grad
    go = torch.randn(M, 3, device="cuda", generator=g)
    y = mlp(x, out_act=0)
    assert type(y.grad_fn).__name__.startswith("_FusedMLP")
    (y * go).sum().backward()
    m64 = copy.deepcopy(mlp).cpu().double()
    m64.basis_p = mlp.basis_p.detach().cpu().double()
    y64 = m64.forward_reference_ops(x.cpu().double())
    (y64 * go.cpu().double()).sum().backward()
    for (pname, p32), p64 in zip(mlp.named_parameters(), m64.parameters()):
        _close(p32.grad, p64.grad, pname)


@pytest.mark.parametrize("K,n", [(1, 64), (77, 64), (4000, 128)])
def test_sphere_set_kernels_match_double_backward(K, n):
    """nrt_sphere_set_forward / _backward (value + gradient of the smooth-min of warped spheres and the reverse pass of
    both outputs) vs FLOAT64 torch autograd of the reference expression (sdfs.py:37-45, utils.py:385-387) with
    autograd.grad(create_graph=True) and a loss on (value, gradient): the double backward of SDF.autograd_diff."""
    import torch
    from neural_raytracing_b200.pathtracer.shapes import sdfs
    torch.manual_seed(K)
    s = sdfs.SphereSDF(n=n, device="cuda")
    with torch.no_grad():
        s.tfs.add_(0.1 * torch.randn_like(s.tfs))
        s.radii.abs_()
    g = torch.Generator("cuda").manual_seed(K + 3)
    p = 0.15 * torch.randn(K, 3, device="cuda", generator=g)
    if K > 1:
        p[K // 2] = 5.0                   # far away: the 1e-4 clamp of smooth_min is active there (zero gradients)
    gv = torch.randn(K, device="cuda", generator=g)
    gn = torch.randn(K, 3, device="cuda", generator=g)
    val, grad = sdfs._SphereSet.apply(s.centers, s.radii, s.tfs, p, True)
    ((val * gv).sum() + (grad * gn).sum()).backward()
    got = [q.grad.clone() for q in (s.centers, s.radii, s.tfs)]
    c64, r64, t64 = (q.detach().double().requires_grad_() for q in (s.centers, s.radii, s.tfs))
    p64 = p.double().requires_grad_()
    tf = t64 + torch.eye(3, device="cuda", dtype=torch.float64).unsqueeze(0)
    q = torch.einsum("ijk,ibk->ibj", tf, p64.unsqueeze(0).expand(n, -1, -1)) - c64.unsqueeze(1)
    sd = q.norm(p=2, dim=-1) - r64.unsqueeze(-1)
    v64 = -torch.exp(-32.0 * sd).sum(0).clamp(min=1e-4).log() / 32.0
    n64, = torch.autograd.grad(v64.sum(), p64, create_graph=True)
    assert (val.double() - v64.detach()).abs().max().item() < 2e-6
    assert (grad.double() - n64.detach()).abs().max().item() < 2e-5
    assert K == 1 or grad[K // 2].abs().max().item() == 0.0
    ((v64 * gv.double()).sum() + (n64 * gn.double()).sum()).backward()
    for a, b, name in zip(got, (c64.grad, r64.grad, t64.grad), ("centers", "radii", "tfs")):
        _close(a, b, name, rtol=2e-4, min_cos=0.99999)
    # first order only (SDF.throughput's sdf(best_pos))
    for q_ in (s.centers, s.radii, s.tfs):
        q_.grad = None
    v1 = sdfs._SphereSet.apply(s.centers, s.radii, s.tfs, p, False)[0]
    (v1 * gv).sum().backward()
    c64.grad = r64.grad = t64.grad = None
    q = torch.einsum("ijk,ibk->ibj", t64 + torch.eye(3, device="cuda", dtype=torch.float64).unsqueeze(0),
                     p.double().unsqueeze(0).expand(n, -1, -1)) - c64.unsqueeze(1)
    v = -torch.exp(-32.0 * (q.norm(p=2, dim=-1) - r64.unsqueeze(-1))).sum(0).clamp(min=1e-4).log() / 32.0
    (v * gv.double()).sum().backward()
    for a, b, name in zip((s.centers.grad, s.radii.grad, s.tfs.grad), (c64.grad, r64.grad, t64.grad), ("centers", "radii", "tfs")):
        _close(a, b, name + " (first order)", rtol=2e-4, min_cos=0.99999)
