"""GPU tests (-m gpu) of the analytic-Jacobian SDF kernel and the shading-glue kernels against the golden
outputs of the unmodified reference (tests/golden/sdf.npz, shading.npz) and the CPU oracle."""
import numpy as np
import pytest

import helpers
from oracle import c_oracle

pytestmark = pytest.mark.gpu


def _t(a):
    import torch
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


def test_sdf_value_grad_matches_autograd_reference():
    """nrt_sdf_value_grad vs torch.autograd.grad of the reference SDF (sdfs.py:184-197) and vs the oracle."""
    from neural_raytracing_b200 import ops
    g = helpers.golden("sdf")
    w = helpers.golden_sdf_weights()
    s = helpers.cuda_sdf(w)
    pts = g["pts"][:96]
    val, grad = ops.sdf_value_grad(s, _t(pts))
    val, grad = val.cpu().numpy(), grad.cpu().numpy()
    assert np.abs(val - g["sdf_vals"][:96]).max() < 2e-6
    assert np.abs(grad - g["grad_pts"]).max() < 2e-4            # fp32 forward mode vs fp32 reverse mode
    vo, go = c_oracle.sdf_value_grad(helpers.oracle_sdf(w), pts)
    assert np.abs(val - vo).max() < 1e-6 and np.abs(grad - go).max() < 5e-5
    # normals at the march's hit points (sdfs.py:152-157)
    hit = g["hit"]
    p_hit = g["rays"][hit, :3] + g["depth"][hit, None] * g["rays"][hit, 3:]
    _, gn = ops.sdf_value_grad(s, _t(p_hit))
    assert np.abs(gn.cpu().numpy() - g["raw_normals"]).max() < 5e-4


@pytest.mark.parametrize("M", [0, 1, 15, 16, 17, 1000])
def test_sdf_value_grad_ragged(M):
    from neural_raytracing_b200 import ops
    w = helpers.golden_sdf_weights()
    rs = np.random.RandomState(M)
    pts = (0.5 * rs.standard_normal((M, 3))).astype(np.float32)
    val, grad = ops.sdf_value_grad(helpers.cuda_sdf(w), _t(pts))
    assert val.shape == (M,) and grad.shape == (M, 3)
    if M:
        v2 = ops.sdf_eval(helpers.cuda_sdf(w), _t(pts))
        assert np.abs(val.cpu().numpy() - v2.cpu().numpy()).max() < 1e-6
        # finite-difference sanity check of the Jacobian (central differences in fp32 through sigma=32 Fourier
        # features: only the bulk of the points can be expected to agree closely)
        eps = 2e-4
        for c in range(3):
            d = np.zeros((1, 3), np.float32); d[0, c] = eps
            fd = (ops.sdf_eval(helpers.cuda_sdf(w), _t(pts + d)) - ops.sdf_eval(helpers.cuda_sdf(w), _t(pts - d))) / (2 * eps)
            err = np.abs(fd.cpu().numpy() - grad[:, c].cpu().numpy())
            assert np.median(err) < 5e-3 and err.max() < 0.5, (np.median(err), err.max())


def test_shading_frame_and_to_local_match_reference():
    import torch
    from neural_raytracing_b200 import ops
    g = helpers.golden("shading")
    frame = ops.shading_frame(_t(g["n"]))
    assert np.abs(frame.cpu().numpy() - g["frame"]).max() < 2e-6
    loc = ops.to_local(frame, _t(g["v"]))
    assert np.abs(loc.cpu().numpy() - g["to_local"]).max() < 2e-6
    rays = torch.cat([torch.zeros(64, 3, device="cuda"), -_t(g["v"])], dim=-1)
    frame2, wi = ops.shading_frame(_t(g["n"]), rays)
    assert torch.equal(frame2, frame)
    assert np.abs(wi.cpu().numpy() - g["to_local"]).max() < 2e-6


def test_param_rusin2_matches_reference():
    from neural_raytracing_b200 import ops
    g = helpers.golden("shading")
    out = ops.param_rusin2(_t(g["rusin_a"]), _t(g["rusin_b"])).cpu().numpy()
    assert np.abs(out - g["rusin"]).max() < 5e-6


def test_package_glue_uses_kernels_without_grad_and_torch_with_grad():
    """Same values either way; with requires_grad the differentiable torch glue is used."""
    import torch
    from neural_raytracing_b200.pathtracer.interaction import coordinate_system, to_local
    from neural_raytracing_b200.pathtracer.utils import param_rusin2
    g = helpers.golden("shading")
    n, v = _t(g["n"]), _t(g["v"])
    f0 = coordinate_system(n)
    f1 = coordinate_system(n.clone().requires_grad_())
    assert np.abs(f0.cpu().numpy() - f1.detach().cpu().numpy()).max() < 2e-6
    l0 = to_local(f0, v)
    l1 = to_local(f1, v)
    assert l1.requires_grad and np.abs(l0.cpu().numpy() - l1.detach().cpu().numpy()).max() < 2e-6
    a, b = _t(g["rusin_a"]), _t(g["rusin_b"])
    r0 = param_rusin2(a, b)
    r1 = param_rusin2(a.clone().requires_grad_(), b)
    assert np.abs(r0.cpu().numpy() - r1.detach().cpu().numpy()).max() < 5e-6
