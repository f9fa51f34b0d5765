"""GPU parity tests (-m gpu): the CUDA fp32 path, called through the C ABI of libnrt_b200.so,
against (a) the CPU oracle on the same inputs -- BIT-EXACT, the fp32 kernels follow the oracle's
fixed fma order -- and (b) the golden outputs of the unmodified reference (fp32 tolerance)."""
import numpy as np
import pytest

import helpers
import synth
from oracle import c_oracle

pytestmark = pytest.mark.gpu


def _t(a, dtype=None):
    import torch
    t = torch.from_numpy(np.ascontiguousarray(a)).cuda()
    return t if dtype is None else t.to(dtype)


def _bits(a):
    return np.ascontiguousarray(a, np.float32).view(np.uint32)


@pytest.mark.parametrize("name", list(helpers.MLP_CASES))
def test_mlp_forward_bitexact_vs_oracle_and_golden(name):
    from neural_raytracing_b200 import ops
    g = helpers.golden("mlp")
    kw, act = helpers.MLP_CASES[name]
    w = synth.mlp_weights(**kw)
    x = g[name + "_x"]
    lat = g[name + "_latent"] if (name + "_latent") in g.files else None
    y = ops.mlp_forward(helpers.cuda_mlp(w, act), _t(x), _t(lat) if lat is not None else None).cpu().numpy()
    yo = c_oracle.mlp_forward(helpers.oracle_mlp(w, act), x, lat)
    assert np.array_equal(_bits(y), _bits(yo)), "max abs diff %g" % np.abs(y - yo).max()
    tol = 5e-4 if name == "sp_var_small" else 2e-5
    assert np.abs(y - g[name + "_y"]).max() < tol


@pytest.mark.parametrize("M", [0, 1, 63, 64, 65, 1000])
def test_mlp_forward_ragged_sizes(M):
    import torch
    from neural_raytracing_b200 import ops
    kw, act = helpers.MLP_CASES["latent_small"]
    w = synth.mlp_weights(**kw)
    rs = np.random.RandomState(M)
    x = rs.standard_normal((M, 3)).astype(np.float32)
    lat = rs.standard_normal((M, 8)).astype(np.float32)
    for out_act in (ops.OUT_NONE, ops.OUT_SIGMOID, ops.OUT_SOFTPLUS, ops.OUT_TANH):
        y = ops.mlp_forward(helpers.cuda_mlp(w, act), _t(x), _t(lat), out_act=out_act)
        assert y.shape == (M, 9)
        if M:
            yo = c_oracle.mlp_forward(helpers.oracle_mlp(w, act), x, lat, out_act=out_act)
            assert np.array_equal(_bits(y.cpu().numpy()), _bits(yo))
    torch.cuda.synchronize()


def test_mlp_rejects_cpu_tensors():
    import torch
    from neural_raytracing_b200 import ops
    kw, act = helpers.MLP_CASES["one_layer"]
    m = helpers.cuda_mlp(synth.mlp_weights(**kw), act)
    with pytest.raises(ops.NrtError):
        ops.mlp_forward(m, torch.zeros(4, 5))            # CPU tensor: no fallback
    with pytest.raises(ops.NrtError):
        ops.mlp_forward(m, torch.zeros(4, 5, device="cuda", dtype=torch.float64))


def test_sdf_eval_bitexact():
    from neural_raytracing_b200 import ops
    g = helpers.golden("sdf")
    w = helpers.golden_sdf_weights()
    v = ops.sdf_eval(helpers.cuda_sdf(w), _t(g["pts"])).cpu().numpy()
    vo = c_oracle.sdf_eval(helpers.oracle_sdf(w), g["pts"])
    assert np.array_equal(_bits(v), _bits(vo)), np.abs(v - vo).max()
    assert np.abs(v - g["sdf_vals"]).max() < 2e-6


def test_sphere_trace_hit_mask_bitexact():
    import torch
    from neural_raytracing_b200 import ops
    g = helpers.golden("sdf")
    w = helpers.golden_sdf_weights()
    steps = torch.zeros(1, dtype=torch.int64, device="cuda")
    depth, hit = ops.sphere_trace(helpers.cuda_sdf(w), _t(g["rays"]), 1e-3, 64, 10.0, steps_counter=steps)
    depth, hit = depth.cpu().numpy(), hit.cpu().numpy()
    do, ho = c_oracle.sphere_trace(helpers.oracle_sdf(w), g["rays"], 1e-3, 64, 10.0)
    assert int((hit != ho).sum()) == 0                       # bit-exact mask vs the fp32 restatement
    assert np.array_equal(_bits(depth), _bits(do))           # and identical depths
    assert int((hit != g["hit"]).sum()) <= 2                 # vs the torch reference (threshold flips)
    same = hit == g["hit"]
    assert np.abs(depth[same] - g["depth"][same]).max() < 5e-4
    n = int(steps.item())
    assert 0 < n < 384 * 64                                  # compaction evaluated fewer samples than the reference


def test_sphere_trace_compaction_is_order_invariant():
    """Size-independent property at a size the oracle cannot reach: a ray's result must not depend
    on which slot / CTA / refill round it was traced in."""
    import torch
    from neural_raytracing_b200 import ops
    w = helpers.golden_sdf_weights()
    s = helpers.cuda_sdf(w)
    rays = _t(synth.camera_rays(5, 30000))
    d1, h1 = ops.sphere_trace(s, rays, 1e-3, 64, 10.0)
    perm = torch.randperm(rays.shape[0], device="cuda", generator=torch.Generator("cuda").manual_seed(0))
    d2, h2 = ops.sphere_trace(s, rays[perm].contiguous(), 1e-3, 64, 10.0)
    assert torch.equal(h1[perm], h2)
    assert torch.equal(d1[perm], d2)
    # and a small prefix agrees with the oracle bit for bit
    do, ho = c_oracle.sphere_trace(helpers.oracle_sdf(w), rays[:200].cpu().numpy(), 1e-3, 64, 10.0)
    assert np.array_equal(h1[:200].cpu().numpy(), ho)
    assert np.array_equal(_bits(d1[:200].cpu().numpy()), _bits(do))
    assert 0.05 < h1.float().mean().item() < 0.95


@pytest.mark.parametrize("R", [1, 5, 64, 65])
def test_sphere_trace_small_and_all_miss(R):
    from neural_raytracing_b200 import ops
    w = helpers.golden_sdf_weights()
    rays = synth.camera_rays(9, R)
    rays[:, 3:] = -rays[:, 3:]                              # looking away: every ray misses
    depth, hit = ops.sphere_trace(helpers.cuda_sdf(w), _t(rays), 1e-3, 16, 10.0)
    do, ho = c_oracle.sphere_trace(helpers.oracle_sdf(w), rays, 1e-3, 16, 10.0)
    assert not hit.any().item()
    assert np.array_equal(_bits(depth.cpu().numpy()), _bits(do))


def test_shadow_test_bitexact():
    from neural_raytracing_b200 import ops
    g = helpers.golden("sdf")
    w = helpers.golden_sdf_weights()
    nb = ops.shadow_test(helpers.cuda_sdf(w), _t(g["shadow_rays"]), _t(g["shadow_max_t"]), 1e-3, 64).cpu().numpy()
    nbo = c_oracle.shadow_test(helpers.oracle_sdf(w), g["shadow_rays"], g["shadow_max_t"], 1e-3, 64)
    assert int((nb != nbo).sum()) == 0
    assert int((nb != g["not_blocked"]).sum()) <= 2


def test_min_scan_bitexact():
    from neural_raytracing_b200 import ops
    g = helpers.golden("sdf")
    w = helpers.golden_sdf_weights()
    step = (float(g["scan_dist"]) + float(g["fixed_random"]) * (2 / 128)) / 128
    idx, pos, mv = ops.min_scan(helpers.cuda_sdf(w), _t(g["rays"]), step, 128)
    io, po, mo = c_oracle.min_scan(helpers.oracle_sdf(w), g["rays"], step, 128)
    assert np.array_equal(idx.cpu().numpy(), io)
    assert np.array_equal(_bits(pos.cpu().numpy()), _bits(po))
    assert np.array_equal(_bits(mv.cpu().numpy()), _bits(mo))
    close = np.abs(pos.cpu().numpy() - g["best_pos"]).max(axis=-1) < 1e-6
    assert close.mean() > 0.98


@pytest.mark.parametrize("tag", ["s64", "s5", "s1"])
def test_composite_forward_backward(tag):
    from neural_raytracing_b200 import ops
    g = helpers.golden("composite")
    sg, c, ts = g[tag + "_sigma"], g[tag + "_rgb"], g[tag + "_ts"]
    out = ops.composite_forward(_t(sg), _t(c), _t(ts)).cpu().numpy()
    assert np.array_equal(_bits(out), _bits(c_oracle.composite(sg, c, ts)))
    assert np.abs(out - g[tag + "_out"]).max() < 2e-6
    gs, gc = ops.composite_backward(_t(sg), _t(c), _t(ts), _t(g[tag + "_gout"]))
    # gradients vs torch autograd of the reference expression (nerf.py:205-213), fp32 tolerance
    assert np.abs(gc.cpu().numpy() - g[tag + "_grgb"]).max() < 1e-5
    ref = g[tag + "_gsigma"]
    assert np.abs(gs.cpu().numpy() - ref).max() < 1e-4 * max(1.0, np.abs(ref).max())


@pytest.mark.parametrize("tag", ["pt", "le"])
def test_nerfle_render_bitexact_vs_oracle_and_golden(tag):
    from neural_raytracing_b200 import ops
    g = helpers.golden("nerfle")
    w1, w2 = helpers.nerfle_weights(tag == "le")
    rays = g[tag + "_rays"]
    N, per_view = rays.shape[0], rays.shape[1] * rays.shape[2]
    view = np.repeat(np.arange(N, dtype=np.int32), per_view)
    code = g["le_light_code"] if tag == "le" else g["pt_light_loc"]
    ts = helpers.nerfle_ts(g["fixed_random"])
    rgb = ops.nerfle_render(helpers.cuda_mlp(w1), helpers.cuda_mlp(w2), _t(rays.reshape(-1, 6)), _t(ts), _t(code),
                            _t(view)).cpu().numpy()
    ro = c_oracle.nerfle_render(helpers.oracle_mlp(w1), helpers.oracle_mlp(w2), rays, ts=ts, light_code=code,
                                view_of_ray=view)
    assert np.array_equal(_bits(rgb), _bits(ro)), np.abs(rgb - ro).max()
    ref = g[tag + "_rgb"].reshape(-1, 3)
    assert np.abs(rgb - ref).max() < 1e-4
    assert helpers.psnr(rgb, ref) > 80


@pytest.mark.parametrize("S,R", [(32, 70), (128, 9), (192, 5), (64, 1)])
def test_nerfle_render_other_sample_counts(S, R):
    """Ragged ray counts and samples/ray that are a divisor / multiple of the 64-sample tile."""
    from neural_raytracing_b200 import ops
    w1, w2 = helpers.nerfle_weights(False)
    rays = synth.camera_rays(77, R)
    ts = np.linspace(0.05, 2.0, S).astype(np.float32)
    code = np.array([[0.4, 1.0, 0.3]], np.float32)
    rgb = ops.nerfle_render(helpers.cuda_mlp(w1), helpers.cuda_mlp(w2), _t(rays), _t(ts), _t(code)).cpu().numpy()
    ro = c_oracle.nerfle_render(helpers.oracle_mlp(w1), helpers.oracle_mlp(w2), rays, ts=ts, light_code=code)
    assert np.array_equal(_bits(rgb), _bits(ro)), np.abs(rgb - ro).max()
