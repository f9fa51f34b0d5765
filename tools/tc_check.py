"""Development check of the tensor-core path against the CPU oracle (prints errors)."""
import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests")); sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
import numpy as np, torch
import helpers, synth
from oracle import c_oracle
from neural_raytracing_b200 import ops

def T(a): return torch.from_numpy(np.ascontiguousarray(a)).cuda()
cases = dict(helpers.MLP_CASES)
cases["occ"] = (dict(seed=18, in_size=5, out=1, num_layers=8, hidden=64, freqs=16, sigma=32.0), None)
for name in ("nerf_first", "nerf_second", "neural_bsdf", "occ"):
    kw, act = cases[name]
    w = synth.mlp_weights(**kw)
    rs = np.random.RandomState(3)
    for M in (1000, 128, 1):
        x = (0.6 * rs.standard_normal((M, kw["in_size"]))).astype(np.float32)
        yo = c_oracle.mlp_forward(helpers.oracle_mlp(w, act), x)
        for prec in ("f16", "bf16"):
            try:
                y = ops.mlp_forward(helpers.cuda_mlp(w, act), T(x), prec=prec).cpu().numpy()
                torch.cuda.synchronize()
                print("%-12s M=%4d %-4s max_abs_err %.3e (|y| max %.3f) finite=%s" % (name, M, prec, np.abs(y - yo).max(), np.abs(yo).max(), np.isfinite(y).all()))
            except Exception as e:
                print(name, M, prec, "FAILED", e)
# SDF residual MLP (softplus, weights streamed) and the full SDF on the tensor cores
wsdf = helpers.golden_sdf_weights()
rs = np.random.RandomState(5)
pts = (0.5 * rs.standard_normal((3000, 3))).astype(np.float32)
yo = c_oracle.mlp_forward(helpers.oracle_mlp(wsdf["shift"], "softplus"), pts)
vo = c_oracle.sdf_eval(helpers.oracle_sdf(wsdf), pts)
for prec in ("f16", "bf16"):
    y = ops.mlp_forward(helpers.cuda_mlp(wsdf["shift"], "softplus"), T(pts), prec=prec).cpu().numpy()
    v = ops.sdf_eval(helpers.cuda_sdf(wsdf), T(pts), prec=prec).cpu().numpy()
    print("sdf_shift %-4s max_abs_err %.3e (|y| max %.3f) | sdf_eval max_abs_err %.3e (|v| max %.3f)" % (prec, np.abs(y - yo).max(), np.abs(yo).max(), np.abs(v - vo).max(), np.abs(vo).max()))
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
big = T((0.5 * rs.standard_normal((148 * 256 * 16, 3))).astype(np.float32)); sdf_c = helpers.cuda_sdf(wsdf)
ops.sdf_eval(sdf_c, big, prec="f16"); torch.cuda.synchronize()
e0.record(); ops.sdf_eval(sdf_c, big, prec="f16"); e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1); print("sdf_eval f16 %d pts: %.3f ms -> %.1f Msamples/s, %.1f TFLOP/s" % (big.shape[0], ms, big.shape[0] / ms / 1e3, big.shape[0] * 331008 / ms / 1e9))
e0.record(); ops.sdf_eval(sdf_c, big, prec="f32"); e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1); print("sdf_eval f32 %d pts: %.3f ms -> %.1f Msamples/s, %.1f TFLOP/s" % (big.shape[0], ms, big.shape[0] / ms / 1e3, big.shape[0] * 331008 / ms / 1e9))
g = helpers.golden("nerfle")
w1, w2 = helpers.nerfle_weights(False)
rays = synth.camera_rays(33, 500)
ts = helpers.nerfle_ts(g["fixed_random"])
code = g["pt_light_loc"][:1]
ro = c_oracle.nerfle_render(helpers.oracle_mlp(w1), helpers.oracle_mlp(w2), rays, ts=ts, light_code=code)
for prec in ("f16", "bf16"):
    rgb = ops.nerfle_render(helpers.cuda_mlp(w1), helpers.cuda_mlp(w2), T(rays), T(ts), T(code), prec=prec).cpu().numpy()
    print("nerfle render %s: max_abs %.3e psnr %.1f dB" % (prec, np.abs(rgb - ro).max(), helpers.psnr(rgb, ro)))
rgb32 = ops.nerfle_render(helpers.cuda_mlp(w1), helpers.cuda_mlp(w2), T(rays), None, T(code), prec="f32", n_coarse=64, n_fine=128, t_near=0.0, t_far=2.05, jitter_seed=0).cpu().numpy()
rgb16 = ops.nerfle_render(helpers.cuda_mlp(w1), helpers.cuda_mlp(w2), T(rays), None, T(code), prec="f16", n_coarse=64, n_fine=128, t_near=0.0, t_far=2.05, jitter_seed=0).cpu().numpy()
print("hierarchical 64+128 f16 vs f32: max_abs %.3e psnr %.1f" % (np.abs(rgb16 - rgb32).max(), helpers.psnr(rgb16, rgb32)))
