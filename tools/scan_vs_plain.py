"""Development: the min scan (IoScanEval: points from rays, sphere set per sample) against the plain forward of the same
residual MLP on the same number of samples, and against the point evaluation of the whole SDF (IoSdfEval)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests")); sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
import numpy as np, torch
import helpers, synth
from neural_raytracing_b200 import ops
w = helpers.golden_sdf_weights()
sdf = helpers.cuda_sdf(w)
mlp = helpers.cuda_mlp(w["shift"], "softplus")
R = 262144
rays = torch.from_numpy(synth.camera_rays(3, R)).cuda()
M = R * 129
x = (0.5 * torch.randn(M, 3, device="cuda"))
def timed(fn, n=3):
    fn(); torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(n): fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / n
step = (2.2 + 0.37 * 2 / 128) / 128
print("min scan (R=%d x 129)      %.2f ms" % (R, timed(lambda: ops.min_scan(sdf, rays, step, 128, prec="f16"))))
print("plain MLP forward, M=%d  %.2f ms" % (M, timed(lambda: ops.mlp_forward(mlp, x, prec="f16"))))
print("sdf_eval (MLP + sphere set)  %.2f ms" % timed(lambda: ops.sdf_eval(sdf, x, prec="f16")))
