"""Development check of the tensor-core march / shadow / min-scan against the fp32 exact kernels
(which are bit-identical to the oracle).  Run on the GPU box: python tools/tc_march_check.py"""
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
import helpers  # noqa: E402
import synth  # noqa: E402
from neural_raytracing_b200 import ops  # noqa: E402


def T(a):
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


def timed(fn, n=3):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        out = fn()
    e1.record(); torch.cuda.synchronize()
    return out, e0.elapsed_time(e1) / n


w = helpers.golden_sdf_weights()
sdf = helpers.cuda_sdf(w)
for R in (1, 100, 4096, 262144):
    rays = T(synth.camera_rays(3, R))
    cnt32 = torch.zeros(1, dtype=torch.int64, device="cuda")
    (d32, h32), ms32 = timed(lambda: ops.sphere_trace(sdf, rays, 1e-3, 64, 10.0, prec="f32", steps_counter=cnt32), 1)
    for prec in ("f16", "bf16"):
        cnt = torch.zeros(1, dtype=torch.int64, device="cuda")
        (d, h), ms = timed(lambda: ops.sphere_trace(sdf, rays, 1e-3, 64, 10.0, prec=prec, steps_counter=cnt), 1)
        both = (h & h32)
        xor = int((h ^ h32).sum())
        derr = float((d - d32)[both].abs().max()) if both.any() else 0.0
        derr_all = float((d - d32).abs().max())
        if both.any():
            q = torch.quantile((d - d32)[both].abs()[:1000000], torch.tensor([0.5, 0.9, 0.99, 0.999], device="cuda"))
            print("   depth err quantiles (both hit) 50/90/99/99.9%%: %s" % " ".join("%.2e" % float(v) for v in q))
        steps = int(cnt.item()) // 2
        print("march R=%7d %-4s: hits %d/%d xor %d  depth err (both hit) %.2e (all) %.2e | %.3f ms (f32 %.3f ms) steps %d -> %.1f Msamples/s %.1f TFLOP/s"
              % (R, prec, int(h.sum()), int(h32.sum()), xor, derr, derr_all, ms, ms32, steps, steps / ms / 1e3,
                 steps * 331008 / ms / 1e9))
    # shadow rays: from the hit points towards a light
    p = rays[:, :3] + d32[:, None] * rays[:, 3:]
    light = torch.tensor([0.3, 1.2, 0.4], device="cuda")
    dirv = light - p
    dist = dirv.norm(dim=-1)
    srays = torch.cat([p, dirv / dist[:, None]], -1).contiguous()
    nb32, ms32 = timed(lambda: ops.shadow_test(sdf, srays, dist, 1e-3, 64, prec="f32"), 1)
    nb16, ms16 = timed(lambda: ops.shadow_test(sdf, srays, dist, 1e-3, 64, prec="f16"), 1)
    print("shadow R=%7d f16: not_blocked %d/%d xor %d | %.3f ms (f32 %.3f ms)" % (R, int(nb16.sum()), int(nb32.sum()), int((nb16 ^ nb32).sum()), ms16, ms32))
    step = (2.2 + 0.37 * 2 / 128) / 128
    (i32, p32, m32), ms32 = timed(lambda: ops.min_scan(sdf, rays, step, 128, prec="f32"), 1)
    (i16, p16, m16), ms16 = timed(lambda: ops.min_scan(sdf, rays, step, 128, prec="f16"), 1)
    print("minscan R=%7d f16: idx equal %.4f  |idx diff| max %d  min_val err %.2e | %.3f ms (f32 %.3f ms) -> %.1f TFLOP/s"
          % (R, float((i16 == i32).float().mean()), int((i16 - i32).abs().max()), float((m16 - m32).abs().max()), ms16, ms32,
             R * 129 * 331008 / ms16 / 1e9))
