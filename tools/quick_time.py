"""Development timing of the fp32 kernels (not the bench)."""
import sys, os, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests")); sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
import numpy as np, torch
import helpers, synth
from neural_raytracing_b200 import ops

def timeit(fn, n=3):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n

prec = sys.argv[1] if len(sys.argv) > 1 else "f32"
w = helpers.golden_sdf_weights(); s = helpers.cuda_sdf(w)
for R in (4096, 65536):
    rays = torch.from_numpy(synth.camera_rays(1, R)).cuda()
    steps = torch.zeros(1, dtype=torch.int64, device="cuda")
    ms = timeit(lambda: ops.sphere_trace(s, rays, 1e-3, 64, 10.0, steps_counter=steps))
    n = steps.item() / 4
    print("sphere_trace R=%d: %.2f ms  useful samples/launch %.0f -> %.1f Msamples/s, %.2f TFLOP/s (ref-executed %.1f Msamples/s)" % (R, ms, n, n / ms / 1e3, n * 331008 / ms / 1e9, R * 64 / ms / 1e3))
    step = 2.2 / 128
    ms = timeit(lambda: ops.min_scan(s, rays, step, 128))
    print("min_scan R=%d: %.2f ms -> %.1f Msamples/s %.2f TFLOP/s" % (R, ms, R * 129 / ms / 1e3, R * 129 * 331008 / ms / 1e9))
w1, w2 = helpers.nerfle_weights(False); m1, m2 = helpers.cuda_mlp(w1), helpers.cuda_mlp(w2)
ts = torch.linspace(0, 2.05, 64).cuda(); code = torch.tensor([[0.4, 1.0, 0.3]]).cuda()
for R in (4096, 65536):
    rays = torch.from_numpy(synth.camera_rays(2, R)).cuda()
    ms = timeit(lambda: ops.nerfle_render(m1, m2, rays, ts, code, prec=prec))
    print("nerfle %s R=%d S=64: %.2f ms -> %.2f Mrays/s, %.1f Msamples/s, %.2f TFLOP/s" % (prec, R, ms, R / ms / 1e3, R * 64 / ms / 1e3, R * 64 * 325504 / ms / 1e9))
rays = torch.from_numpy(synth.camera_rays(2, 65536)).cuda()
ms = timeit(lambda: ops.nerfle_render(m1, m2, rays, None, code, prec=prec, n_coarse=64, n_fine=128, t_near=0.0, t_far=2.05, jitter_seed=7))
print("nerfle %s hierarchical 64+128 R=65536: %.2f ms -> %.2f Mrays/s %.2f TFLOP/s" % (prec, ms, 65536 / ms / 1e3, 65536 * 192 * 325504 / ms / 1e9))
S, R = 64, 1 << 20
sg = torch.randn(S, R, device="cuda"); c = torch.rand(S, R, 3, device="cuda")
ms = timeit(lambda: ops.composite_forward(sg, c, ts), 5)
print("composite_fwd S=64 R=1M: %.3f ms -> %.0f GB/s" % (ms, (S * R * 16 + R * 12) / ms / 1e6))
go = torch.randn(R, 3, device="cuda")
ms = timeit(lambda: ops.composite_backward(sg, c, ts, go), 5)
print("composite_bwd S=64 R=1M: %.3f ms -> %.0f GB/s (algorithmic 32 B/sample)" % (ms, (S * R * 32 + R * 12) / ms / 1e6))
