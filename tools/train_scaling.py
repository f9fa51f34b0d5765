"""Development: NeRFLE tensor-core training step vs batch size, with the per-kernel split."""
import os, sys, random
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests")); sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
import synth
from neural_raytracing_b200 import config, ops
from neural_raytracing_b200.pathtracer.lights import PointLights
from neural_raytracing_b200.pathtracer.shapes.nerf import NeRFLE
config.set_train_precision("f16")
n = NeRFLE(device="cuda"); synth.fill_module(n, 3)
with torch.no_grad():
    n.first.out.bias[0] = 0.8
opt = torch.optim.AdamW(n.parameters(), lr=8e-5, weight_decay=0)
lights = PointLights(device="cuda", location=torch.randn(1, 3, device="cuda"), scale=10)
for R in (4096, 16384, 32768, 65536, 131072):
    rays = torch.from_numpy(synth.camera_rays(6, R).reshape(1, R, 1, 1, 6)).cuda()
    target = torch.full((1, R, 1, 1, 3), 0.5, device="cuda")
    def step():
        opt.zero_grad()
        loss = torch.nn.functional.mse_loss(n(rays, lights), target)
        loss.backward(); opt.step()
    for _ in range(3): step()
    torch.cuda.synchronize()
    ops.profile_collect(); ops.profile_enable(True)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(3): step()
    e1.record(); torch.cuda.synchronize()
    prof = ops.profile_collect(); ops.profile_enable(False)
    ms = e0.elapsed_time(e1) / 3
    ks = {k: v[0] / 3 for k, v in prof.items() if v[1]}
    print("R=%6d: %.2f ms/step (%.2f us per 1000 samples) kernels: %s | sum %.2f | peak mem %.1f GB" % (
        R, ms, ms * 1e3 / (R * 64 / 1000), ", ".join("%s %.2f" % kv for kv in ks.items()), sum(ks.values()), torch.cuda.max_memory_allocated() / 2**30))
