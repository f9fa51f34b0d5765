"""Development: where does a DTU-style training step (cfg4) spend its time?  torch profiler, GPU time by kernel."""
import os, sys, random
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests")); sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
import scenes, synth
import neural_raytracing_b200.pathtracer as P
from neural_raytracing_b200 import config
from neural_raytracing_b200.pathtracer.cameras import NeRFCamera
from neural_raytracing_b200.pathtracer.utils import eikonal_loss
random.random = lambda: 0.37
config.set_precision("f16")
shape, sphere, bsdf, lights, integrator, w_isect = scenes.build_pipeline(P, "dtu", device="cuda")
params = list(sphere.parameters()) + list(bsdf.parameters()) + list(lights.parameters())
opt = torch.optim.AdamW(params, lr=8e-5, weight_decay=0)
size, crop = 512, 128
c2w, focal = synth.nerf_cameras(1, size, device="cuda")
cam = NeRFCamera(cam_to_world=c2w, focal=focal, device="cuda")
def step():
    opt.zero_grad()
    got, mi = P.pathtrace_sample(shape, size=size, chunk_size=crop, bundle_size=1, crop_size=crop, uv=(190, 200), bsdf=bsdf,
                                 integrator=integrator, lights=lights, cameras=cam, device="cuda", silent=True, background=0,
                                 w_isect=w_isect, with_noise=False, addition=lambda it: it, squeeze_first=False)
    loss = (got[..., :3] - 0.5).square().mean() + 0.1 * eikonal_loss(mi.raw_normals)
    loss.backward(); opt.step()
for _ in range(3): step()
torch.cuda.synchronize()
import time
t0 = time.perf_counter()
for _ in range(5): step()
torch.cuda.synchronize()
print("wall ms/step", (time.perf_counter() - t0) / 5 * 1e3)
from torch.profiler import profile, ProfilerActivity
with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as prof:
    for _ in range(3): step()
    torch.cuda.synchronize()
ka = prof.key_averages()
rows = sorted([(e.self_device_time_total / 3e3, e.count // 3, e.key[:90]) for e in ka if e.self_device_time_total > 0], reverse=True)
tot = sum(r[0] for r in rows)
print("GPU busy ms/step %.2f in %d launches" % (tot, sum(r[1] for r in rows)))
for r in rows[:28]: print("%8.3f ms %5d x  %s" % r)
