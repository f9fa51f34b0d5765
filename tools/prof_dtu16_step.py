"""Development: where does a cfg4 training step (dtu.py's model: 10 NeuralBSDF + 6 Diffuse, 16-way sp_var, LightField,
DTUCamera, masked_loss without SSIM + eikonal, AdamW) spend its time?  usage: prof_dtu16_step.py [crop] [f16|f32]"""
import os, sys, random, time
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests")); sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
import scenes
import neural_raytracing_b200.pathtracer as P
from neural_raytracing_b200 import config, ops
from neural_raytracing_b200.pathtracer.cameras import DTUCamera
from neural_raytracing_b200.pathtracer.utils import eikonal_loss, masked_loss
crop = int(sys.argv[1]) if len(sys.argv) > 1 else 128
prec = sys.argv[2] if len(sys.argv) > 2 else "f16"
random.random = lambda: 0.37
config.set_precision(prec); config.set_train_precision(prec)
shape, sphere, bsdf, lights, integrator = scenes.build_dtu16(P, device="cuda")
params = list(sphere.parameters()) + list(bsdf.parameters()) + list(lights.parameters())
FLAT = os.environ.get("NRT_FLAT") == "1"    # flat parameter / gradient buffers + one fused AdamW kernel (training.FlatParameters)
if FLAT:
    from neural_raytracing_b200 import training
    mlps = [sphere.shift, bsdf.sp_var_fn, lights.light_field_approx] + [k.mlp for k in bsdf.bsdfs if hasattr(k, "mlp")]
    mlp_params = {id(q) for m in mlps for q in m.parameters()}
    flat = training.FlatParameters(mlps, [q for q in params if id(q) not in mlp_params])
    opt = torch.optim.AdamW([flat.param], lr=8e-5, weight_decay=0, fused=True)
    opt.zero_grad = lambda *a, **k: flat.zero_grad()
else:
    opt = torch.optim.AdamW([{"params": list(sphere.parameters())}, {"params": list(bsdf.parameters())}, {"params": list(lights.parameters())}],
                            lr=8e-5, weight_decay=0)
size = crop
pose, K = scenes.dtu_cameras(1, device="cuda")
cam = DTUCamera(pose=pose, intrinsic=K, device="cuda")
exp, mask = scenes.dtu_targets(1, crop, device="cuda")
def step():
    opt.zero_grad()
    got, mi = P.pathtrace_sample(shape, size=size, chunk_size=size, bundle_size=1, crop_size=crop, bsdf=bsdf, integrator=integrator,
                                 cameras=cam, lights=lights, device="cuda", uv=(0, 0), background=0, addition=lambda mi: mi,
                                 squeeze_first=False, silent=True)
    loss = masked_loss(got[..., :3], exp, mi.throughput.squeeze(-1), mask, mask_weight=10, with_logits=mi.with_logits, ssim_fn=None) \
        + eikonal_loss(mi.raw_normals)
    loss.backward(); opt.step()
    return loss
for _ in range(3): step()
torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(5): l = step()
torch.cuda.synchronize()
print("crop %d (%d rays) %s: wall %.2f ms/step, loss %.4f, peak mem %.1f GB" % (crop, crop * crop, prec, (time.perf_counter() - t0) / 5 * 1e3,
      float(l.detach()), torch.cuda.max_memory_allocated() / 2**30))
ops.profile_collect(); ops.profile_enable(True)
for _ in range(3): step()
torch.cuda.synchronize()
pr = ops.profile_collect(); ops.profile_enable(False)
print("library kernels per step:", {k: (round(v[0] / 3, 3), v[1] // 3) for k, v in pr.items() if v[1]})
from torch.profiler import profile, ProfilerActivity
with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as prof:
    for _ in range(3): step()
    torch.cuda.synchronize()
ka = prof.key_averages()
rows = sorted([(e.self_device_time_total / 3e3, e.count // 3, e.key[:100]) for e in ka if e.self_device_time_total > 0], reverse=True)
tot = sum(r[0] for r in rows)
print("GPU busy ms/step %.2f in %d launches" % (tot, sum(r[1] for r in rows)))
for r in rows[:22]: print("%8.3f ms %5d x  %s" % r)
if os.environ.get("NRT_PROF_BY_COUNT"):
    print("--- by launch count")
    for r in sorted(rows, key=lambda r: -r[1])[:45]: print("%8.3f ms %5d x  %s" % r)
    cpu = sorted([(e.self_cpu_time_total / 3e3, e.count // 3, e.key[:90]) for e in ka], reverse=True)
    print("--- host side (self CPU ms per step)")
    for r in cpu[:40]: print("%8.3f ms %5d x  %s" % r)
