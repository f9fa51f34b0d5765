"""Supplementary measurements of the other BASELINE.json configs through the drop-in package
(not the bench.py contract; results go to profiles/).  CUDA-event timed, 3 warm-ups."""
import json, os, random, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests")); sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
import numpy as np, torch
import scenes, synth
import neural_raytracing_b200.pathtracer as P
from neural_raytracing_b200 import config, ops
from neural_raytracing_b200.pathtracer.cameras import NeRFCamera
from neural_raytracing_b200.pathtracer.shapes.nerf import NeRFLE
from neural_raytracing_b200.pathtracer.lights import PointLights
from neural_raytracing_b200.pathtracer.utils import eikonal_loss

random.random = lambda: 0.37
FLOP_SDF = 331008


def timed(fn, n=5, warm=3):
    for _ in range(warm): fn()
    torch.cuda.synchronize()
    e = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(n)]
    for a, b in e:
        a.record(); fn(); b.record()
    torch.cuda.synchronize()
    return sum(a.elapsed_time(b) for a, b in e) / n


out = {}
# ---- cfg1: colocate-style forward render, 64x64 (and 512x512) ----
for size, chunk in ((64, 64), (512, 128)):
    shape, sphere, bsdf, lights, integrator, w_isect = scenes.build_pipeline(P, "colocate", device="cuda")
    c2w, focal = synth.nerf_cameras(1, size, device="cuda")
    cam = NeRFCamera(cam_to_world=c2w, focal=focal, device="cuda")
    for prec in ("f32", "f16"):
        config.set_precision(prec)
        def render():
            with torch.no_grad():
                return P.pathtrace(shape, size=size, chunk_size=chunk, bundle_size=1, bsdf=bsdf, integrator=integrator, lights=lights,
                                   cameras=cam, device="cuda", silent=True, background=0, w_isect=w_isect, with_noise=False)
        ops.profile_collect(); ops.profile_enable(True)
        ms = timed(render, n=3, warm=2)
        prof = ops.profile_collect(); ops.profile_enable(False)
        R = size * size
        out["cfg1_colocate_%dx%d_%s" % (size, size, prec)] = {
            "ms_per_frame": ms, "rays_per_sec": R / ms * 1e3,
            "reference_executed_sdf_samples_per_sec": R * (64 + 130 + 64) / ms * 1e3,
            "kernel_ms_per_frame": {k: round(v[0] / 5, 3) for k, v in prof.items() if v[1]},
            "note": "prec=f32: everything on the fp32 exact kernels; prec=f16: march/scan/shadow + NeuralBSDF + occ MLPs on tcgen05, sp_var (16x256) still fp32"}
config.set_precision("f32")
# ---- cfg3: nerfle.py training step, 4 views x 32x32 crop = 4096 rays, S = 64, fwd + bwd + AdamW ----
n = NeRFLE(device="cuda"); synth.fill_module(n, 3)
with torch.no_grad():
    n.first.out.bias[0] = 0.8     # positive density: non-trivial gradients
opt = torch.optim.AdamW(n.parameters(), lr=8e-5, weight_decay=0)
rays = torch.from_numpy(synth.camera_rays(5, 4 * 32 * 32).reshape(4, 32, 32, 1, 6)).cuda()
lights = PointLights(device="cuda", location=torch.randn(4, 3, device="cuda"), scale=10)
target = torch.full((4, 32, 32, 1, 3), 0.5, device="cuda")
def step():
    opt.zero_grad()
    loss = torch.nn.functional.mse_loss(n(rays, lights), target)
    loss.backward(); opt.step()
for tprec in ("f32", "f16"):
    config.set_train_precision(tprec)
    ops.profile_collect(); ops.profile_enable(True)
    ms = timed(step)
    prof = ops.profile_collect(); ops.profile_enable(False)
    out["cfg3_nerfle_train_4096rays_" + tprec] = {
        "ms_per_step": ms, "rays_per_sec": 4096 / ms * 1e3, "mlp_samples_per_sec": 4096 * 64 / ms * 1e3,
        "model_tflops_fwd_bwd": 4096 * 64 * 325504 * 3 / ms / 1e9,
        "kernel_ms_per_step": {k: round(v[0] / 8, 3) for k, v in prof.items() if v[1]},
        "note": "train_precision=%s: fused MLP forward (saves activations) + fused backward + CUDA compositing fwd/bwd + torch AdamW" % tprec}
# same at the 65,536-ray batch of cfg5
rays2 = torch.from_numpy(synth.camera_rays(6, 4 * 128 * 128).reshape(4, 128, 128, 1, 6)).cuda()
target2 = torch.full((4, 128, 128, 1, 3), 0.5, device="cuda")
def step2():
    opt.zero_grad()
    loss = torch.nn.functional.mse_loss(n(rays2, lights), target2)
    loss.backward(); opt.step()
for tprec in ("f32", "f16"):
    config.set_train_precision(tprec)
    ops.profile_collect(); ops.profile_enable(True)
    ms = timed(step2, n=3, warm=2)
    prof = ops.profile_collect(); ops.profile_enable(False)
    out["cfg5_nerfle_train_65536rays_1gpu_" + tprec] = {
        "ms_per_step": ms, "rays_per_sec": 65536 / ms * 1e3, "model_tflops_fwd_bwd": 65536 * 64 * 325504 * 3 / ms / 1e9,
        "kernel_ms_per_step": {k: round(v[0] / 5, 3) for k, v in prof.items() if v[1]}}
config.set_train_precision("f32")
# ---- cfg4: DTU-style training step on a crop (SDF + 3-basis BSDF + LightField, eikonal + BCE) ----
shape, sphere, bsdf, lights, integrator, w_isect = scenes.build_pipeline(P, "dtu", device="cuda")
# the BSDF of dtu.py:101-106: 10 NeuralBSDF (Sigmoid) + 6 Diffuse(sigmoid) bases under a 16-way spatially varying blend
import torch.nn as nn
kids = [P.bsdf.NeuralBSDF(activation=nn.Sigmoid(), device="cuda") for _ in range(10)] + \
       [P.bsdf.Diffuse(preprocess=torch.sigmoid, device="cuda").random() for _ in range(6)]
for i, k in enumerate(kids[:10]):
    synth.fill_module(k, 70 + i)
bsdf = P.bsdf.ComposeSpatialVarying(kids, device="cuda")
bsdf.sp_var_fn._synth_sigma = 128.0
synth.fill_module(bsdf.sp_var_fn, 64)
params = list(sphere.parameters()) + list(bsdf.parameters()) + list(lights.parameters())
opt2 = torch.optim.AdamW(params, lr=8e-5, weight_decay=0)
c2w, focal = synth.nerf_cameras(1, 512, device="cuda")
cam = NeRFCamera(cam_to_world=c2w, focal=focal, device="cuda")
for size, crop, uv in ((512, 128, (190, 200)), (512, 512, (0, 0))):
    def dtu_step():
        opt2.zero_grad()
        got, mi = P.pathtrace_sample(shape, size=size, chunk_size=size, bundle_size=1, crop_size=crop, uv=uv, bsdf=bsdf,
                                     integrator=integrator, lights=lights, cameras=cam, device="cuda", silent=True, background=0,
                                     w_isect=w_isect, with_noise=False, addition=lambda it: it, squeeze_first=False)
        loss = (got[..., :3] - 0.5).square().mean() + 0.1 * eikonal_loss(mi.raw_normals)
        loss.backward(); opt2.step()
    for prec in ("f32", "f16"):
        if prec == "f32" and crop > 128:
            continue
        config.set_precision(prec)
        config.set_train_precision(prec)      # f16: the 256-wide nets' training forward on the tensor cores
        torch.cuda.reset_peak_memory_stats()
        ops.profile_collect(); ops.profile_enable(True)
        ms = timed(dtu_step, n=3, warm=2)
        prof = ops.profile_collect(); ops.profile_enable(False)
        out["cfg4_dtu_style_step_%dx%dcrop_%s" % (crop, crop, prec)] = {
            "ms_per_step": ms, "rays_per_sec": crop * crop / ms * 1e3, "peak_mem_gb": round(torch.cuda.max_memory_allocated() / 2**30, 2),
            "kernel_ms_per_step": {k: round(v[0] / 5, 3) for k, v in prof.items() if v[1]},
            "note": "%d-ray crop, dtu.py BSDF (10 NeuralBSDF + 6 Diffuse, 16-way sp_var), LightField, fwd + bwd + AdamW; gradient-free march + min-scan in `prec`; MLP backward fused fp32 kernels (f16: the 256-wide nets' training forward on tcgen05); normals through the analytic-Jacobian kernels (forward mode + hand-written reverse pass)" % (crop * crop)}
config.set_precision("f32")
config.set_train_precision("f32")
print(json.dumps(out, indent=1))
