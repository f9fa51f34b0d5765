import os, sys, random
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests")); sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
import helpers, synth
from neural_raytracing_b200 import config, ops
from neural_raytracing_b200.pathtracer.lights import PointLights
from neural_raytracing_b200.pathtracer.shapes.nerf import NeRFLE
n = NeRFLE(device="cuda"); synth.fill_module(n, 3)
with torch.no_grad():
    n.first.out.bias[0] = 0.8
rays = torch.from_numpy(synth.camera_rays(5, 4 * 16 * 16).reshape(4, 16, 16, 1, 6)).cuda()
lights = PointLights(device="cuda", location=torch.randn(4, 3, device="cuda"), scale=10)
target = torch.full((4, 16, 16, 1, 3), 0.5, device="cuda")
res = {}
for prec in ("f32", "f16"):
    config.set_train_precision(prec)
    random.seed(0)
    n.zero_grad()
    loss = torch.nn.functional.mse_loss(n(rays, lights), target)
    loss.backward()
    res[prec] = [p.grad.clone() for p in n.parameters()]
    print(prec, "loss", float(loss.detach()))
names = [k for k, _ in n.named_parameters()]
for k, a, b in zip(names, res["f32"], res["f16"]):
    c = float((a.double().flatten() @ b.double().flatten()) / (a.double().norm() * b.double().norm() + 1e-300))
    print("%-24s |f32| %.3e |f16| %.3e finite %s cos %.6f" % (k, float(a.norm()), float(b.norm()), bool(torch.isfinite(b).all()), c))
