import sys, os, copy
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests")); sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
import numpy as np, torch
import synth
from neural_raytracing_b200.pathtracer import neural_blocks as nb
torch.manual_seed(0)
mlp = nb.SkipConnMLP(device="cuda", in_size=70, out=3, num_layers=8, hidden_size=64, freqs=16).to("cuda")
synth.fill_module(mlp, 7)
M = 200
g = torch.Generator("cuda").manual_seed(M)
x = (0.5 * torch.randn(M, 70, device="cuda", generator=g))
go = torch.randn(M, 3, device="cuda", generator=g)
xx = x.clone().requires_grad_()
y = mlp(xx, None, out_act=1); (y * go).sum().backward()
def ref(dtype):
    m = copy.deepcopy(mlp).cpu().to(dtype); m.basis_p = mlp.basis_p.cpu().to(dtype)
    for p in m.parameters(): p.grad = None
    xr = x.cpu().to(dtype).requires_grad_()
    yr = m.forward_reference_ops(xr, None).sigmoid(); (yr * go.cpu().to(dtype)).sum().backward()
    return yr.detach().double(), xr.grad.double(), m.init.weight.grad.double(), m.layers[3].weight.grad.double()
y64, gx64, gi64, gl64 = ref(torch.float64)
y32, gx32, gi32, gl32 = ref(torch.float32)
def rel(a, b): return ((a.double().cpu() - b).abs().max() / b.abs().max()).item()
print("forward: fused %.2e  cpu-fp32 %.2e" % (rel(y.detach(), y64), rel(y32, y64)))
print("g_x    : fused %.2e  cpu-fp32 %.2e" % (rel(xx.grad, gx64), rel(gx32, gx64)))
print("g_init : fused %.2e  cpu-fp32 %.2e" % (rel(mlp.init.weight.grad, gi64), rel(gi32, gi64)))
print("g_l3   : fused %.2e  cpu-fp32 %.2e" % (rel(mlp.layers[3].weight.grad, gl64), rel(gl32, gl64)))
