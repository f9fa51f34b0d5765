import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests")); sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
import numpy as np, torch, copy
import synth
from neural_raytracing_b200.pathtracer import neural_blocks as nb
torch.manual_seed(0)
mlp = nb.SkipConnMLP(device="cuda", in_size=3, out=65, num_layers=5, hidden_size=128, freqs=16).to("cuda")
synth.fill_module(mlp, 7)
for M in (64, 65, 200, 256):
    g = torch.Generator("cuda").manual_seed(M)
    x = (0.5 * torch.randn(M, 3, device="cuda", generator=g))
    go = torch.randn(M, 65, device="cuda", generator=g)
    def run(fused):
        nb._FUSED_BACKWARD[0] = fused
        for p in mlp.parameters(): p.grad = None
        xx = x.clone().requires_grad_()
        y = mlp(xx, None)
        (y * go).sum().backward()
        return mlp.init.weight.grad.clone().cpu().double(), xx.grad.clone().cpu().double()
    g1, gx1 = run(True); g0, gx0 = run(False)
    # float64 CPU reference
    m64 = copy.deepcopy(mlp).cpu().double(); m64.basis_p = mlp.basis_p.cpu().double()
    x64 = x.cpu().double().requires_grad_()
    y64 = m64.forward_reference_ops(x64, None)
    (y64 * go.cpu().double()).sum().backward()
    r = m64.init.weight.grad
    print("M=%d init.weight grad: fused err max %.3e (x cols %.3e, rest %.3e) | torch-fp32 err max %.3e (x cols %.3e) | gx fused %.3e torch %.3e" % (
        M, (g1 - r).abs().max(), (g1 - r)[:, :3].abs().max(), (g1 - r)[:, 3:].abs().max(), (g0 - r).abs().max(), (g0 - r)[:, :3].abs().max(),
        (gx1 - x64.grad).abs().max(), (gx0 - x64.grad).abs().max()))
