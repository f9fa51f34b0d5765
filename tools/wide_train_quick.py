"""Development: the 256-wide tensor-core training kernels: per-tensor gradient cosines vs float64 autograd and kernel times.
usage: wide_train_quick.py [sp_var4|sp_var16|light_field]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests")); sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
import torch, helpers, synth
import test_gpu_tc_train as T
from neural_raytracing_b200 import ops
name = sys.argv[1] if len(sys.argv) > 1 else "sp_var16"
kw = T.WIDE_KW[name]
w = synth.mlp_weights(**kw); m = helpers.cuda_mlp(w)
for M in (1, 129, 5000):
    g = torch.Generator(device="cuda").manual_seed(M + 5)
    x = 0.6 * torch.randn(M, 3, device="cuda", generator=g); gy = torch.randn(M, kw["out"], device="cuda", generator=g) * 3e-4
    out, ws = ops.mlp_forward_train_tc(m, x, 0, prec="f16")
    gp, gx = ops.mlp_backward_tc(m, M, out, gy, ws, 0, need_input_grad=True, prec="f16")
    gW, gb = m.unpack(gp)
    for quantised in (True, False):
        y, xr, Ws, bs = T._ref(w, x, False, quantised, True)
        (y * gy.double()).sum().backward()
        print(name, "M", M, "quantised" if quantised else "exact", "fwd err %.2e" % float((out.double() - y.detach()).abs().max()))
        print("   W cos", " ".join("%.5f" % T._cos(a, r.grad) for a, r in zip(gW, Ws)))
        print("   b cos", " ".join("%.5f" % T._cos(a, r.grad) for a, r in zip(gb, bs)))
        print("   gx cos %.6f  |gx| %.3e ref %.3e" % (T._cos(gx, xr.grad), float(gx.norm()), float(xr.grad.norm())))
for M in (262144,):
    x = 0.6 * torch.randn(M, 3, device="cuda"); gy = torch.randn(M, kw["out"], device="cuda") * 3e-4
    ops.profile_collect(); ops.profile_enable(True)
    for _ in range(3):
        out, ws = ops.mlp_forward_train_tc(m, x, 0, prec="f16")
        gp, gx = ops.mlp_backward_tc(m, M, out, gy, ws, 0, need_input_grad=True, prec="f16")
    torch.cuda.synchronize()
    pr = ops.profile_collect(); ops.profile_enable(False)
    print("M", M, "ws GB %.2f" % (ws.numel() / 2**30), {k: (round(v[0] / 3, 3), v[1] // 3) for k, v in pr.items() if v[1]})
