"""Development check of the tensor-core training path (forward-with-save, dgrad, wgrad) against float64 autograd."""
import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests")); sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
import helpers, synth
from neural_raytracing_b200 import ops


def ref_mlp(w, x, out_act):
    """float64 torch restatement of neural_blocks.py:75-86 with autograd."""
    Ws = [torch.tensor(a, dtype=torch.float64, device="cuda", requires_grad=True) for a in w["W"]]
    bs = [torch.tensor(a, dtype=torch.float64, device="cuda", requires_grad=True) for a in w["b"]]
    B = torch.tensor(w["basis"], dtype=torch.float64, device="cuda")
    x = x.double().requires_grad_()
    ph = x @ B
    enc = torch.cat([x, ph.sin(), ph.cos()], -1)
    act = torch.nn.functional.leaky_relu
    h = enc @ Ws[0].t() + bs[0]
    L = w["num_layers"]
    for i in range(L):
        if i != L - 1 and i % w["skip"] == 0:
            h = torch.cat([h, enc], -1)
        h = act(h) @ Ws[1 + i].t() + bs[1 + i]
    y = act(h) @ Ws[-1].t() + bs[-1]
    if out_act == ops.OUT_SIGMOID:
        y = y.sigmoid()
    return y, x, Ws, bs


def cos(a, b):
    a, b = a.double().flatten(), b.double().flatten()
    return float((a @ b) / (a.norm() * b.norm() + 1e-300))


for name, out_act, need_x in (("nerf_first", ops.OUT_NONE, False), ("nerf_second", ops.OUT_SIGMOID, True), ("nerf_second", ops.OUT_SIGMOID, False)):
    kw, _ = helpers.MLP_CASES[name]
    w = synth.mlp_weights(**kw)
    m = helpers.cuda_mlp(w)
    for M, prec in ((1000, "f16"), (128 * 148 * 6 + 77, "f16"), (128 * 148 * 6 + 77, "bf16")):
        g = torch.Generator(device="cuda").manual_seed(M)
        scale = 0.6 if name == "nerf_first" else 0.1
        x = scale * torch.randn(M, kw["in_size"], device="cuda", generator=g)
        gy = torch.randn(M, kw["out"], device="cuda", generator=g) * 1e-3
        out, ws = ops.mlp_forward_train_tc(m, x, out_act, prec=prec)
        gp, gx = ops.mlp_backward_tc(m, M, out, gy, ws, out_act, need_input_grad=need_x, prec=prec)
        torch.cuda.synchronize()
        y, xr, Ws, bs = ref_mlp(w, x, out_act)
        (y * gy.double()).sum().backward()
        print("%s M=%d need_x=%s %s: fwd max err %.3e" % (name, M, need_x, prec, float((out.double() - y.detach()).abs().max())))
        gW, gb = m.unpack(gp)
        worst = 1.0
        for i, (a, b, ra, rb) in enumerate(zip(gW, gb, Ws, bs)):
            cw, cb = cos(a, ra.grad), cos(b, rb.grad)
            worst = min(worst, cw, cb)
            print("   linear %2d: cos(gW) %.6f  cos(gb) %.6f   |gW| %.3e (ref %.3e)  max abs err W %.2e b %.2e" % (
                i, cw, cb, float(a.norm()), float(ra.grad.norm()), float((a.double() - ra.grad).abs().max()), float((b.double() - rb.grad).abs().max())))
        if need_x:
            print("   g_x: cos %.6f  max abs err %.3e (|ref| max %.3e)" % (cos(gx, xr.grad), float((gx.double() - xr.grad).abs().max()), float(xr.grad.abs().max())))
        print("   worst cosine %.6f" % worst)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(3):
            out, ws = ops.mlp_forward_train_tc(m, x, out_act, prec=prec)
            gp, gx = ops.mlp_backward_tc(m, M, out, gy, ws, out_act, need_input_grad=need_x, prec=prec)
        e1.record(); torch.cuda.synchronize()
        ops.profile_collect(); ops.profile_enable(True)
        out, ws = ops.mlp_forward_train_tc(m, x, out_act, prec=prec)
        gp, gx = ops.mlp_backward_tc(m, M, out, gy, ws, out_act, need_input_grad=need_x, prec=prec)
        prof = ops.profile_collect(); ops.profile_enable(False)
        print("   kernels (ms): " + ", ".join("%s %.3f" % (k, v[0]) for k, v in prof.items() if v[1]))
        flop = 3 * 2 * sum(k * n for k, n in m.dims) * M
        print("   fwd+bwd %.3f ms -> %.1f TFLOP/s (3x forward FLOP)" % (e0.elapsed_time(e1) / 3, flop / (e0.elapsed_time(e1) / 3) / 1e9))
