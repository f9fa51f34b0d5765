"""Development: run the fused tensor-core MLP kernel a few times (target for ncu)."""
import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests")); sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
import torch, helpers, synth
from neural_raytracing_b200 import ops
which = sys.argv[1] if len(sys.argv) > 1 else "nerf_first"
kw, act = helpers.MLP_CASES[which]
m = helpers.cuda_mlp(synth.mlp_weights(**kw), act)
M = 148 * 2 * 128 * 24
x = torch.randn(M, kw["in_size"], device="cuda") * 0.5
for _ in range(3):
    ops.mlp_forward(m, x, prec="f16")
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(); ops.mlp_forward(m, x, prec="f16"); e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1)
flop = 2 * sum(k * n for k, n in m.dims) * M
print("%s M=%d: %.3f ms -> %.1f TFLOP/s" % (which, M, ms, flop / ms / 1e9))
