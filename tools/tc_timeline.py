"""Development: dump the clock64 timeline of CTA 0 of the fused tensor-core MLP kernel."""
import ctypes, sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests")); sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
import numpy as np, torch
import helpers, synth
from neural_raytracing_b200 import ops, _native as N
which = sys.argv[1] if len(sys.argv) > 1 else "nerf_first"
kw, act = helpers.MLP_CASES[which]
m = helpers.cuda_mlp(synth.mlp_weights(**kw), act)
STAGES = kw["num_layers"] + 3
M = 148 * 2 * 128 * 12
x = torch.randn(M, kw["in_size"], device="cuda") * 0.5
ops.mlp_forward(m, x, prec="f16"); torch.cuda.synchronize()
buf = torch.zeros(4 * STAGES * 2 * 8, dtype=torch.int64, device="cuda")
N.lib().nrtdbg_set_timeline(ctypes.c_void_p(buf.data_ptr()))
ops.mlp_forward(m, x, prec="f16"); torch.cuda.synchronize()
N.lib().nrtdbg_set_timeline(None)
t = buf.cpu().numpy().reshape(4, STAGES, 2, 8)
t0 = t[t > 0].min()
names = ["mma:wait_ready", "mma:ready", "mma:committed", "epi:wait_done", "epi:done", "epi:loaded", "epi:stored", "epi:arrived"]
for it in range(2):
    for st in range(STAGES):
        for slot in range(2):
            row = t[it, st, slot]
            print("it%d st%d slot%d " % (it, st, slot) + "  ".join("%s=%6d" % (names[k].split(":")[1], row[k] - t0) if row[k] > 0 else "%s=     -" % names[k].split(":")[1] for k in range(8)))
per_tilepair = (t[1:, 0, 0, 1] - t[:-1, 0, 0, 1])
print("cycles per tile-pair iteration:", per_tilepair)
