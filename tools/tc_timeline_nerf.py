"""Development: clock64 timeline of CTA 0 of the NeRFLE.second kernel with its real IO policy (IoNerfSecond) inside
nerfle_render (the first kernel's stamps are overwritten by the second's; slots 0 and 1 only)."""
import ctypes, sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests")); sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
import numpy as np, torch
import helpers, synth
from neural_raytracing_b200 import ops, _native as N
w1, w2 = helpers.nerfle_weights(False)
m1, m2 = helpers.cuda_mlp(w1), helpers.cuda_mlp(w2)
R = 148 * 3 * 2 * 12
rays = torch.from_numpy(synth.camera_rays(4, R)).cuda()
ts = torch.linspace(0, 2.037, 64, device="cuda")
code = torch.tensor([[0.4, 1.0, 0.3]], device="cuda")
ops.nerfle_render(m1, m2, rays, ts, code, prec="f16"); torch.cuda.synchronize()
STAGES = 11
buf = torch.zeros(4 * STAGES * 2 * 8, dtype=torch.int64, device="cuda")
N.lib().nrtdbg_set_timeline(ctypes.c_void_p(buf.data_ptr()))
ops.nerfle_render(m1, m2, rays, ts, code, prec="f16"); torch.cuda.synchronize()
N.lib().nrtdbg_set_timeline(None)
t = buf.cpu().numpy().reshape(4, STAGES, 2, 8)
t0 = t[:, :, :, 1:][t[:, :, :, 1:] > 0].min()
names = ["-", "ready", "committed", "wait_done", "done", "loaded", "stored", "arrived"]
for it in range(1, 3):
    for st in range(STAGES):
        for slot in range(2):
            row = t[it, st, slot]
            print("it%d st%d slot%d " % (it, st, slot) + "  ".join("%s=%6d" % (names[k], row[k] - t0) if row[k] > 0 else "%s=     -" % names[k] for k in range(1, 8)))
print("cycles per slot iteration:", t[1:, 0, 0, 1] - t[:-1, 0, 0, 1])
