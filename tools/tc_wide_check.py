"""Development check of the 256-wide tensor-core forward (sp_var 16x256, LightField 10x256) vs the C oracle."""
import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests")); sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
import helpers, synth
from neural_raytracing_b200 import ops
from oracle import c_oracle
CASES = {"sp_var": dict(seed=51, in_size=3, out=4, num_layers=16, hidden=256, freqs=128, sigma=128.0),
         "light_field": dict(seed=52, in_size=3, out=3, num_layers=10, hidden=256, freqs=16, sigma=32.0)}
T = lambda a: torch.from_numpy(np.ascontiguousarray(a)).cuda()
for name, kw in CASES.items():
    w = synth.mlp_weights(**kw)
    m = helpers.cuda_mlp(w)
    for M in (1, 129, 3000):
        x = (0.5 * np.random.RandomState(M).standard_normal((M, 3))).astype(np.float32)
        yo = c_oracle.mlp_forward(helpers.oracle_mlp(w), x)
        for prec in ("f16", "bf16"):
            y = ops.mlp_forward(m, T(x), prec=prec).cpu().numpy()
            print("%-12s M=%5d %-4s max_abs_err %.3e (|y| max %.3f) finite=%s" % (name, M, prec, np.abs(y - yo).max(), np.abs(yo).max(), np.isfinite(y).all()), flush=True)
    big = torch.randn(148 * 128 * 14, 3, device="cuda") * 0.5
    for prec in ("f16", "f32"):
        ops.mlp_forward(m, big, prec=prec); torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); ops.mlp_forward(m, big, prec=prec); e1.record(); torch.cuda.synchronize()
        flop = 2 * sum(k * n for k, n in m.dims) * big.shape[0]
        print("%-12s %s %d samples: %.3f ms -> %.1f TFLOP/s" % (name, prec, big.shape[0], e0.elapsed_time(e1), flop / e0.elapsed_time(e1) / 1e9), flush=True)
