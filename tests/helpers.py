"""Test helpers: golden loading and conversion of synth weights into oracle objects."""
import os

import numpy as np

import synth
from oracle import c_oracle

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")

# must mirror tests/golden/make_golden.py
MLP_CASES = {
    "sdf_shift": (dict(seed=11, in_size=3, out=1, num_layers=8, hidden=128, freqs=32, sigma=32.0), "softplus"),
    "nerf_first": (dict(seed=12, in_size=3, out=65, num_layers=5, hidden=128, freqs=16, sigma=32.0), None),
    "nerf_second": (dict(seed=13, in_size=70, out=3, num_layers=8, hidden=64, freqs=16, sigma=32.0), None),
    "neural_bsdf": (dict(seed=14, in_size=3, out=3, num_layers=6, hidden=96, freqs=64, sigma=32.0), None),
    "latent_small": (dict(seed=15, in_size=3, out=9, num_layers=5, hidden=32, freqs=16, sigma=32.0, latent=8), None),
    "sp_var_small": (dict(seed=16, in_size=3, out=4, num_layers=7, hidden=256, freqs=128, sigma=128.0), None),
    "one_layer": (dict(seed=17, in_size=5, out=1, num_layers=1, hidden=64, freqs=16, sigma=32.0), None),
}


def golden(name):
    return np.load(os.path.join(GOLDEN, name + ".npz"))


def oracle_mlp(w, act=None):
    return c_oracle.Mlp(w["in_size"], w["out"], w["num_layers"], w["hidden"], w["freqs"], w["basis"],
                        w["W"], w["b"], latent_size=w["latent"], skip=w["skip"],
                        act=c_oracle.ACT_SOFTPLUS if act == "softplus" else c_oracle.ACT_LEAKY_RELU)


def oracle_sdf(w):
    return c_oracle.SphereSdf(w["centers"], w["radii"], w["tfs"], oracle_mlp(w["shift"], "softplus"))


def golden_sdf_weights():
    return synth.sdf_weights(seed=21)


def nerfle_weights(envmap):
    w1 = synth.mlp_weights(seed=31, in_size=3, out=65, num_layers=5, hidden=128, freqs=16, sigma=32.0)
    in2 = 64 + (6 if not envmap else 3 + 16 * 3)
    w2 = synth.mlp_weights(seed=32 + int(envmap), in_size=in2, out=3, num_layers=8, hidden=64, freqs=16, sigma=32.0)
    w1["b"][-1][0] = 0.8
    return w1, w2


def nerfle_ts(fixed_random, S=64):
    # torch.linspace(0, 2 + random.random()*0.1, 64) (nerf.py:178), computed like torch does in fp32
    import torch
    return torch.linspace(0, 2 + float(fixed_random) * 0.1, S).numpy()


def psnr(a, b):
    mse = float(np.mean((np.asarray(a, np.float64) - np.asarray(b, np.float64)) ** 2))
    return 200.0 if mse == 0 else -10.0 * np.log10(mse)


# ---- CUDA-side objects (import torch lazily so CPU-only collection stays cheap) ----
def cuda_mlp(w, act=None, device="cuda"):
    import torch
    from neural_raytracing_b200 import ops
    Ws = [torch.from_numpy(x).to(device) for x in w["W"]]
    bs = [torch.from_numpy(x).to(device) for x in w["b"]]
    params = ops.PackedMLP.pack(Ws, bs)
    return ops.PackedMLP(w["in_size"], w["latent"], w["freqs"], w["hidden"], w["num_layers"], w["skip"], w["out"],
                         ops.ACT_SOFTPLUS if act == "softplus" else ops.ACT_LEAKY_RELU,
                         torch.from_numpy(w["basis"]).to(device), params)


def cuda_sdf(w, device="cuda"):
    import torch
    from neural_raytracing_b200 import ops
    return ops.PackedSDF(torch.from_numpy(w["centers"]).to(device), torch.from_numpy(w["radii"]).to(device),
                         torch.from_numpy(w["tfs"]).to(device), cuda_mlp(w["shift"], "softplus", device))
