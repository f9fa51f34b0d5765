"""CPU tests of include/nrt_shade_math.h, the scalar shading functions the fused Direct-integrator kernels compile (forward
on float, backward on dual numbers): values against the mirror's torch expressions -- themselves pinned to the unmodified
reference by tests/golden/shading.npz / pipeline.npz -- and Jacobians against torch autograd in float64.  The header is
host-compilable; tests/host/shade_host.cpp wraps it (test infrastructure, g++ only)."""
import ctypes
import os
import subprocess

import numpy as np
import pytest
import torch
import torch.nn.functional as F

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    out = os.path.join(ROOT, "build", "host")
    os.makedirs(out, exist_ok=True)
    so = os.path.join(out, "libshade_host.so")
    subprocess.check_call(["g++", "-O2", "-std=c++17", "-ffp-contract=off", "-shared", "-fPIC", "-I" + os.path.join(ROOT, "include"),
                           os.path.join(ROOT, "tests", "host", "shade_host.cpp"), "-o", so])
    L = ctypes.CDLL(so)
    L.shade_eval.restype = ctypes.c_int
    return L


def _eval(L, fn, x, c=(0, 0, 0), n_in=None):
    x = np.asarray(x, np.float32)
    c = np.asarray(c, np.float32)
    y, J, yf = np.zeros(8, np.float32), np.zeros(64, np.float32), np.zeros(8, np.float32)
    P = lambda a: a.ctypes.data_as(ctypes.c_void_p)
    n_out = L.shade_eval(fn, P(x), P(c), P(y), P(J), P(yf))
    assert n_out > 0
    return y[:n_out], J[:n_out * len(x)].reshape(n_out, len(x)), yf[:n_out]


def _torch_fns():
    from neural_raytracing_b200.pathtracer import interaction as I, utils as U
    from neural_raytracing_b200.pathtracer.bsdf import bsdfs as B

    def normalize6(x, c):
        return F.normalize(x, eps=1e-6, dim=-1)

    def to_local_n(x, c):
        return I.to_local(I.coordinate_system(x[:3]), x[3:])

    def rusin(x, c):
        return U.param_rusin2(x[:3], x[3:])

    def elaz(x, c):
        return U.dir_to_elev_azim(x)

    def fresnel(x, c):
        return B.fresnel_conductor(x[0], x[1], 0.0).reshape(1)

    def point_light(x, c):
        d = c - x
        return torch.cat([F.normalize(d, eps=1e-6, dim=-1), torch.linalg.norm(d, dim=-1, keepdim=True)])

    def light_field(x, c):
        return torch.cat([F.normalize(x, eps=1e-6, dim=-1).clamp(min=1e-6, max=1), torch.linalg.norm(x, dim=-1, keepdim=True)])

    def geom(x, c):
        n = F.normalize(x, eps=1e-6, dim=-1)
        return torch.cat([n, I.to_local(I.coordinate_system(n), -c)])

    def denom(x, c):
        return (c[0] + c[1] * x + c[2] * x.square()).clamp(min=1e-6)
    return [normalize6, to_local_n, rusin, elaz, fresnel, point_light, light_field, geom, denom]


CASES = [  # fn, n_in, input scale, constants
    (0, 3, 1.0, None), (0, 3, 1e-7, None), (1, 6, 1.0, None), (2, 6, 1.0, None), (3, 3, 1.0, None), (4, 2, None, None),
    (5, 3, 0.5, (0.9, 0.5, 0.7)), (6, 3, 1.0, None), (7, 3, 1.0, (0.3, -0.5, 0.8)), (8, 1, None, (1e-6, 1e-6, 1.0)),
]


@pytest.mark.parametrize("fn,n_in,scale,const", CASES)
def test_values_and_jacobians_match_torch(lib, fn, n_in, scale, const):
    fns = _torch_fns()
    rs = np.random.RandomState(100 + fn)
    worst_v = worst_j = 0.0
    for trial in range(200):
        if fn == 4:
            x = np.array([rs.uniform(0.02, 1.0), rs.uniform(0.3, 3.0)], np.float32)
        elif fn == 8:
            x = np.array([rs.uniform(0.05, 3.0)], np.float32)
        else:
            x = (scale * rs.standard_normal(n_in)).astype(np.float32)
            if fn == 6 and trial % 2 == 0:
                x = np.abs(x)          # the light field's direction is clamped to [1e-6, 1]: exercise the unclamped side too
        c = np.asarray(const if const is not None else (0, 0, 0), np.float32)
        if fn == 7:
            c = c / np.linalg.norm(c)
        y, J, yf = _eval(lib, fn, x, c)
        assert np.array_equal(y, yf), "the dual-number value differs from the float evaluation"
        xt = torch.tensor(x, dtype=torch.float64, requires_grad=True)
        ct = torch.tensor(c, dtype=torch.float64)
        f = lambda v: fns[fn](v, ct)
        yt = f(xt)
        Jt = torch.autograd.functional.jacobian(f, xt)
        worst_v = max(worst_v, float(np.abs(y - yt.detach().numpy()).max()))
        # compare Jacobians relative to their scale; skip inputs that sit on a clamp / branch of the function (the two
        # implementations may then legitimately pick different one-sided derivatives)
        Jt = Jt.numpy()
        denom_ = max(1.0, float(np.abs(Jt).max()))
        err = float(np.abs(J - Jt).max()) / denom_
        if err > 1e-3:
            xs = torch.tensor(x.astype(np.float64) * (1 + 1e-6))
            if np.abs(torch.autograd.functional.jacobian(f, xs).numpy() - Jt).max() / denom_ > 1e-3:
                continue          # derivative discontinuous here
        worst_j = max(worst_j, err)
    assert worst_v < 2e-5, worst_v
    assert worst_j < 1e-3, worst_j
