"""ctypes front-end of the C oracle (oracle/c/nrt_oracle.c).  TEST INFRASTRUCTURE ONLY."""
import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_build", "libnrt_oracle.so")

ACT_LEAKY_RELU, ACT_SOFTPLUS = 0, 1
OUT_NONE, OUT_SIGMOID, OUT_SOFTPLUS, OUT_TANH = 0, 1, 2, 3

_f32p = ctypes.POINTER(ctypes.c_float)


class NrtMlp(ctypes.Structure):
    _fields_ = [("in_size", ctypes.c_int32), ("latent_size", ctypes.c_int32),
                ("freqs", ctypes.c_int32), ("hidden", ctypes.c_int32),
                ("num_layers", ctypes.c_int32), ("skip", ctypes.c_int32),
                ("out_size", ctypes.c_int32), ("act", ctypes.c_int32),
                ("basis", ctypes.c_void_p), ("params", ctypes.c_void_p),
                ("params_tc", ctypes.c_void_p)]


class NrtSphereSdf(ctypes.Structure):
    _fields_ = [("n", ctypes.c_int32), ("centers", ctypes.c_void_p), ("radii", ctypes.c_void_p),
                ("tfs", ctypes.c_void_p), ("shift", NrtMlp)]


def _cpu_has(flag):
    try:
        with open("/proc/cpuinfo") as f:
            for line in f:
                if line.startswith("flags"):
                    return flag in line.split()
    except OSError:
        pass
    return False


def build(force=False):
    """Compiles the oracle with gcc (no-op if up to date)."""
    src = os.path.join(_HERE, "c", "nrt_oracle.c")
    hdrs = [os.path.join(_HERE, "..", "include", h) for h in ("nrt_detmath.h", "nrt_b200.h")]
    stamp = _SO + ".flags"
    want = "avx2fma" if (_cpu_has("avx2") and _cpu_has("fma")) else "generic"
    have = open(stamp).read().strip() if os.path.exists(stamp) else ""
    fresh = os.path.exists(_SO) and all(os.path.getmtime(_SO) >= os.path.getmtime(p)
                                        for p in [src] + hdrs)
    if fresh and have == want and not force:
        return _SO
    target = "_build/libnrt_oracle.so" if want == "avx2fma" else "generic"
    if os.path.exists(_SO):
        os.remove(_SO)
    subprocess.check_call(["make", "-C", _HERE, target], stdout=subprocess.DEVNULL)
    with open(stamp, "w") as f:
        f.write(want)
    return _SO


_lib = None


def lib():
    global _lib
    if _lib is None:
        _lib = ctypes.CDLL(build())
        _lib.oracle_mlp_param_count.restype = ctypes.c_int64
    return _lib


class Mlp:
    """Host-side SkipConnMLP parameters in the packed-f32 layout of include/nrt_b200.h."""

    def __init__(self, in_size, out_size, num_layers, hidden, freqs, basis, weights, biases,
                 latent_size=0, skip=3, act=ACT_LEAKY_RELU):
        # weights/biases in torch layout and order [init, layers..., out]: W [N,K], b [N]
        self.in_size, self.out_size, self.num_layers, self.hidden = in_size, out_size, num_layers, hidden
        self.freqs, self.latent_size, self.skip, self.act = freqs, latent_size, skip, act
        self.basis = np.ascontiguousarray(basis, np.float32).reshape(in_size, freqs)
        order = [0] + list(range(1, num_layers + 1)) + [num_layers + 1]
        chunks = []
        for i in order:
            chunks.append(np.ascontiguousarray(np.asarray(weights[i], np.float32).T).ravel())
            chunks.append(np.asarray(biases[i], np.float32).ravel())
        self.params = np.ascontiguousarray(np.concatenate(chunks), np.float32)
        c = self.c_struct()
        n = lib().oracle_mlp_param_count(ctypes.byref(c))
        assert n == self.params.size, (n, self.params.size)

    def c_struct(self):
        return NrtMlp(self.in_size, self.latent_size, self.freqs, self.hidden, self.num_layers,
                      self.skip, self.out_size, self.act, self.basis.ctypes.data,
                      self.params.ctypes.data, None)


class SphereSdf:
    def __init__(self, centers, radii, tfs, shift: Mlp):
        self.centers = np.ascontiguousarray(centers, np.float32)
        self.radii = np.ascontiguousarray(radii, np.float32)
        self.tfs = np.ascontiguousarray(tfs, np.float32)
        self.shift = shift
        self.n = self.radii.shape[0]

    def c_struct(self):
        return NrtSphereSdf(self.n, self.centers.ctypes.data, self.radii.ctypes.data,
                            self.tfs.ctypes.data, self.shift.c_struct())


def _p(a):
    return a.ctypes.data_as(ctypes.c_void_p)


def _f32(a):
    return np.ascontiguousarray(a, np.float32)


def mlp_forward(m: Mlp, x, latent=None, out_act=OUT_NONE):
    x = _f32(x).reshape(-1, m.in_size)
    M = x.shape[0]
    lat = _f32(latent).reshape(M, m.latent_size) if m.latent_size else None
    out = np.empty((M, m.out_size), np.float32)
    c = m.c_struct()
    rc = lib().oracle_mlp_forward(ctypes.byref(c), ctypes.c_int(out_act), _p(x),
                                  _p(lat) if lat is not None else None, ctypes.c_int64(M), _p(out))
    assert rc == 0
    return out


def sdf_eval(s: SphereSdf, p):
    p = _f32(p).reshape(-1, 3)
    out = np.empty(p.shape[0], np.float32)
    c = s.c_struct()
    assert lib().oracle_sdf_eval(ctypes.byref(c), _p(p), ctypes.c_int64(p.shape[0]), _p(out)) == 0
    return out


def sdf_value_grad(s: SphereSdf, p):
    p = _f32(p).reshape(-1, 3)
    val = np.empty(p.shape[0], np.float32)
    grad = np.empty((p.shape[0], 3), np.float32)
    c = s.c_struct()
    assert lib().oracle_sdf_value_grad(ctypes.byref(c), _p(p), ctypes.c_int64(p.shape[0]), _p(val),
                                       _p(grad)) == 0
    return val, grad


def sphere_trace(s: SphereSdf, rays, eps=1e-3, max_steps=64, max_t=10.0):
    rays = _f32(rays).reshape(-1, 6)
    R = rays.shape[0]
    depth = np.empty(R, np.float32)
    hit = np.empty(R, np.uint8)
    c = s.c_struct()
    ev = ctypes.c_ulonglong(0)
    assert lib().oracle_sphere_trace(ctypes.byref(c), _p(rays), ctypes.c_int64(R), ctypes.c_float(eps),
                                     ctypes.c_int(max_steps), ctypes.c_float(max_t), _p(depth), _p(hit),
                                     ctypes.byref(ev)) == 0
    return depth, hit.astype(bool)


def shadow_test(s: SphereSdf, rays, max_t, eps=1e-3, max_steps=64):
    rays = _f32(rays).reshape(-1, 6)
    R = rays.shape[0]
    max_t = _f32(max_t).reshape(R)
    nb = np.empty(R, np.uint8)
    c = s.c_struct()
    assert lib().oracle_shadow_test(ctypes.byref(c), _p(rays), _p(max_t), ctypes.c_int64(R),
                                    ctypes.c_float(eps), ctypes.c_int(max_steps), _p(nb)) == 0
    return nb.astype(bool)


def min_scan(s: SphereSdf, rays, step, n_steps=128):
    rays = _f32(rays).reshape(-1, 6)
    R = rays.shape[0]
    idx = np.empty(R, np.int32)
    pos = np.empty((R, 3), np.float32)
    mv = np.empty(R, np.float32)
    c = s.c_struct()
    assert lib().oracle_min_scan(ctypes.byref(c), _p(rays), ctypes.c_int64(R), ctypes.c_double(step),
                                 ctypes.c_int(n_steps), _p(idx), _p(pos), _p(mv)) == 0
    return idx, pos, mv


def composite(sigma_raw, rgb, ts):
    sigma_raw = _f32(sigma_raw)
    S, R = sigma_raw.shape
    rgb = _f32(rgb).reshape(S, R, 3)
    ts = _f32(ts).reshape(S)
    out = np.empty((R, 3), np.float32)
    assert lib().oracle_composite(_p(sigma_raw), _p(rgb), _p(ts), ctypes.c_int(S), ctypes.c_int64(R),
                                  _p(out)) == 0
    return out


def nerfle_render(first: Mlp, second: Mlp, rays, ts=None, ts_per_ray=None, light_code=None,
                  view_of_ray=None, second_out_act=OUT_SIGMOID, store=False):
    rays = _f32(rays).reshape(-1, 6)
    R = rays.shape[0]
    if ts is not None:
        ts = _f32(ts).ravel()
        S = ts.size
    else:
        ts_per_ray = _f32(ts_per_ray).reshape(R, -1)
        S = ts_per_ray.shape[1]
    light_code = _f32(light_code).reshape(-1, second.in_size - (first.out_size - 1) - 3)
    light_dim = light_code.shape[1]
    vor = np.ascontiguousarray(view_of_ray, np.int32) if view_of_ray is not None else None
    out = np.empty((R, 3), np.float32)
    sig = np.empty((R, S), np.float32) if store else None
    srgb = np.empty((R, S, 3), np.float32) if store else None
    c1, c2 = first.c_struct(), second.c_struct()
    rc = lib().oracle_nerfle_render(
        ctypes.byref(c1), ctypes.byref(c2), _p(rays), ctypes.c_int64(R),
        _p(ts) if ts is not None else None, _p(ts_per_ray) if ts_per_ray is not None else None,
        ctypes.c_int(S), _p(light_code), ctypes.c_int(light_dim), _p(vor) if vor is not None else None,
        ctypes.c_int(second_out_act), _p(out), _p(sig) if store else None, _p(srgb) if store else None)
    assert rc == 0, rc
    return (out, sig, srgb) if store else out
