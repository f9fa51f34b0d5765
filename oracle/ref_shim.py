"""Import shim for the UNMODIFIED reference (`/root/reference/pytorch3d/pathtracer`).

TEST INFRASTRUCTURE ONLY.  Used in the build container (where /root/reference exists)
to (a) validate the restatements under oracle/ and (b) generate the golden fixtures in
tests/golden/ (see tests/golden/make_golden.py).  Nothing on the product path, in
`-m gpu` tests, smoke() or bench.py imports this file: /root/reference does not exist
on the GPU box.

Why a shim is needed: the reference creates CUDA tensors at import
time (interaction.py:64, utils.py:188-190), imports pytorch_msssim / matplotlib /
fvcore / iopath / pytorch3d._C which are absent here, and trips a Python>=3.11
dataclass check (lights.py:113-116).
"""
import os
import sys
import types

REFERENCE_ROOT = os.environ.get("NRT_REFERENCE_ROOT", "/root/reference")


def available() -> bool:
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "pytorch3d", "pathtracer"))


_loaded = None


def load():
    """Returns the reference `pytorch3d.pathtracer` package, patched to run on CPU."""
    global _loaded
    if _loaded is not None:
        return _loaded
    if not available():
        raise RuntimeError("reference tree not present at %s" % REFERENCE_ROOT)
    os.environ.setdefault("PYTORCH_JIT", "0")
    import torch

    def _cpuify(fn):
        def wrapped(*a, **k):
            d = k.get("device")
            if d is not None and "cuda" in str(d):
                k["device"] = "cpu"
            return fn(*a, **k)
        wrapped.__wrapped_by_nrt__ = True
        return wrapped

    for name in ["tensor", "zeros", "ones", "rand", "randn", "full", "eye", "linspace",
                 "arange", "empty", "normal"]:
        f = getattr(torch, name)
        if not getattr(f, "__wrapped_by_nrt__", False):
            setattr(torch, name, _cpuify(f))
    _to = torch.nn.Module.to
    if not getattr(_to, "__wrapped_by_nrt__", False):
        def to(self, *a, **k):
            a = tuple("cpu" if isinstance(x, (str, torch.device)) and "cuda" in str(x) else x
                      for x in a)
            return _to(self, *a, **k)
        to.__wrapped_by_nrt__ = True
        torch.nn.Module.to = to

    class _Any(types.ModuleType):
        def __getattr__(self, k):
            if k.startswith("__"):
                raise AttributeError(k)
            return object

    for m in ["pytorch_msssim", "matplotlib", "matplotlib.pyplot", "fvcore", "fvcore.common",
              "fvcore.common.file_io", "iopath", "iopath.common", "iopath.common.file_io"]:
        if m not in sys.modules:
            sys.modules[m] = _Any(m)
    sys.modules["pytorch_msssim"].ssim = lambda *a, **k: None
    sys.modules["pytorch_msssim"].ms_ssim = lambda *a, **k: None
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    import pytorch3d
    if "pytorch3d._C" not in sys.modules:
        pytorch3d._C = _Any("pytorch3d._C")
        sys.modules["pytorch3d._C"] = pytorch3d._C
    import pytorch3d.pathtracer.interaction  # noqa: F401
    import pytorch3d.pathtracer.shapes.shapes as _s
    _s.Shape.__hash__ = _s.Sphere.__hash__ = object.__hash__
    import pytorch3d.pathtracer as pt
    _loaded = pt
    return pt
