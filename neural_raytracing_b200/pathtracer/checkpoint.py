"""Checkpoint interchange with the reference (scripts/dtu.py:93-108, 159-170).

What the reference's scripts write:
  * `torch.jit.save(density_field.sdf, path)`: a TorchScript archive of the (scripted) SphereSDF.  The archive is
    self-contained (code + tensors); what identifies the model is its parameter naming -- `centers`, `radii`, `tfs`,
    `shift.init.*`, `shift.layers.N.*`, `shift.out.*` (the layout tests/test_state_dict_cpu.py pins to the reference) --
    plus the tensor attribute `shift.basis_p` (a plain attribute in neural_blocks.py:38, not a buffer: it is in the
    archive but not in state_dict()).
  * `torch.save(learned_bsdf / lights, path)`: pickles of whole objects.  With the reference as shipped these cannot be
    produced: SkipConnMLP's default activation is a lambda (neural_blocks.py:26) and pickling raises PicklingError
    (reproduced with the unmodified reference in this container: ComposeSpatialVarying([NeuralBSDF(), ...]) and
    LightField() both fail), so there is no such format to read.  The interchange for BSDFs and light fields is
    `state_dict()` / `load_state_dict()`, whose layout is identical in both packages.

This module: `load_sdf_archive` (archive -> SphereSDF of this package, on the fused kernels), `save_sdf_archive`
(SphereSDF -> archive that the reference's `torch.jit.load` + `SDF(sdf=...)` runs, dtu.py:93-94) and `ScriptSphereSDF`, the
scriptable restatement of sdfs.py:37-46 + neural_blocks.py:75-86 that the saved archive carries as its code."""
from typing import List

import torch
import torch.nn as nn
import torch.nn.functional as F


class _ScriptSkipConnMLP(nn.Module):
    """neural_blocks.py:75-86 for the softplus / leaky_relu nets, TorchScript-compatible; same attribute names."""

    def __init__(self, in_size: int, dim_p: int, hidden: int, out: int, num_layers: int, skip: int, softplus: bool):
        super().__init__()
        self.in_size = in_size
        self.skip = skip
        self.softplus = softplus
        self.basis_p = torch.zeros(in_size, (dim_p - in_size) // 2)
        widths = [hidden + dim_p if (i % skip) == 0 and i != num_layers - 1 else hidden for i in range(num_layers)]
        self.init = nn.Linear(dim_p, hidden)
        self.layers = nn.ModuleList([nn.Linear(w, hidden) for w in widths])
        self.out = nn.Linear(hidden, out)

    def act(self, x):
        return F.softplus(x) if self.softplus else F.leaky_relu(x)

    def forward(self, p):
        batches: List[int] = list(p.shape[:-1])
        x0 = p.reshape(-1, self.in_size)
        mapped = x0 @ self.basis_p                                    # utils.py:37-40
        init = torch.cat([x0, mapped.sin(), mapped.cos()], dim=-1)
        x = self.init(init)
        n = len(self.layers)
        for i, layer in enumerate(self.layers):
            if i != n - 1 and (i % self.skip) == 0:
                x = torch.cat([x, init], dim=-1)
            x = layer(self.act(x))
        return self.out(self.act(x)).reshape(batches + [self.out.out_features])


class ScriptSphereSDF(nn.Module):
    """sdfs.py:11-46 (SphereSDF.forward = smooth-min of the warped spheres + residual MLP), TorchScript-compatible."""

    def __init__(self, n: int, in_size: int, dim_p: int, hidden: int, num_layers: int, skip: int):
        super().__init__()
        self.centers = nn.Parameter(torch.zeros(n, 3))
        self.radii = nn.Parameter(torch.zeros(n))
        self.tfs = nn.Parameter(torch.zeros(n, 3, 3))
        self.shift = _ScriptSkipConnMLP(in_size, dim_p, hidden, 1, num_layers, skip, True)

    def forward(self, p):
        shape: List[int] = list(p.shape[:-1])
        pts = p.reshape(-1, 3)
        tfs = self.tfs + torch.eye(3, device=p.device).unsqueeze(0)
        q = torch.einsum("ijk,bk->ibj", tfs, pts) - self.centers.unsqueeze(1)
        sd = q.norm(p=2, dim=-1) - self.radii.unsqueeze(-1)
        sm = -(torch.exp(-32.0 * sd).sum(dim=0).clamp(min=1e-4)).log() / 32.0          # utils.py:385-387, k = 32
        return sm.reshape(shape) + self.shift(p).reshape(shape)


def _shape_of(sd, basis_p):
    n = sd["centers"].shape[0]
    hidden, dim_p = sd["shift.init.weight"].shape
    num_layers = len([k for k in sd if k.startswith("shift.layers.") and k.endswith(".weight")])
    in_size = basis_p.shape[0]
    if dim_p != in_size + 2 * basis_p.shape[1]:
        raise ValueError("archive: shift.init expects %d inputs, basis_p %s gives %d" %
                         (dim_p, tuple(basis_p.shape), in_size + 2 * basis_p.shape[1]))
    return n, in_size, dim_p, hidden, num_layers


def load_sdf_archive(path, device="cuda"):
    """TorchScript archive written by the reference (`torch.jit.save(density_field.sdf, ...)`, dtu.py:159) or by
    `save_sdf_archive` -> SphereSDF of this package with the archive's tensors (use as `SDF(sdf=load_sdf_archive(p))`:
    the fused march / scan / normals kernels).  An archive that is not a SphereSDF is an error naming what is supported
    (a generic callable still works as `SDF(sdf=torch.jit.load(p))` on the unfused march)."""
    from .shapes.sdfs import SphereSDF
    m = torch.jit.load(path, map_location="cpu")
    sd = {k: v.detach() for k, v in m.state_dict().items()}
    need = {"centers", "radii", "tfs", "shift.init.weight", "shift.init.bias", "shift.out.weight", "shift.out.bias"}
    if not need.issubset(sd.keys()):
        raise ValueError("load_sdf_archive: %s is not a SphereSDF archive (parameters %s...); supported: the reference's "
                         "SphereSDF (sdfs.py:11-46).  Other SDF callables: SDF(sdf=torch.jit.load(path))" %
                         (path, sorted(sd.keys())[:6]))
    basis_p = m.shift.basis_p.detach().to("cpu", torch.float32)
    n, in_size, dim_p, hidden, num_layers = _shape_of(sd, basis_p)
    out = SphereSDF(n=n, device=device)
    ref = out.shift
    if (ref.in_size, ref.init.in_features, ref.init.out_features, len(ref.layers)) != (in_size, dim_p, hidden, num_layers):
        raise ValueError("load_sdf_archive: residual MLP %dx%d on %d inputs; SphereSDF builds %dx%d on %d (sdfs.py:23-31)" %
                         (num_layers, hidden, dim_p, len(ref.layers), ref.init.out_features, ref.init.in_features))
    missing, unexpected = out.load_state_dict({k: v.to(device) for k, v in sd.items()}, strict=False)
    if missing or unexpected:
        raise ValueError("load_sdf_archive: state mismatch, missing %s unexpected %s" % (missing, unexpected))
    out.shift.basis_p = basis_p.to(device)
    out.invalidate_packed()
    return out


def save_sdf_archive(sphere_sdf, path):
    """SphereSDF of this package -> TorchScript archive in the reference's format: `torch.jit.load(path, device)` gives a
    callable with the reference's parameter names (dtu.py:93-94 wraps it in SDF(sdf=...); `.parameters()` feeds AdamW)."""
    sd = {k: v.detach().to("cpu", torch.float32) for k, v in sphere_sdf.state_dict().items()}
    basis_p = sphere_sdf.shift.basis_p.detach().to("cpu", torch.float32)
    n, in_size, dim_p, hidden, num_layers = _shape_of(sd, basis_p)
    m = ScriptSphereSDF(n, in_size, dim_p, hidden, num_layers, int(sphere_sdf.shift.skip))
    m.load_state_dict(sd, strict=True)
    m.shift.basis_p = basis_p.clone()
    scripted = torch.jit.script(m)
    torch.jit.save(scripted, path)
    return path


def adopt_script_sphere_sdf(scripted):
    """A TorchScript SphereSDF archive as the scripts hold it (`shape = torch.jit.load(path, device); SDF(sdf=shape)`,
    dtu.py:93-94, nerf_synthetic.py:63-64, colocate.py:65-66) -> a SphereSDF of this package whose parameters ARE the
    archive's parameter tensors (same objects: an optimizer step or a load_state_dict on either side is seen by both, and
    `torch.jit.save(density_field.sdf, ...)` later writes the trained weights), so that SDF runs the fused march / scan /
    normals kernels on it.  None when `scripted` is not a SphereSDF archive (SDF then marches the callable generically)."""
    from .shapes.sdfs import SphereSDF
    if not isinstance(scripted, torch.jit.ScriptModule):
        return None
    params = dict(scripted.named_parameters())
    need = {"centers", "radii", "tfs", "shift.init.weight", "shift.init.bias", "shift.out.weight", "shift.out.bias"}
    if not need.issubset(params.keys()) or not hasattr(scripted, "shift") or not hasattr(scripted.shift, "basis_p"):
        return None
    basis_p = scripted.shift.basis_p
    try:
        n, in_size, dim_p, hidden, num_layers = _shape_of(params, basis_p)
    except ValueError:
        return None
    device = params["centers"].device
    out = SphereSDF(n=n, device=device)
    ref = out.shift
    if (ref.in_size, ref.init.in_features, ref.init.out_features, len(ref.layers)) != (in_size, dim_p, hidden, num_layers):
        return None
    own = dict(out.named_parameters())
    if set(own.keys()) != set(params.keys()):
        return None
    for name, tensor in params.items():
        if tuple(own[name].shape) != tuple(tensor.shape):
            return None
        mod = out
        *path, leaf = name.split(".")
        for part in path:
            mod = getattr(mod, part)
        mod._parameters[leaf] = tensor                 # share, do not copy
    out.shift.basis_p = basis_p.detach().to(device=device, dtype=torch.float32)
    out.invalidate_packed()
    return out


def scriptable_view(sphere_sdf):
    """`torch.jit.script(SphereSDF(n=...))` (colocate.py:63, dtu.py:95, nerf_synthetic.py:65): the module torch.jit.script
    compiles instead of this package's SphereSDF (whose forward dispatches to the library and is not TorchScript) -- the
    scriptable restatement of sdfs.py:37-46 over the SAME parameter tensors.  `SDF(sdf=<the scripted module>)` adopts it back
    (adopt_script_sphere_sdf), so the script's line runs on the fused kernels and trains the tensors it was given."""
    params = dict(sphere_sdf.named_parameters())
    basis_p = sphere_sdf.shift.basis_p.detach()
    n, in_size, dim_p, hidden, num_layers = _shape_of(params, basis_p)
    m = ScriptSphereSDF(n, in_size, dim_p, hidden, num_layers, int(sphere_sdf.shift.skip))
    for name, tensor in params.items():
        mod = m
        *path, leaf = name.split(".")
        for part in path:
            mod = getattr(mod, part)
        mod._parameters[leaf] = tensor
    m.shift.basis_p = basis_p.to(torch.float32)
    return m
