"""SURVEY 8f rank 3 (dtu.py:93-94, 159): TorchScript SDF archives.  The archive written by save_sdf_archive carries a scriptable
restatement of the reference's SphereSDF; its output is checked against the reference's own values (tests/golden/sdf.npz:
SphereSDF.forward of the unmodified reference on the same weights), and load_sdf_archive brings the tensors back into this
package's SphereSDF.  CPU only: archives are host-side interchange."""
import numpy as np
import pytest
import torch

import helpers


def _sphere_sdf_from(w, device="cpu"):
    from neural_raytracing_b200.pathtracer.shapes.sdfs import SphereSDF
    s = SphereSDF(n=w["n"], device=device)
    with torch.no_grad():
        s.centers.copy_(torch.from_numpy(w["centers"])); s.radii.copy_(torch.from_numpy(w["radii"])); s.tfs.copy_(torch.from_numpy(w["tfs"]))
        m = w["shift"]
        lins = [s.shift.init] + list(s.shift.layers) + [s.shift.out]
        for lin, W, b in zip(lins, m["W"], m["b"]):
            lin.weight.copy_(torch.from_numpy(W)); lin.bias.copy_(torch.from_numpy(b))
    s.shift.basis_p = torch.from_numpy(m["basis"]).clone()
    return s


def test_sdf_archive_round_trip_and_reference_values(tmp_path):
    from neural_raytracing_b200.pathtracer import checkpoint
    g = helpers.golden("sdf")
    s = _sphere_sdf_from(helpers.golden_sdf_weights())
    path = str(tmp_path / "dtu_sdf.pt")
    checkpoint.save_sdf_archive(s, path)
    # what the reference's script does with the file (dtu.py:93): torch.jit.load -> a callable with parameters
    m = torch.jit.load(path, "cpu")
    pts = torch.from_numpy(g["pts"])
    got = m(pts).detach().numpy()
    assert got.shape == g["sdf_vals"].shape
    assert np.abs(got - g["sdf_vals"]).max() < 2e-6, np.abs(got - g["sdf_vals"]).max()      # the unmodified reference's values
    assert m(pts.reshape(10, 30, 3)).shape == (10, 30)                                      # batch dims like SphereSDF.forward
    names = {k for k, _ in m.named_parameters()}
    assert {"centers", "radii", "tfs", "shift.init.weight", "shift.layers.7.bias", "shift.out.weight"} <= names
    assert len(list(m.parameters())) == len(list(s.parameters())) == 23
    # gradients flow to the archive's parameters (the reference optimises density_field.parameters(), dtu.py:125)
    m(pts).sum().backward()
    assert all(p.grad is not None for p in m.parameters())
    # and back into this package
    s2 = checkpoint.load_sdf_archive(path, device="cpu")
    for (k, a), (k2, b) in zip(s.state_dict().items(), s2.state_dict().items()):
        assert k == k2 and torch.equal(a, b), k
    assert torch.equal(s2.shift.basis_p, s.shift.basis_p)
    assert np.abs(s2(pts).detach().numpy() - g["sdf_vals"]).max() < 2e-6


def test_sdf_archive_of_another_module_is_rejected(tmp_path):
    from neural_raytracing_b200.pathtracer import checkpoint
    path = str(tmp_path / "other.pt")
    torch.jit.save(torch.jit.script(torch.nn.Linear(3, 1)), path)
    with pytest.raises(ValueError, match="not a SphereSDF archive"):
        checkpoint.load_sdf_archive(path, device="cpu")


def test_script_archive_is_adopted_with_shared_parameters(tmp_path):
    """What the scripts literally do -- `shape = torch.jit.load(path, device); density_field = SDF(sdf=shape)` (dtu.py:93-94),
    optimise `density_field.parameters()`, `torch.jit.save(density_field.sdf, ...)` (dtu.py:159): the archive is evaluated as a
    SphereSDF of this package that SHARES its parameter tensors, `.sdf` stays the archive."""
    from neural_raytracing_b200.pathtracer import checkpoint
    from neural_raytracing_b200.pathtracer.shapes.sdfs import SDF, SphereSDF
    g = helpers.golden("sdf")
    path = str(tmp_path / "dtu_sdf.pt")
    checkpoint.save_sdf_archive(_sphere_sdf_from(helpers.golden_sdf_weights()), path)
    shape = torch.jit.load(path, "cpu")
    field = SDF(sdf=shape, device="cpu")
    assert field.sdf is shape and isinstance(field._impl, SphereSDF)
    mine, theirs = dict(field._impl.named_parameters()), dict(shape.named_parameters())
    assert mine.keys() == theirs.keys()
    assert all(mine[k].data_ptr() == theirs[k].data_ptr() for k in mine)
    assert {p.data_ptr() for p in field.parameters()} == {p.data_ptr() for p in shape.parameters()}
    pts = torch.from_numpy(g["pts"])
    assert np.abs(field._impl(pts).detach().numpy() - g["sdf_vals"]).max() < 2e-6
    # an optimizer step on density_field.parameters() changes what the archive computes and what torch.jit.save writes
    opt = torch.optim.SGD(field.parameters(), lr=0.1)
    field._impl(pts).sum().backward()
    assert all(p.grad is not None for p in shape.parameters())          # the gradient lands on the archive's tensors
    before = shape(pts).detach().clone()
    opt.step()
    after = shape(pts).detach()
    assert (after - before).abs().max().item() > 1e-3
    assert torch.allclose(field._impl(pts).detach(), after, atol=2e-6)
    out = str(tmp_path / "trained.pt")
    torch.jit.save(field.sdf, out)
    assert torch.allclose(torch.jit.load(out, "cpu")(pts).detach(), after, atol=1e-7)
    # something that is not a SphereSDF archive stays a generic callable
    other = torch.jit.script(torch.nn.Sequential(torch.nn.Linear(3, 1)))
    generic = SDF(sdf=other, device="cpu")
    assert generic._impl is other and generic._fused() is None
    # plain assignment after construction goes through the same adoption
    generic.sdf = shape
    assert isinstance(generic._impl, SphereSDF)


def test_torch_jit_script_of_sphere_sdf_shares_parameters():
    """colocate.py:63 / dtu.py:95 / nerf_synthetic.py:65: `SDF(sdf=torch.jit.script(SphereSDF(n=...)))`.  torch.jit.script
    compiles the TorchScript restatement over the module's own tensors; SDF adopts it back as a SphereSDF on the same tensors."""
    from neural_raytracing_b200.pathtracer.shapes.sdfs import SDF, SphereSDF
    s = SphereSDF(n=2 << 5, device="cpu")
    with torch.no_grad():
        for q in s.shift.parameters():
            q.normal_(0, 0.05)
    js = torch.jit.script(s)
    assert isinstance(js, torch.jit.ScriptModule)
    assert all(a.data_ptr() == b.data_ptr() for a, b in zip(s.parameters(), js.parameters()))
    pts = torch.randn(40, 3) * 0.5
    assert (js(pts) - s.forward_reference_ops(pts)).abs().max().item() < 2e-6
    field = SDF(sdf=js, device="cpu")
    assert field.sdf is js and isinstance(field._impl, SphereSDF)
    assert all(a.data_ptr() == b.data_ptr() for a, b in zip(field.parameters(), s.parameters()))


def test_bsdf_and_light_objects_pickle_like_the_scripts_do():
    """dtu.py:99-113, 163-170: `torch.save(learned_bsdf, path)` / `torch.load(path)` of whole objects, `setattr(bsdf, "act",
    nn.Sigmoid())` on the children.  (The reference as shipped cannot write these files -- lambda activation, INTEGRATION 2b --
    but the scripts' lines must work on this package's classes.)"""
    import io
    import torch.nn as nn
    from neural_raytracing_b200.pathtracer.bsdf import ComposeSpatialVarying, Diffuse, NeuralBSDF
    from neural_raytracing_b200.pathtracer.lights import LightField
    bsdf = ComposeSpatialVarying([NeuralBSDF(device="cpu") for _ in range(2)] +
                                 [Diffuse(preprocess=torch.sigmoid, device="cpu").random()], device="cpu")
    for child in bsdf.bsdfs[:2]:
        setattr(child, "act", nn.Sigmoid())
    lights = LightField(device="cpu")
    for obj in (bsdf, lights):
        buf = io.BytesIO()
        torch.save(obj, buf)
        buf.seek(0)
        back = torch.load(buf, weights_only=False)
        assert type(back) is type(obj)
        for (k, a), (k2, b) in zip(obj.state_dict().items(), back.state_dict().items()):
            assert k == k2 and torch.equal(a, b), k
    pts = torch.randn(7, 3)
    buf = io.BytesIO(); torch.save(bsdf, buf); buf.seek(0)
    back = torch.load(buf, weights_only=False)
    assert isinstance(back.bsdfs[0].act, nn.Sigmoid)
    assert torch.equal(back.sp_var_fn(pts), bsdf.sp_var_fn(pts))
