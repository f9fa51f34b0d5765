"""Scene construction shared by tests/golden/make_golden.py (run on the UNMODIFIED reference package)
and the GPU pipeline tests (run on the drop-in mirror): the same code builds both, which is itself a
check of the drop-in surface (constructor kwargs, attributes that scripts mutate)."""
import torch

import synth


def build_pipeline(P, kind, device="cpu"):
    """Builds the colocate-style ('colocate') or DTU-style ('dtu') scene out of the package `P`
    (either the reference `pytorch3d.pathtracer` or the drop-in mirror): identical code for both."""
    import torch.nn as nn
    sdfs, bsdfm, lightsm, integ = P.shapes.sdfs, P.bsdf, P.lights, P.integrators
    sphere = sdfs.SphereSDF(n=64, device=device)
    synth.fill_module(sphere, 61, shift_std=0.02)
    shape = sdfs.SDF(sdf=sphere, device=device, max_steps=64)
    if kind == "colocate":
        # colocate.py:63-85: 2 neural + diffuse + conductor bases, point light, learned occlusion MLP
        kids = [bsdfm.NeuralBSDF(device=device), bsdfm.NeuralBSDF(device=device),
                bsdfm.Diffuse(preprocess=nn.Softplus(), device=device), bsdfm.Conductor(activation=nn.Softplus(), device=device)]
        kids[2].reflectance = torch.tensor([0.3, 0.6, 0.2], device=device, requires_grad=True)
        kids[3].specular = torch.tensor([0.7, 0.4, 0.9], device=device, requires_grad=True)
        lights = lightsm.PointLights(device=device, location=[[0.9, 0.5, 0.7]], scale=5)
        occ = P.neural_blocks.SkipConnMLP(in_size=5, out=1, device=device).to(device)
        synth.fill_module(occ, 65)
        integrator, w_isect = integ.Direct(), occ
    else:
        # dtu.py:95-106: neural + diffuse(sigmoid) bases, learned light field, NeRFIntegrator(Direct())
        kids = [bsdfm.NeuralBSDF(activation=nn.Sigmoid(), device=device), bsdfm.NeuralBSDF(activation=nn.Sigmoid(), device=device),
                bsdfm.Diffuse(preprocess=torch.sigmoid, device=device)]
        kids[2].reflectance = torch.tensor([0.2, -0.4, 0.5], device=device, requires_grad=True)
        lights = lightsm.LightField(device=device)
        synth.fill_module(lights, 66)
        integrator, w_isect = integ.NeRFIntegrator(integ.Direct()), False
    for i, k in enumerate(kids[:2]):
        synth.fill_module(k, 62 + i)
    bsdf = bsdfm.ComposeSpatialVarying(kids, device=device)
    bsdf.sp_var_fn._synth_sigma = 128.0
    synth.fill_module(bsdf, 64)
    return shape, sphere, bsdf, lights, integrator, w_isect


