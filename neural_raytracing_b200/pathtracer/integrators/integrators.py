"""Integrators on the hot path (pytorch3d/pathtracer/integrators/integrators.py): Direct,
NeRFIntegrator, NeRFReproduce, plus the trivial debug views built on the same protocol."""
import torch
import torch.nn as nn

from .. import fused_shading
from ..neural_blocks import SkipConnMLP
from ..shapes.sdfs import SDF
from ..scene import sample_emitter_dir_w_isect, sample_emitter_dir_w_learned_occ, sample_emitter_dir_wo_isect


class Integrator(nn.Module):
    def __init__(self, max_depth=2, russian_roulette_depth=5, sampler=None, lights=None):
        super().__init__()
        self.max_depth = max_depth
        self.rr_depth = russian_roulette_depth
        self.sampler = sampler
        self.lights = lights

    def dims(self):
        raise NotImplementedError()

    def sample(self, shapes, rays, bsdf, **kwargs):
        raise NotImplementedError()


class Debug(Integrator):
    def dims(self):
        return 3

    def sample(self, shapes, rays, bsdf, **kwargs):
        si, active = shapes.intersect(rays)
        return torch.where(active.unsqueeze(-1), (si.n + 1) / 2, torch.zeros_like(si.n)), active, si


class Silhouette(Integrator):
    def dims(self):
        return 1

    def sample(self, shapes, rays, bsdf, **kwargs):
        si, active = shapes.intersect(rays)
        return 1 - active.unsqueeze(-1).float(), active, si


class Mask(Integrator):
    def __init__(self, sub_integrator, **kwargs):
        super().__init__(**kwargs)
        self.sub_integrator = sub_integrator

    def dims(self):
        return self.sub_integrator.dims() + 1

    def sample(self, density_field, rays, bsdf, **kwargs):
        result, active, si = self.sub_integrator.sample(density_field, rays, bsdf, **kwargs)
        result = torch.cat([result, active.float().unsqueeze(-1)], dim=-1)
        return result, torch.ones_like(active), si


class Direct(Integrator):
    """Direct lighting: intersect, sample the emitter (optionally shadow-tested / with learned
    occlusion), evaluate the BSDF, accumulate on the hits (integrators.py:139-206)."""
    DEFAULT_EMITTER_SAMPLES = 1
    DEFAULT_BSDF_SAMPLES = 0

    def dims(self):
        return 3

    def __init__(self, emitter_samples=DEFAULT_EMITTER_SAMPLES, bsdf_samples=DEFAULT_BSDF_SAMPLES, training=True,
                 **kwargs):
        self.emitter_samples = emitter_samples
        self.bsdf_samples = bsdf_samples
        self.training = training
        self.fused = True       # fused shading on the compacted hits where it applies (not in the reference)
        super().__init__(**kwargs)

    def sample(self, shapes, rays, bsdf, **kwargs):
        sampler = kwargs.get("sampler", self.sampler)
        lights = kwargs.get("lights", self.lights)
        w_isect = kwargs.get("w_isect")
        emit = sample_emitter_dir_wo_isect
        if w_isect is True:
            emit = sample_emitter_dir_w_isect
        if isinstance(w_isect, SkipConnMLP):
            def emit(it, s, lights, sampler, active):
                return sample_emitter_dir_w_learned_occ(it, s, lights, sampler, w_isect, active)
        if self.fused and rays.is_cuda and isinstance(shapes, SDF) and fused_shading.supported(bsdf, lights, w_isect):
            # compacted hits + three fused elementwise stages around the MLPs (fused_shading.py): same image, the K hits
            # instead of all R rays through every shading MLP, ~25 launches instead of ~150
            it, active = shapes.intersect(rays, primary=self.training, fused_hits=True)
            result = fused_shading.shade_direct(shapes, rays, it, active, bsdf, lights, w_isect, self.emitter_samples)
            if it._hits is not None:
                self._bsdf_sample_loop(shapes, it, active, bsdf, lights, sampler)
            return result, active, it
        result = torch.zeros(*rays.shape[:-1], 3, device=rays.device)
        it, active = shapes.intersect(rays, primary=self.training)
        if not active.any():
            return result, active, it
        for _ in range(self.emitter_samples):
            ds, emitter_val = emit(it, shapes, lights=lights, sampler=sampler, active=active)
            lit = active & (ds.pdf > 0)
            wo = it.to_local(ds.d)
            bsdf_val, _pdf = bsdf.eval_and_pdf(it, wo, active=lit)
            result[lit] = result[lit] + bsdf_val[lit] * emitter_val[lit] / self.emitter_samples
        self._bsdf_sample_loop(shapes, it, active, bsdf, lights, sampler)
        return result, active, it

    def _bsdf_sample_loop(self, shapes, it, active, bsdf, lights, sampler):
        """integrators.py:189-204: the unfinished BSDF-sampling loop.  It samples the BSDF, intersects the bounce with
        the emitters and raises only if a bounce actually hits one; PointLights / LightField.intersect return
        (None, False), so with those a script configured with bsdf_samples > 0 runs to completion (without any
        contribution), exactly as in the reference."""
        for _ in range(self.bsdf_samples):
            bs, _bsdf_val = bsdf.sample(it, sampler=sampler, active=active)
            _it2, snd_active = lights.intersect(it.spawn_rays(it.from_local(bs.wo)))
            bsdf_active = active & snd_active
            if bsdf_active.any():
                raise NotImplementedError("BSDF-sampled emitter hits are not implemented in the reference either "
                                          "(integrators.py:198)")


class NeRFIntegrator(Integrator):
    """Appends sigmoid(throughput) as an alpha channel and marks every pixel valid (integrators.py:243-257)."""

    def __init__(self, sub_integrator, **kwargs):
        super().__init__(**kwargs)
        self.sub_integrator = sub_integrator

    def dims(self):
        return self.sub_integrator.dims() + 1

    def sample(self, density_field, rays, bsdf, **kwargs):
        result, active, mi = self.sub_integrator.sample(density_field, rays, bsdf, **kwargs)
        alpha = mi.throughput.unsqueeze(-1)
        if mi.with_logits:
            alpha = alpha.sigmoid()
        return torch.cat([result, alpha], dim=-1), torch.tensor(True, device=result.device), mi


class NeRFReproduce(Integrator):
    """Uses a volumetric NeRF module in place of surface integration (integrators.py:260-267)."""

    def dims(self):
        return 3

    def sample(self, nerf, rays, lights, **kwargs):
        result = nerf(rays, lights)

        class Dummy:
            ...
        return result, torch.tensor(True, device=result.device), Dummy()


class Path(Integrator):
    """Multi-bounce light path integrator (integrators.py:274-354; scripts/path_nerv.py): direct lighting at every
    vertex, then a BSDF-sampled bounce whose secondary rays go through the same march kernels
    (`intersect(primary=False)`: no silhouette scan).  Kept from the reference: MIS weights are 1, the throughput is
    detached between bounces, Russian roulette is not implemented (the reference asserts when depth exceeds
    `rr_depth`; so does this), emitters are sampled at the bounce vertices only through the light sampler."""

    def __init__(self, training=False, **kwargs):
        super().__init__(**kwargs)
        self.training = training

    def dims(self):
        return 3

    def _emitter_fn(self, kwargs):
        w_isect = kwargs.get("w_isect", False)
        if isinstance(w_isect, SkipConnMLP):
            return lambda it, shapes, lights, sampler, active: \
                sample_emitter_dir_w_learned_occ(it, shapes, lights, sampler, w_isect, active)
        if w_isect is True:
            return sample_emitter_dir_w_isect
        return kwargs.get("sample_emitter_fn", sample_emitter_dir_wo_isect)

    def sample(self, shapes, rays, bsdf, **kwargs):
        sampler = kwargs.get("sampler", self.sampler)
        lights = kwargs.get("lights", self.lights)
        sample_emitter = self._emitter_fn(kwargs)
        throughput = torch.ones(*rays.shape[:-1], 3, device=rays.device)
        result = torch.zeros_like(throughput)
        first_it, active = shapes.intersect(rays, primary=self.training)
        if not active.any():
            return result, active, first_it
        first_active = active.clone()
        it = first_it
        for depth in range(self.max_depth):
            assert depth <= self.rr_depth, "Russian roulette is not implemented (neither in the reference)"
            if active.any():
                ds, emitter_val = sample_emitter(it, shapes, lights=lights, sampler=sampler, active=active)
                lit = active & (ds.pdf > 0)
                bsdf_val, _pdf = bsdf.eval_and_pdf(it, it.to_local(ds.d), active=lit)
                result = result + torch.where(lit.unsqueeze(-1), throughput * bsdf_val * emitter_val,
                                              torch.zeros_like(result))
            bs, bounce_val = bsdf.sample(it, sampler=sampler, active=active)
            throughput = (bounce_val.clamp(min=1e-10) * throughput).detach()
            active = active & (throughput > 0).any(-1)
            if not active.any():
                break
            it, hits = shapes.intersect(it.spawn_rays(it.from_local(bs.wo)), active=active, primary=False)
            active = active & hits
            if not active.any():
                break
        return result, first_active, first_it
