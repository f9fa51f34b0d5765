"""Visualisation integrators the vis scripts put on top of the hot path (pytorch3d/pathtracer/integrators/integrators.py:57-136;
dtu_vis.py:125-142, nerv_vis.py:119-156, visualize.py:95-120): each one is the sphere-trace march (shapes.intersect) plus a few
elementwise ops on the interaction, so they share one template -- intersect, then `_view(it, active, ...)` -- instead of five
copies of the protocol."""
import torch
import torch.nn.functional as F

from ..scene import sample_emitter_dir_wo_isect
from .integrators import Integrator


class _SurfaceView(Integrator):
    """intersect -> per-ray view of the interaction.  Subclasses return the [.., dims()] values; rays that miss keep whatever
    `_view` wrote for them (the reference's conventions differ per integrator and are kept)."""

    def sample(self, shapes, rays, bsdf=None, **kwargs):
        it, active = shapes.intersect(rays)
        return self._view(shapes, rays, it, active, bsdf=bsdf, **kwargs), active, it


class Depth(_SurfaceView):
    """Hit distance along the ray, `empty_val` where nothing was hit (integrators.py:57-67; the optional rescale by the
    maximum is the reference's, including its exclusion of exact zeros)."""

    def __init__(self, scale=False, empty_val=-1, **kwargs):
        super().__init__(**kwargs)
        self.empty_val, self.scale = empty_val, scale

    def dims(self):
        return 1

    def _view(self, shapes, rays, it, active, **kwargs):
        t = torch.where(active, it.t, torch.full_like(it.t, self.empty_val))
        if self.scale:
            nz = t != 0
            t[nz] = t[nz] / t[nz].max()
        return t.unsqueeze(-1)


class BasisBRDF(_SurfaceView):
    """Weight map of a spatially varying BSDF: sigmoid(sp_var_fn(p)) at the hits, zero elsewhere (integrators.py:79-90).
    The 256-wide sp_var network runs on the compacted hits only."""

    def __init__(self, multi_basis_bsdf):
        super().__init__()
        self.bsdf = multi_basis_bsdf

    def dims(self):
        return len(self.bsdf.bsdfs)

    def _view(self, shapes, rays, it, active, **kwargs):
        out = torch.zeros(*rays.shape[:-1], self.dims(), device=rays.device)
        if active.any():
            out[active] = self.bsdf.normalized_weights(it.p[active], it)
        return out


class _EmitterView(_SurfaceView):
    """Views of the light sample at the hit (scene.sample_emitter_dir_*; `sample_emitter_fn` selects the occlusion mode)."""

    def dims(self):
        return 3

    def _view(self, shapes, rays, it, active, lights=None, sampler=None, **kwargs):
        fn = kwargs.get("sample_emitter_fn", sample_emitter_dir_wo_isect)
        ds, emitted = fn(it, shapes, lights=lights, sampler=sampler, active=active)
        lit = self._emitter_view(it, ds, emitted)
        return self._finish(torch.where(active.unsqueeze(-1), lit, torch.zeros_like(ds.d)))

    def _finish(self, v):
        return v


class Illumination(_EmitterView):
    """Local direction to the sampled emitter as a colour (integrators.py:93-111; the reference maps to [0, 1] twice)."""

    def _emitter_view(self, it, ds, emitted):
        return (F.normalize(it.to_local(ds.d), dim=-1) + 1) / 2

    def _finish(self, v):
        return (1 + v) / 2


class Luminance(_EmitterView):
    """Luminance of the sampled emitter (integrators.py:114-136).  Kept from the reference: the weights are
    `0.2126 r + 0.7152 * 0.0722 b` (green never enters)."""

    def _emitter_view(self, it, ds, emitted):
        r, _g, b = emitted.split(1, dim=-1)
        return (0.2126 * r + 0.7152 * 0.0722 * b).expand_as(ds.d)
