"""Emitter sampling with optional shadow rays / learned occlusion (pytorch3d/pathtracer/scene.py:290-324)."""
import torch

from .. import ops
from .utils import dir_to_elev_azim


def sample_emitter_dir_wo_isect(it, shapes, lights, sampler, active=True):
    ds, spectrum = lights.sample_direction(it, sampler=sampler, active=active)
    spectrum[~active] = 0
    return ds, spectrum


def _not_blocked(it, shapes, ds, active):
    rays = torch.cat([it.p, ds.d], dim=-1)
    return shapes.intersect_test(rays, max_t=ds.dist.reshape_as(active)[..., None], active=active)


def sample_emitter_dir_w_isect(it, shapes, lights, sampler, active=True):
    ds, spectrum = lights.sample_direction(it, sampler=sampler, active=active)
    nb = _not_blocked(it, shapes, ds, active)
    spectrum[~nb | ~active] = 0
    return ds, spectrum


def sample_emitter_dir_w_learned_occ(it, shapes, lights, sampler, occ, active=True):
    ds, spectrum = lights.sample_direction(it, sampler=sampler, active=active)
    occluded = ~_not_blocked(it, shapes, ds, active)
    occ_in = torch.cat([it.p, dir_to_elev_azim(ds.d)], dim=-1)
    spectrum = torch.where(occluded[..., None], occ(occ_in, out_act=ops.OUT_SIGMOID) * spectrum, spectrum)
    return ds, active[..., None] * spectrum
