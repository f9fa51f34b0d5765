"""Fused Direct-lighting path on the compacted hit rays (SURVEY 8b(8) `shade_direct`; rows a8-a16).

The reference's Direct.sample (integrators/integrators.py:156-206) shades ALL R rays with ~150 elementwise torch launches
around its MLP evaluations and masks the misses afterwards (bsdfs.py:521-525, lights.py:109, scene.py:290-324).  Here the
hits are compacted once (K = #hits, the one data-dependent host sync that `raw_normals [K,3]` needs anyway, sdfs.py:152-159)
and everything between the MLP evaluations runs as three elementwise CUDA stages on [K,...] arrays -- geom, light, blend
(csrc/nrt_shade_direct.cu) -- each with a hand-written backward kernel registered below; the MLPs (LightField, occlusion,
sp_var, NeuralBSDF children) are evaluated on the K hits only.  Masked outputs of the reference are exactly 0 on the
misses, so the image is unchanged; `it.normalized_weights`, which the reference returns for all rays (colocate.py:104-105
takes a std over all of it), is completed lazily for the misses the first time it is read.

Falls back (returns None) for anything it does not cover: BSDFs other than ComposeSpatialVarying of
NeuralBSDF / Diffuse / (one) Conductor, lights other than PointLights / LightField, callables it cannot classify."""
import torch
import torch.nn as nn
import torch.nn.functional as F

from .. import ops
from .neural_blocks import SkipConnMLP


# ---- autograd wrappers of the three stages -----------------------------------------------------------------------
class _ShadeGeom(torch.autograd.Function):
    @staticmethod
    def forward(ctx, raw_n, p_hit, rays_hit, eps5):
        rn = raw_n.detach().float().contiguous()
        rh = rays_hit.detach().float().contiguous()
        n, p_off, wi, frame = ops.shade_geom_forward(rn, p_hit.detach().float().contiguous(), rh, eps5)
        ctx.save_for_backward(rn, rh)
        ctx.eps5 = eps5
        ctx.mark_non_differentiable(frame)
        return n, p_off, wi, frame

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, g_n, g_p_off, g_wi, _g_frame):
        rn, rh = ctx.saved_tensors
        c = lambda g: None if g is None else g.contiguous().float()
        return ops.shade_geom_backward(rn, rh, ctx.eps5, c(g_n), c(g_p_off), c(g_wi)), None, None, None


class _ShadeLight(torch.autograd.Function):
    """(n, wi, p_off | v, light parameters) -> d, dist, wo, rusin, e, elaz."""

    @staticmethod
    def forward(ctx, mode, n, wi, p_off, v, location, amp, coef, view_of_hit, sig_color, want_elaz):
        det = lambda t: None if t is None else t.detach().float().contiguous()
        n_, wi_, p_, v_ = det(n), det(wi), det(p_off), det(v)
        loc_, amp_, coef_, sig_ = det(location), det(amp), det(coef), det(sig_color)
        d, dist, wo, ru, e, elaz = ops.shade_light_forward(mode, n_, wi_, p_, loc_, amp_, coef_, view_of_hit, v_, sig_, want_elaz)
        ctx.mode = mode
        ctx.view = view_of_hit
        ctx.save_for_backward(*[t if t is not None else n_.new_empty(0) for t in (n_, wi_, p_, v_, loc_, amp_, coef_, sig_)])
        ctx.has = [t is not None for t in (n_, wi_, p_, v_, loc_, amp_, coef_, sig_)]
        ctx.mark_non_differentiable(d, dist)     # the shadow march is gradient-free (sdfs.py:169) and dist only feeds it
        if elaz is None:
            elaz = n_.new_empty(0)
        return d, dist, wo, ru, e, elaz

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, _g_d, _g_dist, g_wo, g_ru, g_e, g_elaz):
        n_, wi_, p_, v_, loc_, amp_, coef_, sig_ = [t if h else None for t, h in zip(ctx.saved_tensors, ctx.has)]
        c = lambda g: None if (g is None or g.numel() == 0) else g.contiguous().float()
        g_n, g_wi, g_pv, g_amp, g_coef, g_sig = ops.shade_light_backward(
            ctx.mode, n_, wi_, p_, c(g_wo), c(g_ru), c(g_e), c(g_elaz), loc_, amp_, coef_, ctx.view, v_, sig_)
        point = ctx.mode == ops.LIGHT_POINT
        return (None, g_n, g_wi, g_pv if point else None, None if point else g_pv, None, g_amp, g_coef, None, g_sig, None)


class _ShadeBlend(torch.autograd.Function):
    @staticmethod
    def forward(ctx, cfg, logits, neural_raw, wi, wo, e, refl, cond_spec, cond_eta, inv_samples):
        det = lambda t: None if t is None else t.detach().float().contiguous()
        args = [det(t) for t in (logits, neural_raw, wi, wo, e, refl, cond_spec, cond_eta)]
        out = ops.shade_blend_forward(cfg[0], cfg[1], cfg[2], *args, inv_samples)
        ctx.cfg, ctx.inv = cfg, inv_samples
        ctx.has = [t is not None for t in args]
        ctx.save_for_backward(*[t if t is not None else args[0].new_empty(0) for t in args])
        return out

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, g_out):
        args = [t if h else None for t, h in zip(ctx.saved_tensors, ctx.has)]
        g = ops.shade_blend_backward(ctx.cfg[0], ctx.cfg[1], ctx.cfg[2], *args, ctx.inv, g_out.contiguous().float())
        g_logits, g_neural, g_wi, g_wo, g_e, g_refl, g_cs, g_ce = g
        return None, g_logits, g_neural, g_wi, g_wo, g_e, g_refl, g_cs, g_ce, None


# ---- classification of the user's modules --------------------------------------------------------------------------
def _act_id(fn):
    if fn is torch.sigmoid or isinstance(fn, nn.Sigmoid) or fn is F.sigmoid:
        return 0
    if isinstance(fn, nn.Softplus) and fn.beta == 1 and fn.threshold == 20:
        return 1
    if fn is F.softplus:
        return 1
    if getattr(fn, "__name__", "") == "identity" or isinstance(fn, nn.Identity):
        return 2
    return None


def _pre_id(fn):
    name = getattr(fn, "__name__", "")
    if name == "identity" or isinstance(fn, nn.Identity):
        return 0
    if name == "identity_div_pi":
        return 1
    a = _act_id(fn)
    if a == 1:
        return 2
    if a == 0:
        return 3
    return None


def blend_config(bsdf):
    """(kinds, neural_act, diffuse_pre, neural children, diffuse children, conductor or None) of a ComposeSpatialVarying
    the fused blend covers, else None."""
    from .bsdf.bsdfs import ComposeSpatialVarying, Conductor, Diffuse, NeuralBSDF
    if type(bsdf) is not ComposeSpatialVarying or not isinstance(bsdf.sp_var_fn, SkipConnMLP) or len(bsdf.bsdfs) > 16:
        return None
    kinds, neural, diffuse, conductor = [], [], [], None
    n_act = d_pre = None
    for b in bsdf.bsdfs:
        if type(b) is NeuralBSDF and isinstance(b.mlp, SkipConnMLP) and b.mlp.in_size == 3 and b.mlp.out.out_features == 3:
            a = _act_id(b.act)
            if a is None or (n_act is not None and a != n_act):
                return None
            n_act = a
            kinds.append(ops.BSDF_NEURAL); neural.append(b)
        elif type(b) is Diffuse:
            a = _pre_id(b.preproc)
            if a is None or (d_pre is not None and a != d_pre):
                return None
            d_pre = a
            kinds.append(ops.BSDF_DIFFUSE); diffuse.append(b)
        elif type(b) is Conductor and conductor is None:
            kinds.append(ops.BSDF_CONDUCTOR); conductor = b
        else:
            return None
    return kinds, (n_act or 0), (d_pre or 0), neural, diffuse, conductor


def supported(bsdf, lights, w_isect):
    from .lights.lights import LightField, PointLights
    if blend_config(bsdf) is None:
        return False
    if type(lights) not in (PointLights, LightField):
        return False
    if type(lights) is LightField and getattr(lights, "preproc", None) is not None and \
            getattr(lights.preproc, "__name__", "") != "identity":
        return False
    if w_isect in (None, False):
        return True
    # shadow rays need the distance to the light: the light field has none (the reference fails there too, scene.py:297)
    return type(lights) is PointLights and (w_isect is True or isinstance(w_isect, SkipConnMLP))


# ---- the fused path -------------------------------------------------------------------------------------------------
def shade_direct(shapes, rays, it, active, bsdf, lights, w_isect, emitter_samples=1):
    """result [N,W,H,B,3] of Direct.sample's emitter loop for an interaction that carries compacted hits (`it._hits`,
    produced by SDF.intersect(..., fused_hits=True))."""
    from .lights.lights import PointLights
    h = it._hits
    result = torch.zeros(*rays.shape[:-1], 3, device=rays.device)
    if h is None:
        return result
    kinds, n_act, d_pre, neural, diffuse, conductor = blend_config(bsdf)
    idx, n, p_off, wi = h["idx"], h["n"], h["p_off"], h["wi"]
    dev = rays.device
    want_elaz = isinstance(w_isect, SkipConnMLP)
    if type(lights) is PointLights:
        # lights.py:93: location [n,1,1,1,3] broadcasts against it.p [N,W,H,B,3]: one light per view, or one for all
        n_loc = lights.location.shape[0]
        if n_loc == rays.shape[0] and n_loc > 1:
            view = (idx // (rays[0].numel() // 6)).to(torch.int32)
        elif n_loc == 1:
            view = None
        else:
            raise ValueError("PointLights.location has %d rows for %d views" % (n_loc, rays.shape[0]))
        color = lights.intensity.to(dev)
        amp = lights.scale * F.normalize(color, dim=-1)                               # lights.py:104 (0-dim CPU leaf x CUDA)
        coef = torch.stack([lights.const.clamp(min=1e-6), lights.linear.clamp(min=1e-6), lights.square.clamp(min=1e-6)]).to(dev)
        d, dist, wo, rusin, e, elaz = _ShadeLight.apply(ops.LIGHT_POINT, n, wi, p_off, None, lights.location.to(dev), amp, coef,
                                                        view, None, want_elaz)
    else:
        v = lights.light_field_approx(p_off)                                          # lights.py:181, on the hits only
        sig = lights.color.sigmoid()
        d, dist, wo, rusin, e, elaz = _ShadeLight.apply(ops.LIGHT_FIELD, n, wi, p_off, v, None, None, None, None, sig, want_elaz)
    # ---- occlusion (scene.py:290-324) ----
    if w_isect is True or want_elaz:
        with torch.no_grad():
            nb = shapes.intersect_test(torch.cat([p_off.detach(), d], dim=-1), max_t=dist, active=None)
        if w_isect is True:
            e = e * nb.unsqueeze(-1).float()
        else:
            occ = w_isect(torch.cat([p_off, elaz], dim=-1), out_act=ops.OUT_SIGMOID)
            e = torch.where(nb.unsqueeze(-1), e, occ * e)
    # ---- BSDF: MLPs on the hits, then one blend kernel ----
    logits = bsdf.sp_var_fn(bsdf.preprocess(p_off)).reshape(p_off.shape[0], len(kinds))
    raws = torch.stack([b.mlp(rusin) for b in neural]) if neural else None
    refl = torch.stack([b.reflectance.to(dev) for b in diffuse]) if diffuse else None
    cond_spec = cond_eta = None
    if conductor is not None:
        cond_spec = conductor.act(conductor.specular).to(dev).reshape(3)
        cond_eta = F.softplus(conductor.eta).to(dev).reshape(1)
    out = _ShadeBlend.apply((tuple(kinds), n_act, d_pre), logits, raws, wi, wo, e, refl, cond_spec, cond_eta, 1.0 / emitter_samples)
    if emitter_samples != 1:
        out = out * emitter_samples          # integrators.py:171-187: the same deterministic sample, accumulated n times
    result = result.reshape(-1, 3).index_copy(0, idx, out).reshape(result.shape)
    it._lazy_weights = (bsdf, logits, idx, torch.is_grad_enabled())
    return result
