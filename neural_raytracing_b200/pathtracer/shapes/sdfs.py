"""SphereSDF and the sphere-tracing SDF shape (pytorch3d/pathtracer/shapes/sdfs.py).

The three gradient-free scan loops that are >95 % of a reference step (march :119-131, min scan
:232-249, shadow march :169-180) run in the fused CUDA kernels of libnrt_b200 (persistent,
slot-compacted).  What autograd has to see: sdf(best_pos) goes through the fused MLP forward /
backward kernels; the normals (sdfs.py:184-197, a create_graph autograd in the reference, whose
double backward carries the eikonal and shading losses into the SDF weights) are the forward-mode
analytic Jacobian kernel with its own hand-written reverse pass (nrt_mlp_value_jac_forward /
_backward).  Only the 3 kFLOP/sample smooth-min of the sphere set stays a torch expression."""
import random

import torch
import torch.nn as nn
import torch.nn.functional as F

from .. import utils as U
from ... import ops
from ..interaction import MixedInteraction
from ..neural_blocks import SkipConnMLP


def SPHERE_SDF(p):
    return torch.norm(p, dim=-1) - 1


class _MlpValueJac(torch.autograd.Function):
    """(value, d value / d p) of SphereSDF.shift at p, forward mode, with the hand-written reverse pass into the
    weights.  p itself carries no gradient: in the reference the hit points come out of a no_grad march and
    autograd_diff makes them leaves (sdfs.py:119-131, 186-187)."""

    @staticmethod
    def forward(ctx, module, p, *params):
        pk = module.packed()
        x = p.detach().float().reshape(-1, 3).contiguous()
        ctx.pk = pk
        ctx.tc_prec = module.train_precision()
        ctx.flat_grad = getattr(module, "_flat_grad", None)
        if ctx.tc_prec != "f32":
            # tensor cores: four rows per point through the streamed-weight forward, activations saved as 16-bit tiles
            val, jac, ws = ops.mlp_value_jac_forward_tc(pk, x, prec=ctx.tc_prec)
            ctx.K = x.shape[0]
            ctx.save_for_backward(ws)
            return val, jac
        val, jac, acts = ops.mlp_value_jac_forward(pk, x, save_acts=True)
        ctx.save_for_backward(x, acts)
        return val, jac

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, g_val, g_jac):
        if ctx.tc_prec != "f32":
            ws, = ctx.saved_tensors
            g_params = ops.mlp_value_jac_backward_tc(ctx.pk, ctx.K, ws, g_val.contiguous().float(), g_jac.contiguous().float(),
                                                     prec=ctx.tc_prec, g_params=ctx.flat_grad)
            if ctx.flat_grad is not None:       # accumulated straight into training.FlatParameters' gradient buffer
                return (None, None) + (None,) * (2 * len(ctx.pk.dims))
        else:
            x, acts = ctx.saved_tensors
            g_params = ops.mlp_value_jac_backward(ctx.pk, x, acts, g_val.contiguous().float(), g_jac.contiguous().float())
        gW, gb = ctx.pk.unpack(g_params)
        flat = []
        for w, b in zip(gW, gb):
            flat += [w, b]
        return (None, None) + tuple(flat)


class _SphereSet(torch.autograd.Function):
    """(value, d value / d p) of the sphere set at leaf points p, with the hand-written reverse pass of both outputs
    into centers / radii / tfs (nrt_sphere_set_forward / _backward): the reference's torch expression, its
    autograd.grad(create_graph=True) and the double backward through it (sdfs.py:37-45, 184-197) as two kernels."""

    @staticmethod
    def forward(ctx, centers, radii, tfs, p, want_grad):
        x = p.detach().float().reshape(-1, 3).contiguous()
        val, grad = ops.sphere_set_forward(centers, radii, tfs, x, want_grad=want_grad)
        ctx.save_for_backward(centers, radii, tfs, x)
        ctx.want_grad = want_grad
        if not want_grad:
            grad = val.new_zeros(0)
            ctx.mark_non_differentiable(grad)
        return val, grad

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, g_val, g_grad):
        centers, radii, tfs, x = ctx.saved_tensors
        gc, gr, gt = ops.sphere_set_backward(centers, radii, tfs, x, g_val.contiguous().float(),
                                             g_grad.contiguous().float() if ctx.want_grad else None)
        return gc, gr, gt, None, None


class SphereSDF(nn.Module):
    """Smooth-min of n affinely warped spheres plus a residual MLP (sdfs.py:16-46)."""

    def __init__(self, n=2 << 6, device="cuda"):
        super().__init__()
        self.centers = nn.Parameter(0.3 * torch.rand(n, 3, device=device) - 0.15)
        self.radii = nn.Parameter(0.2 * torch.rand(n, device=device) - 0.1)
        self.tfs = nn.Parameter(torch.zeros(n, 3, 3, device=device))
        self.shift = SkipConnMLP(num_layers=8, hidden_size=128, in_size=3, out=1, device=device, freqs=32,
                                 activation=F.softplus, zero_init=True).to(device)

    def __prepare_scriptable__(self):
        # torch.jit.script(SphereSDF(...)) compiles the TorchScript restatement over the same parameter tensors
        from ..checkpoint import scriptable_view
        return scriptable_view(self)

    def set_center(self, at):
        self.centers = nn.Parameter(at.expand_as(self.centers).clone().detach())

    def transform(self, p):
        tfs = self.tfs + torch.eye(3, device=p.device).unsqueeze(0)
        return torch.einsum("ijk,ibk->ibj", tfs, p.expand(tfs.shape[0], -1, -1))

    def packed(self) -> "ops.PackedSDF":
        return ops.PackedSDF(self.centers, self.radii, self.tfs, self.shift.packed())

    def invalidate_packed(self):
        """See SkipConnMLP.invalidate_packed (the sphere parameters are borrowed, never cached)."""
        self.shift.invalidate_packed()

    def _needs_grad(self, p):
        return torch.is_grad_enabled() and (p.requires_grad or any(q.requires_grad for q in self.parameters()))

    def sphere_set(self, p):
        """Smooth-min of the warped spheres (sdfs.py:37-45, utils.py:385-387) as a torch expression: 3 kFLOP per
        sample against 331 kFLOP for `shift`, differentiable to any order."""
        q = self.transform(p.reshape(-1, 3).unsqueeze(0)) - self.centers.unsqueeze(1)
        sd = q.norm(p=2, dim=-1) - self.radii.unsqueeze(-1)
        return U.smooth_min(sd, k=32.).reshape(p.shape[:-1])

    def forward_reference_ops(self, p):
        out = self.sphere_set(p)
        return out + self.shift.forward_reference_ops(p).reshape_as(out)

    def _fusable_grad(self, p):
        return p.is_cuda and not p.requires_grad and ops.HAS_SDF_VALUE_GRAD and self.shift.in_size == 3 and \
            self.shift.latent_size == 0 and self.shift.init.out_features in (32, 64, 128)

    def value_and_normal(self, p):
        """sdf(p) and d sdf / d p for leaf points p [K,3], both carrying the graph to the parameters: the residual MLP
        through the analytic-Jacobian kernels, the sphere set through its value + gradient kernel and that kernel's
        hand-written reverse pass (_SphereSet)."""
        pl = p.detach().reshape(-1, 3)
        with torch.enable_grad():
            s, n_s = _SphereSet.apply(self.centers, self.radii, self.tfs, pl, True)
            v_m, j_m = _MlpValueJac.apply(self.shift, pl, *self.shift._flat_params())
        return s + v_m[:, 0], n_s + j_m[:, 0, :]

    def precision(self):
        """Arithmetic of the gradient-free SDF evaluations: config.precision when the residual MLP has the shape
        the tensor-core path instantiates (the default 8x128 softplus net), else fp32."""
        return self.shift.precision()

    def forward(self, p):
        if p.is_cuda and not self._needs_grad(p):
            return ops.sdf_eval(self.packed(), p.detach().float(), prec=self.precision())
        if p.is_cuda:
            # first-order graph: the residual MLP through the fused forward / backward kernels (_FusedMLP), the sphere
            # set through its own kernel pair when the points are leaves without gradient (SDF.throughput, sdfs.py:249)
            if not p.requires_grad and self.centers.is_cuda:
                out = _SphereSet.apply(self.centers, self.radii, self.tfs, p, False)[0].reshape(p.shape[:-1])
            else:
                out = self.sphere_set(p)
            return out + self.shift(p).reshape_as(out)
        return self.forward_reference_ops(p)


class SDF:
    """General SDF shape: ray marching, soft-silhouette throughput, normals (sdfs.py:89-249)."""

    def __init__(self, device="cuda", sdf=SPHERE_SDF, epsilon=1e-3, max_steps=32, dist=2.2, **kwargs):
        self.device = torch.device(device)
        self.sdf = sdf
        self.epsilon = epsilon
        self.max_steps = max_steps
        self.dist = dist

    def __len__(self):
        return 1

    # `sdf` is what the caller handed over (and what `torch.jit.save(density_field.sdf, ...)` writes back, dtu.py:159);
    # `_impl` is what is evaluated: the same object, except for a TorchScript SphereSDF archive -- the way the scripts hold
    # their shape (`SDF(sdf=torch.jit.load(path, device))`, dtu.py:93-94) -- which is adopted as a SphereSDF of this package
    # sharing the archive's parameter tensors (checkpoint.adopt_script_sphere_sdf), so that it runs on the fused kernels
    # instead of the generic march over a scripted callable.
    @property
    def sdf(self):
        return self._sdf

    @sdf.setter
    def sdf(self, model):
        self._sdf = model
        self._impl = model
        if isinstance(model, torch.jit.ScriptModule):
            from ..checkpoint import adopt_script_sphere_sdf
            first = next(model.parameters(), None)
            if first is not None and self.device.type == "cuda" and first.device.type == "cpu" and torch.cuda.is_available():
                # colocate.py:65 loads its archive without a device argument; one written from CPU tensors
                # (checkpoint.save_sdf_archive) then sits on the host while the rays are on the GPU: move it to the shape's device
                model.to(self.device)
            adopted = adopt_script_sphere_sdf(model)
            if adopted is not None:
                self._impl = adopted

    def parameters(self):
        return self._impl.parameters()

    # ---- fused / unfused dispatch -----------------------------------------------------------
    def _fused(self):
        """PackedSDF if self._impl is (or wraps) a SphereSDF the kernels can evaluate, else None."""
        s = self._impl
        if isinstance(s, SphereSDF) and s.centers.is_cuda:
            return s.packed()
        return None

    def _march_generic(self, r_o, r_d, max_t):
        # arbitrary callables (edit scripts: SDF(sdf=bend)) cannot be fused: the reference's own loop
        depths = torch.zeros(r_o.shape[:-1] + (1,), device=r_o.device)
        remaining = torch.ones(depths.shape[:-1], dtype=torch.bool, device=r_o.device)
        hit_any = torch.zeros_like(remaining)
        with torch.no_grad():
            for _ in range(self.max_steps):
                remaining = remaining & (depths < max_t).squeeze(-1)
                d = self._impl(r_o + r_d * depths)
                hits = remaining & (d <= self.epsilon)
                hit_any = hit_any | hits
                remaining = remaining & ~hits
                depths = torch.where(remaining.unsqueeze(-1), depths + d.unsqueeze(-1), depths)
        return depths, hit_any

    def intersect(self, rays, max_t=10, active=True, primary: bool = True, fused_hits: bool = False):
        """sdfs.py:102-160.  fused_hits (not in the reference; used by the fused Direct integrator): the per-hit
        quantities -- normal, offset point, local incoming direction -- are produced on the COMPACTED hits by one
        kernel with a registered backward and handed over in `si._hits`; the full-size `si.n / si.frame / si.wi / si.p`
        are then scattered copies without an autograd graph (gradients flow through `_hits` and `raw_normals`)."""
        r_o, r_d = rays.split(3, dim=-1)
        packed = self._fused()
        if packed is not None:
            d, out_active = ops.sphere_trace(packed, rays.detach(), self.epsilon, self.max_steps, float(max_t),
                                             prec=self._impl.precision())
            depths = d.unsqueeze(-1)
        else:
            depths, out_active = self._march_generic(r_o, r_d, max_t)
        p = r_o + depths * r_d
        throughput = 0
        if primary:
            throughput, _best = self.throughput(r_o, r_d)
            throughput = -1000 * throughput
        si = MixedInteraction(p=p, t=depths.squeeze(), obj=self, throughput=throughput)
        normals = torch.zeros_like(p)
        if fused_hits and rays.is_cuda:
            from ..fused_shading import _ShadeGeom
            idx = out_active.reshape(-1).nonzero().squeeze(-1)      # host sync: K = #hits (raw_normals is [K,3])
            si._hits = None
            if idx.numel() > 0:
                rays_hit = rays.detach().reshape(-1, 6)[idx]
                p_hit = p.detach().reshape(-1, 3)[idx]
                raw = self.autograd_diff(p_hit)
                setattr(si, "raw_normals", raw)
                n_k, p_off, wi_k, frame_k = _ShadeGeom.apply(raw, p_hit, rays_hit, self.epsilon * 5)
                si._hits = {"idx": idx, "raw": raw, "n": n_k, "p_off": p_off, "wi": wi_k, "rays": rays_hit}
                with torch.no_grad():
                    normals = normals.reshape(-1, 3).index_copy(0, idx, n_k.detach()).reshape(p.shape)
                    p = p.detach().reshape(-1, 3).index_copy(0, idx, p_off.detach()).reshape(p.shape)
                si.p = p
            with torch.no_grad():
                si.set_normals(normals)
                si.wi = si.to_local(-r_d.detach())
            return si, out_active
        if out_active.any():          # host sync, as in the reference: raw_normals is [K,3], K = #hits
            raw = self.autograd_diff(p[out_active])
            setattr(si, "raw_normals", raw)
            normals[out_active] = F.normalize(raw, eps=1e-6, dim=-1)
            p[out_active] = p[out_active] + normals[out_active] * self.epsilon * 5
        si.set_normals(normals)
        si.wi = si.to_local(-r_d)
        return si, out_active

    def intersect_test(self, rays, max_t=10, active=True):
        packed = self._fused()
        if packed is not None:
            mt = max_t if torch.is_tensor(max_t) else torch.full(rays.shape[:-1], float(max_t), device=rays.device)
            act = active if torch.is_tensor(active) else None
            return ops.shadow_test(packed, rays.detach(), mt.detach().float().reshape(rays.shape[:-1]), self.epsilon,
                                   self.max_steps, active=act, prec=self._impl.precision())
        r_o, r_d = rays.split(3, dim=-1)
        depths = torch.zeros(r_o.shape[:-1] + (1,), device=rays.device) + 1e2 * self.epsilon
        remaining = torch.ones(depths.shape[:-1], dtype=torch.bool, device=rays.device)
        with torch.no_grad():
            for _ in range(self.max_steps):
                d = self._impl(r_o + r_d * depths)
                hits = remaining & (d < self.epsilon)
                depths = torch.where(remaining.unsqueeze(-1), depths + d.unsqueeze(-1), depths)
                remaining = remaining & ~hits
        return (depths >= max_t).squeeze(-1) | remaining

    def autograd_diff(self, p):
        """d sdf / d p at p (sdfs.py:184-197).  Without a graph: the fused analytic-Jacobian kernel.  With gradients
        enabled: the same forward-mode kernel for the residual MLP with its hand-written reverse pass (so eikonal /
        shading losses reach the SDF weights), the sphere set through create_graph autograd.  Callables other than
        SphereSDF (and points that themselves require grad) take the reference's create_graph autograd."""
        s = self._impl
        wants_graph = torch.is_grad_enabled() and (not isinstance(s, SphereSDF) or
                                                   any(q.requires_grad for q in s.parameters()))
        if isinstance(s, SphereSDF) and p.is_cuda and not wants_graph and ops.HAS_SDF_VALUE_GRAD:
            prec = s.precision()
            if prec != "f32":
                # 16-bit inference: the residual MLP's Jacobian on the tensor cores (four rows per point through the
                # streamed-weight kernel the training path uses), the sphere set's gradient from its own kernel; the
                # fp32 analytic-Jacobian kernel was 3.0 of 40 ms of the 512x512 colocate frame
                x = p.detach().float().reshape(-1, 3).contiguous()
                _v, g_s = ops.sphere_set_forward(s.centers, s.radii, s.tfs, x, want_grad=True)
                _vm, jac, _ws = ops.mlp_value_jac_forward_tc(s.shift.packed(), x, prec=prec)
                return (g_s + jac[:, 0, :]).reshape(p.shape)
            return ops.sdf_value_grad(s.packed(), p.detach())[1]
        if isinstance(s, SphereSDF) and s._fusable_grad(p):
            return s.value_and_normal(p)[1].reshape(p.shape)
        with torch.enable_grad():
            if not p.requires_grad:
                p = p.requires_grad_()
            out = s.forward_reference_ops(p) if isinstance(s, SphereSDF) else s(p)
            n, = torch.autograd.grad(inputs=p, outputs=out, grad_outputs=torch.ones_like(out), create_graph=True,
                                     retain_graph=True, only_inputs=True)
        return n

    def throughput(self, r_o_local, d):
        """Minimum of the SDF along the ray (sdfs.py:232-249): scan without grad, then one
        differentiable evaluation at the argmin."""
        n = 128
        max_t = getattr(self, "dist", 2.2) + random.random() * (2 / n)
        step = max_t / n
        packed = self._fused()
        if packed is not None:
            rays = torch.cat([r_o_local.expand_as(d), d], dim=-1).detach()
            _idx, best_pos, _mv = ops.min_scan(packed, rays, step, n, prec=self._impl.precision())
        else:
            with torch.no_grad():
                sd = self._impl(r_o_local).squeeze(-1)
                cur, idxs = sd, torch.zeros_like(sd, dtype=torch.long)
                for i in range(n):
                    sd = self._impl(r_o_local + (step * (i + 1)) * d).squeeze(-1)
                    idxs = torch.where(sd < cur, i + 1, idxs)
                    cur = torch.minimum(cur, sd)
            best_pos = r_o_local + idxs.unsqueeze(-1).unsqueeze(-1) * step * d
        return self._impl(best_pos), best_pos
