"""BSDFs used by the hot path (pytorch3d/pathtracer/bsdf/bsdfs.py): Diffuse, Conductor,
NeuralBSDF and the spatially varying composition; `eval_and_pdf` for the Direct integrator, `sample` for the
multi-bounce Path integrator."""
import math
from dataclasses import dataclass
from itertools import chain

import torch
import torch.nn as nn
import torch.nn.functional as F

from ... import ops
from ..neural_blocks import SkipConnMLP
from ..utils import param_rusin2
from ..warps import square_to_cos_hemisphere


def identity(x):
    return x


def identity_div_pi(x):
    return x / math.pi


def square_to_cos_hemisphere_pdf(v):
    """warps.py: cosine-hemisphere pdf = cos(theta) / pi."""
    return v[..., 2] / math.pi


def local_reflect(v):
    """bsdfs.py:126-129: mirror about the local normal (0, 0, 1)."""
    return torch.cat([-v[..., 0:1], -v[..., 1:2], v[..., 2:3]], dim=-1)


@dataclass
class BSDFSample:
    """One sampled bounce (bsdfs.py:22-63): local outgoing direction and its pdf."""
    pdf: torch.Tensor = 0
    wo: torch.Tensor = 0
    eta: torch.Tensor = 1

    @classmethod
    def zeros_like(cls, like):
        return cls(pdf=torch.zeros(like.shape[:-1], device=like.device), wo=torch.zeros_like(like), eta=1)

    @staticmethod
    def compose(samples, k, selections):
        """Per point, the sample of the selected child; its pdf times the selection weight (bsdfs.py:44-63)."""
        nb = k.shape[-1]
        rows = torch.arange(selections.shape[0], device=selections.device)
        pdfs = torch.stack([s.pdf for s in samples], dim=-1).reshape(-1, nb)
        pdf = (pdfs[rows, selections] * k.reshape(-1, nb)[rows, selections]).reshape_as(samples[0].pdf)
        wos = torch.stack([s.wo for s in samples], dim=-1).reshape(-1, 3, nb)
        wo = wos[rows, :, selections].reshape_as(samples[0].wo)
        return BSDFSample(pdf=pdf, wo=F.normalize(wo, dim=-1), eta=samples[0].eta)


def _cos_hemisphere_sample(it, sampler):
    wo = F.normalize(square_to_cos_hemisphere(sampler.sample(it.shape()[:-1] + (2,), device=it.device())), dim=-1)
    return BSDFSample(pdf=square_to_cos_hemisphere_pdf(wo), wo=wo, eta=1.0)


class BSDF(nn.Module):
    def sample(self, it, sampler, active=True):
        raise NotImplementedError()

    def eval_and_pdf(self, it, wo, active=True):
        raise NotImplementedError()

    def joint_eval_pdf(self, it, wo, active=True):
        spectrum, pdf = self.eval_and_pdf(it, wo, active)
        return torch.cat([spectrum, pdf.reshape(spectrum.shape[:-1] + (1,))], dim=-1)

    def eval(self, it, wo, active=True):
        return self.eval_and_pdf(it, wo, active)[0]

    def pdf(self, it, wo, active=True):
        return self.eval_and_pdf(it, wo, active)[1]


class Diffuse(BSDF):
    """bsdfs.py:78-118: preproc(cos(theta_o) * reflectance)."""

    def __init__(self, reflectance=[0.25, 0.2, 0.7], preprocess=identity_div_pi, device="cuda"):
        super().__init__()
        if type(reflectance) == list:
            self.reflectance = torch.tensor(reflectance, device=device, requires_grad=True)
        else:
            self.reflectance = reflectance
        self.preproc = preprocess

    def parameters(self):
        return [self.reflectance]

    def random(self):
        self.reflectance = torch.rand_like(self.reflectance, requires_grad=True)
        return self

    def sample(self, it, sampler, active=True):
        """bsdfs.py:88-106: cosine-hemisphere direction; the spectrum is preproc(reflectance) on every point (the
        reference leaves the masking of inactive points commented out)."""
        active = (it.wi[..., 2] > 0) & active
        if not active.any():
            return BSDFSample.zeros_like(it.p), torch.zeros_like(it.p)
        bs = _cos_hemisphere_sample(it, sampler)
        return bs, self.preproc(self.reflectance).expand(*it.shape()).clone()

    def eval_and_pdf(self, it, wo, active=True):
        spectrum = self.preproc(wo[..., 2].unsqueeze(-1) * self.reflectance)
        return spectrum, square_to_cos_hemisphere_pdf(wo)


def fresnel_conductor(cos_t, eta_r, eta_i):
    """bsdfs.py:327-341."""
    ct2 = cos_t * cos_t
    st2 = (1 - ct2).clamp(min=1e-10)
    st4 = st2 * st2
    tmp = eta_r * eta_r - eta_i * eta_i - st2
    a2b2 = (tmp * tmp + 4 * eta_i * eta_i * eta_r * eta_r).clamp(min=1e-10).sqrt()
    a = (0.5 * (a2b2 + tmp)).clamp(min=1e-10).sqrt()
    t1 = a2b2 + ct2
    t2 = 2 * cos_t * a
    r_s = (t1 - t2) / (t1 + t2)
    t3 = a2b2 * ct2 + st4
    t4 = t2 * st2
    r_p = r_s * (t3 - t4) / (t3 + t4)
    return 0.5 * (r_s + r_p)


class Conductor(BSDF):
    """bsdfs.py:345-402: mirror lobe gated by dot(reflect(wi), wo) > 0.94 times a Fresnel term."""

    def __init__(self, specular=[1., 1., 1.], eta: float = 1.3, k: float = 1, device="cuda", activation=torch.sigmoid):
        super().__init__()
        self.eta = torch.tensor(eta, requires_grad=True, dtype=torch.float)
        self.k = torch.tensor(k, requires_grad=True, dtype=torch.float)
        self.specular = torch.tensor(specular, device=device, requires_grad=True) if type(specular) == list else specular
        self.act = activation

    def random(self):
        self.specular = torch.rand_like(self.specular, requires_grad=True)
        return self

    def parameters(self):
        return [self.eta, self.k, self.specular]

    def sample(self, it, sampler, active=True):
        """bsdfs.py:390-400: the mirror direction with pdf 1, specular * Fresnel on the front-facing active points.
        (The reference calls the two-argument `reflect(n, v)` with one argument there and raises; this is the
        local-frame mirror `local_reflect` that the call evidently means.)"""
        cos_i = it.wi[..., 2]
        active = (cos_i > 0) & active
        bs = BSDFSample(pdf=torch.ones_like(active), wo=local_reflect(it.wi), eta=1)
        f = fresnel_conductor(cos_i, self.eta.to(it.wi.device), self.k.to(it.wi.device)).unsqueeze(-1)
        spectrum = torch.where(active.unsqueeze(-1), self.specular * f, torch.zeros_like(it.p))
        return bs, spectrum

    def eval_and_pdf(self, it, wo, active=True):
        wi = it.wi
        refl = torch.cat([-wi[..., 0:1], -wi[..., 1:2], wi[..., 2:3]], dim=-1)
        thresh = (refl * wo).sum(dim=-1, keepdim=True) > 0.94
        fresnel = fresnel_conductor(wi[..., 2], F.softplus(self.eta).to(wi.device), 0.0).reshape_as(thresh)
        spectrum = torch.where(thresh, fresnel * self.act(self.specular), torch.zeros_like(it.p))
        pdf = thresh.squeeze(-1).float()
        if torch.is_tensor(active):
            spectrum = torch.where(active.unsqueeze(-1), spectrum, torch.zeros_like(spectrum))
        return spectrum, pdf


class NeuralBSDF(BSDF):
    """bsdfs.py:613-645: act(MLP(rusinkiewicz(wi, wo)))."""

    def __init__(self, activation=torch.sigmoid, device="cuda"):
        super().__init__()
        self.mlp = SkipConnMLP(in_size=3, out=3, num_layers=6, hidden_size=96, freqs=64, device=device).to(device)
        self.act = activation

    def parameters(self):
        return chain(self.mlp.parameters())

    def random(self):
        return self

    def sample(self, it, sampler, active=True):
        """bsdfs.py:625-633: cosine-hemisphere direction, spectrum = act(MLP(rusinkiewicz(wi, wo))) on every point."""
        bs = _cos_hemisphere_sample(it, sampler)
        return bs, self.act(self.mlp(param_rusin2(it.wi, bs.wo)))

    def eval_and_pdf(self, it, wo, active=True, rusin=None):
        coords = param_rusin2(it.wi, wo) if rusin is None else rusin
        spectrum = self.act(self.mlp(coords))
        return spectrum, torch.ones(spectrum.shape[:-1], device=spectrum.device)

    def zero(self):
        class Zero(nn.Module):
            def forward(self, x):
                return torch.zeros_like(x)
        self.mlp = Zero()


class ComposeSpatialVarying(BSDF):
    """bsdfs.py:482-540: per-point sigmoid weights (a 16x256 MLP on p) blending the child BSDFs."""

    def __init__(self, bsdfs, spatial_varying_fn=None, device="cuda"):
        super().__init__()
        self.bsdfs = bsdfs
        if spatial_varying_fn is None:
            self.sp_var_fn = SkipConnMLP(num_layers=16, hidden_size=256, freqs=128, sigma=2 << 6, in_size=3,
                                         out=len(bsdfs), device=device, xavier_init=True).to(device)
        else:
            self.sp_var_fn = spatial_varying_fn
        self.preprocess = identity

    def normalized_weights(self, p, it):
        w = self.sp_var_fn(self.preprocess(p)).reshape(p.shape[:-1] + (len(self.bsdfs),))
        setattr(it, "nonnormalized_weights", w)
        return w.sigmoid()

    def sample(self, it, sampler, active=True):
        """bsdfs.py:500-513: every child proposes a bounce, one child per point is drawn from the (un-normalised,
        sigmoid) spatial weights with torch.multinomial, and that child's direction / spectrum is kept."""
        samples, spectrums = zip(*[b.sample(it, sampler, active) for b in self.bsdfs])
        nb = len(self.bsdfs)
        k = self.normalized_weights(it.p, it)
        selections = torch.multinomial(k.reshape(-1, nb), num_samples=1).squeeze(-1)
        stacked = torch.stack(spectrums, dim=-1).reshape(-1, 3, nb)
        rows = torch.arange(stacked.shape[0], device=stacked.device)
        spectrum = stacked[rows, :, selections].reshape_as(it.p)
        return BSDFSample.compose(samples, k, selections), spectrum

    def eval_and_pdf(self, it, wo, active=True):
        k = self.normalized_weights(it.p, it)
        # the Rusinkiewicz coordinates are shared by every NeuralBSDF child (the reference recomputes
        # them per child; same values)
        rusin = None
        parts = []
        for b in self.bsdfs:
            if isinstance(b, NeuralBSDF):
                if rusin is None:
                    rusin = param_rusin2(it.wi, wo)
                s, pdf = b.eval_and_pdf(it, wo, active, rusin=rusin)
            else:
                s, pdf = b.eval_and_pdf(it, wo, active)
            parts.append(torch.cat([s, pdf.reshape(s.shape[:-1] + (1,))], dim=-1))
        spec_pdf = torch.stack(parts, dim=-1)
        setattr(it, "normalized_weights", k)
        spec_pdf = torch.where(active[..., None, None], spec_pdf * k.unsqueeze(-2), torch.zeros_like(spec_pdf))
        spectrum, pdf = spec_pdf.sum(dim=-1).split([3, 1], dim=-1)
        return spectrum, pdf.squeeze(-1)

    def parameters(self):
        return chain(self.own_parameters(), self.child_parameters())

    def own_parameters(self):
        return self.sp_var_fn.parameters()

    def child_parameters(self):
        return chain(*[b.parameters() for b in self.bsdfs])
