"""The fork's point light of `pytorch3d.renderer` as the vis helpers use it (renderer/lighting.py:220-305: upstream's
PointLights plus `scale` and the path tracer's `sample_direction`); `utils.sphere_examples` lights its sphere with it.
Only the path-tracing side is mirrored (the rasteriser's diffuse / specular shading is out of scope)."""
import torch
import torch.nn.functional as F

from ..pathtracer.interaction import DirectionSample


def _rows3(v, device):
    t = v if torch.is_tensor(v) else torch.tensor(v, dtype=torch.float32)
    t = t.to(device=device, dtype=torch.float32)
    if t.dim() == 1:
        t = t.unsqueeze(0)
    if t.shape[-1] != 3:
        raise ValueError("Expected a colour / location of shape (N, 3); got %r" % (tuple(t.shape),))
    return t


class PointLights:
    def __init__(self, ambient_color=((0.5, 0.5, 0.5),), diffuse_color=((0.3, 0.3, 0.3),),
                 specular_color=((0.2, 0.2, 0.2),), location=((0, 1, 0),), device="cpu", scale=1e-2):
        self.device = torch.device(device)
        self.ambient_color = _rows3(ambient_color, self.device)
        self.diffuse_color = _rows3(diffuse_color, self.device)
        self.specular_color = _rows3(specular_color, self.device)
        self.location = _rows3(location, self.device)
        self.intensity = self.ambient_color.clone()          # lighting.py:255: the emitted colour is the ambient colour
        self.scale = scale

    def __len__(self):
        return self.location.shape[0]

    def sample_towards(self, points):
        return F.normalize(self.location - points, dim=-1)

    def sample_direction(self, it, sampler, active=True):
        """lighting.py:289-305: unit direction and distance to the light, spectrum = scale * colour / dist^2 (the
        reciprocal of 1e-7 + dist, as there)."""
        ds = DirectionSample()
        ds.p, ds.n, ds.obj = self.location, 0, self
        ds.delta = torch.tensor(True, device=self.device)
        to_light = ds.p - it.p
        ds.dist = (to_light * to_light).sum(dim=-1, keepdim=True).sqrt()
        inv = (1e-7 + ds.dist).reciprocal()
        ds.d = to_light * inv
        return ds, self.scale * self.intensity * inv * inv
