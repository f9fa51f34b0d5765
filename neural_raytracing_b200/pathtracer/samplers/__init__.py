import torch


class Sampler:
    """Uniform [0,1) sampler (pytorch3d/pathtracer/samplers/samplers.py:14-20)."""

    def __init__(self, device="cuda"):
        self.device = device

    def sample(self, shape, device=None):
        return torch.rand(shape, device=device if device is not None else self.device)
