import torch

from . import defaults


def rational_damping(order, distances, qq, param, **kwargs):
    a1 = param.get("a1", torch.tensor(defaults.A1, device=distances.device, dtype=distances.dtype))
    a2 = param.get("a2", torch.tensor(defaults.A2, device=distances.device, dtype=distances.dtype))
    return 1.0 / (distances.pow(order) + (a1 * torch.sqrt(qq) + a2).pow(order))
