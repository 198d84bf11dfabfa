import torch


def gaussian_weight(dcn, factor: float = 4.0):
    return torch.exp(-factor * dcn.pow(2))


def weight_references(numbers, cn, reference, weighting_function=gaussian_weight, epsilon=None, **kwargs):
    """Normalised Gaussian weights of every atom's reference systems, (..., nat, 7)."""
    refcn = reference.cn[numbers]
    mask = refcn >= 0
    dcn = refcn - cn.unsqueeze(-1)
    weights = torch.where(mask, weighting_function(dcn, **kwargs), torch.tensor(0.0, device=cn.device, dtype=cn.dtype))
    eps = torch.finfo(cn.dtype).eps if epsilon is None else epsilon
    norms = torch.add(torch.sum(weights, dim=-1), eps)
    return weights / norms.unsqueeze(-1)


def atomic_c6(numbers, weights, reference, chunk_size=None):
    rc6 = reference.c6[numbers.unsqueeze(-1), numbers.unsqueeze(-2)]
    gw = weights.unsqueeze(-1).unsqueeze(-3) * weights.unsqueeze(-2).unsqueeze(-4)
    return torch.sum(torch.sum(gw * rc6, dim=-1), dim=-1)
