import torch


def cdist(x, y=None, p=2):
    """Pairwise p-norm distances, sqrt(clamp(sum |xi-yj|^p, min=eps)): finite gradient on the diagonal."""
    if y is None:
        y = x
    eps = torch.tensor(torch.finfo(x.dtype).eps, device=x.device, dtype=x.dtype)
    diff = torch.abs(x.unsqueeze(-2) - y.unsqueeze(-3))
    if p == 2:
        d = torch.einsum("...ijk,...ijk->...ij", diff, diff)
    else:
        d = torch.sum(torch.pow(diff, p), -1)
    return torch.pow(torch.clamp(d, min=eps), 1.0 / p)
