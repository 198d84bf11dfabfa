import torch


def is_linear(numbers, positions, atol: float = 1e-8, rtol: float = 1e-5):
    """True if all atoms lie on one line (rank of the centred coordinates <= 1)."""
    mask = numbers != 0
    p = positions - (positions * mask.unsqueeze(-1)).sum(-2, keepdim=True) / mask.sum(-1, keepdim=True).unsqueeze(-1)
    s = torch.linalg.svdvals(p * mask.unsqueeze(-1))
    return s[..., 1] <= atol + rtol * s[..., 0]
