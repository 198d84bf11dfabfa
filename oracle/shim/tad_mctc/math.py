"""``einsum``: the real package dispatches to opt_einsum when available; plain torch otherwise."""
import torch


def einsum(equation, *operands, **kwargs):
    kwargs.pop("optimize", None)
    return torch.einsum(equation, *operands)
