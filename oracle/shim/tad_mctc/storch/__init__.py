""""Safe" torch ops: arguments are clamped / shifted away from the singular point by the dtype's eps."""
import torch

from . import linalg
from .distance import cdist
from .linalg import eighb

__all__ = ["cdist", "divide", "eighb", "linalg", "pow", "reciprocal", "sqrt"]


def _eps(x, eps=None):
    if eps is None:
        return torch.tensor(torch.finfo(x.dtype).eps, device=x.device, dtype=x.dtype)
    return torch.as_tensor(eps, device=x.device, dtype=x.dtype)


def sqrt(x, *, eps=None):
    return torch.sqrt(torch.clamp(x, min=_eps(x, eps)))


def divide(x, y, *, eps=None, **kwargs):
    e = _eps(y if torch.is_tensor(y) and y.is_floating_point() else x, eps)
    y = torch.as_tensor(y, device=x.device)
    y_safe = torch.where(y == 0, e.to(y.dtype) if y.is_floating_point() else e, y)
    return torch.divide(x, y_safe, **kwargs)


def reciprocal(x, *, eps=None, **kwargs):
    x_safe = torch.where(x == 0, _eps(x, eps), x)
    return torch.reciprocal(x_safe, **kwargs)


def pow(x, exponent, *, eps=None, **kwargs):
    """x**exponent with zeros shifted by eps (negative and fractional exponents stay finite in value and gradient)."""
    x_safe = torch.where(x == 0, _eps(x, eps), x)
    return torch.pow(x_safe, exponent, **kwargs)
