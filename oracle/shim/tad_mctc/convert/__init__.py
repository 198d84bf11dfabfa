"""Conversions between python / numpy / torch objects."""
from __future__ import annotations

import numpy as np
import torch
from torch import Tensor

__all__ = ["any_to_tensor", "normalize_device", "numpy_to_tensor", "str_to_device", "symbol_to_number", "number_to_symbol",
           "symmetrize", "tensor_to_numpy", "reshape_fortran"]


def any_to_tensor(x, device=None, dtype=None) -> Tensor:
    if isinstance(x, Tensor):
        return x.to(device=device, dtype=dtype)
    if isinstance(x, np.ndarray):
        return torch.from_numpy(x).to(device=device, dtype=dtype)
    if isinstance(x, str):
        raise ValueError(f"Cannot convert string '{x}' to tensor.")
    if isinstance(x, (bool, int, float, list, tuple)):
        return torch.tensor(x, device=device, dtype=dtype)
    raise TypeError(f"Tensor-incompatible type '{type(x)}'.")


def numpy_to_tensor(x, device=None, dtype=None) -> Tensor:
    return torch.from_numpy(x).to(device=device, dtype=dtype)


def tensor_to_numpy(x: Tensor, dtype=None):
    a = x.detach().cpu().numpy()
    return a if dtype is None else a.astype(dtype)


def str_to_device(s):
    return torch.device(s) if s is not None else None


def normalize_device(device):
    if device is None:
        return torch.tensor(1.0).device
    return torch.device(device)


def symmetrize(x: Tensor, force: bool = False) -> Tensor:
    """(x + x^T)/2 after checking that x is symmetric within 10*eps (unless ``force``)."""
    sym = 0.5 * (x + x.mT)
    if not force:
        tol = torch.finfo(x.dtype).eps * 10
        if not torch.allclose(x, x.mT, atol=tol, rtol=tol):
            raise RuntimeError("Matrix appears to be not symmetric. Use `force=True` to symmetrize anyway.")
    return sym


def reshape_fortran(x: Tensor, shape) -> Tensor:
    if len(x.shape) > 0:
        x = x.permute(*reversed(range(len(x.shape))))
    return x.reshape(*reversed(shape)).permute(*reversed(range(len(shape))))


def symbol_to_number(symbols):
    from ..data.pse import S2Z

    return torch.tensor([S2Z.get(s.title(), 0) for s in symbols])


def number_to_symbol(numbers):
    from ..data.pse import Z2S

    return [Z2S.get(int(z), "X") for z in numbers]
