"""Padding / masking helpers for batches of differently sized systems (zero padding is the convention)."""
from __future__ import annotations

import torch
from torch import Tensor

__all__ = ["deflate", "eye", "index", "pack", "real_atoms", "real_pairs", "real_triples", "unpack"]


def pack(tensors, axis: int = 0, value=0, size=None, return_mask: bool = False):
    """Stack tensors of different shapes into one, padding with ``value``; the new batch dimension is ``axis``."""
    if isinstance(tensors, Tensor):
        return tensors
    count = len(tensors)
    dev, dt = tensors[0].device, tensors[0].dtype
    if size is None:
        size = torch.tensor([list(t.shape) for t in tensors]).max(0).values.tolist() if tensors[0].ndim > 0 else []
    padded = torch.full((count, *size), value, dtype=dt, device=dev)
    mask = torch.zeros((count, *size), dtype=torch.bool, device=dev) if return_mask else None
    for n, src in enumerate(tensors):
        sl = (n, *[slice(0, s) for s in src.shape])
        padded[sl] = src
        if mask is not None:
            mask[sl] = True
    if axis != 0:
        ax = padded.dim() + axis if axis < 0 else axis
        order = list(range(1, padded.dim()))
        order.insert(ax, 0)
        padded = padded.permute(order)
        if mask is not None:
            mask = mask.permute(order)
    return (padded, mask) if return_mask else padded


def unpack(tensor: Tensor, value=0, axis: int = 0):
    return tuple(deflate(t, value) for t in tensor.movedim(axis, 0))


def deflate(tensor: Tensor, value=0, axis: int | None = None) -> Tensor:
    """Strip trailing padding: along every dimension the trailing slices that only hold ``value`` are cut."""
    if isinstance(value, float) and value != value:  # NaN padding
        mask = torch.isnan(tensor)
    else:
        mask = tensor == value
    if axis is not None:
        mask = mask.all(axis)
    slices = []
    nd = mask.ndim
    for d in range(nd):
        other = tuple(i for i in range(nd) if i != d)
        full = mask.all(other) if other else mask  # padding-only slices along dimension d
        keep = int(full.numel())
        while keep > 0 and bool(full[keep - 1]):
            keep -= 1
        slices.append(slice(None, keep))
    if axis is not None:
        slices.insert(axis if axis >= 0 else tensor.ndim + axis, slice(None))
    return tensor[tuple(slices)]


def real_atoms(numbers: Tensor) -> Tensor:
    return numbers != 0


def real_pairs(numbers: Tensor, mask_diagonal: bool = True) -> Tensor:
    real = real_atoms(numbers)
    mask = real.unsqueeze(-2) * real.unsqueeze(-1)
    if mask_diagonal:
        mask = mask * ~torch.diag_embed(torch.ones_like(real))
    return mask


def real_triples(numbers: Tensor, mask_diagonal: bool = True, mask_self: bool = True) -> Tensor:
    real = real_pairs(numbers, mask_diagonal=False)
    mask = real.unsqueeze(-3) * real.unsqueeze(-2) * real.unsqueeze(-1)
    if mask_diagonal:
        mask = mask * ~torch.diag_embed(torch.ones_like(real))
    if mask_self:
        mask = mask * ~torch.diag_embed(torch.ones_like(real), offset=0, dim1=-3, dim2=-2)
        mask = mask * ~torch.diag_embed(torch.ones_like(real), offset=0, dim1=-3, dim2=-1)
    return mask


def eye(size, value: float = 1.0, device=None, dtype=None) -> Tensor:
    """Batched identity: ``size`` = (..., n, n)."""
    out = torch.zeros(*size, device=device, dtype=dtype)
    out.diagonal(dim1=-2, dim2=-1).fill_(value)
    return out


def index(inp: Tensor, idx: Tensor) -> Tensor:
    """inp[idx] for a (possibly batched) 1-D parameter table and an index tensor."""
    if idx.ndim > 1 and inp.ndim > 1:
        return torch.stack([inp[b][idx[b]] for b in range(idx.shape[0])])
    return inp[idx]
