"""Thin wrappers over torch.autograd / torch.func (only the signatures dxtb imports)."""
from __future__ import annotations

import torch

from . import checks
from .checks import is_batched, is_gradtracking  # noqa: F401

jacrev = torch.func.jacrev
vmap = torch.func.vmap
functorch_jacobian = torch.func.jacrev


def jac(f, argnums: int = 0):
    """Row-by-row Jacobian of ``f`` w.r.t. argument ``argnums`` with torch.autograd.grad (graph kept)."""

    def wrap(*args):
        x = args[argnums]
        y = f(*args)
        flat = y.reshape(-1)
        rows = []
        for i in range(flat.numel()):
            (g,) = torch.autograd.grad(flat[i], x, retain_graph=True, create_graph=True, allow_unused=True)
            rows.append(torch.zeros_like(x) if g is None else g)
        return torch.stack(rows).reshape(*y.shape, *x.shape)

    return wrap


def hessian(f, inputs, argnums: int, is_batched: bool = False):
    def grad_fn(*a):
        x = a[argnums]
        (g,) = torch.autograd.grad(f(*a).sum(), x, create_graph=True)
        return g

    return jac(grad_fn, argnums)(*inputs)
