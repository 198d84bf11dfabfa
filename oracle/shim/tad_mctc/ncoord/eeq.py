"""EEQ coordination number: erf counting, smoothly capped at cn_max."""
import math

import torch

from .. import storch
from ..batch import real_pairs
from ..data.radii import COV_D3
from .count import erf_count

CUTOFF_EEQ = 25.0
CUTOFF_EEQ_MAX = 8.0
KCN_EEQ = 7.5


def cut_coordination_number(cn, cn_max=CUTOFF_EEQ_MAX):
    cn_max = torch.as_tensor(cn_max, device=cn.device, dtype=cn.dtype)
    return torch.log(1.0 + torch.exp(cn_max)) - torch.log(1.0 + torch.exp(cn_max - cn))


def cn_eeq(numbers, positions, *, counting_function=erf_count, rcov=None, cutoff=None, cn_max=CUTOFF_EEQ_MAX, **kwargs):
    dd = {"device": positions.device, "dtype": positions.dtype}
    if cutoff is None:
        cutoff = torch.tensor(CUTOFF_EEQ, **dd)
    if rcov is None:
        rcov = COV_D3(**dd)[numbers]
    mask = real_pairs(numbers, mask_diagonal=True)
    eps = torch.tensor(torch.finfo(positions.dtype).eps, **dd)
    distances = torch.where(mask, storch.cdist(positions, positions, p=2), eps)
    rc = rcov.unsqueeze(-2) + rcov.unsqueeze(-1)
    cf = torch.where(mask * (distances <= cutoff), counting_function(distances, rc, **kwargs), torch.tensor(0.0, **dd))
    cn = torch.sum(cf, dim=-1)
    return cn if cn_max is None else cut_coordination_number(cn, cn_max)
