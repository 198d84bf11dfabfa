"""D3-type coordination number: sum over neighbours of a counting function of r / (rcov_i + rcov_j)."""
import torch

from .. import storch
from ..batch import real_pairs
from ..data.radii import COV_D3
from .count import dexp_count, exp_count

CUTOFF_D3 = 25.0


def cn_d3(numbers, positions, *, counting_function=exp_count, rcov=None, cutoff=None, **kwargs):
    dd = {"device": positions.device, "dtype": positions.dtype}
    if cutoff is None:
        cutoff = torch.tensor(CUTOFF_D3, **dd)
    if rcov is None:
        rcov = COV_D3(**dd)[numbers]
    if numbers.shape != rcov.shape:
        raise ValueError(f"Shape of covalent radii {rcov.shape} is not consistent with ({numbers.shape}).")
    if numbers.shape != positions.shape[:-1]:
        raise ValueError(f"Shape of positions ({positions.shape[:-1]}) is not consistent with atomic numbers ({numbers.shape}).")
    mask = real_pairs(numbers, mask_diagonal=True)
    eps = torch.tensor(torch.finfo(positions.dtype).eps, **dd)
    distances = torch.where(mask, storch.cdist(positions, positions, p=2), eps)
    rc = rcov.unsqueeze(-2) + rcov.unsqueeze(-1)
    cf = torch.where(mask * (distances <= cutoff), counting_function(distances, rc, **kwargs), torch.tensor(0.0, **dd))
    return torch.sum(cf, dim=-1)


def cn_d3_gradient(numbers, positions, *, dcounting_function=dexp_count, rcov=None, cutoff=None, **kwargs):
    """dCN_i/dR_j as (..., nat, nat, 3)."""
    dd = {"device": positions.device, "dtype": positions.dtype}
    if cutoff is None:
        cutoff = torch.tensor(CUTOFF_D3, **dd)
    if rcov is None:
        rcov = COV_D3(**dd)[numbers]
    mask = real_pairs(numbers, mask_diagonal=True)
    eps = torch.tensor(torch.finfo(positions.dtype).eps, **dd)
    distances = torch.where(mask, storch.cdist(positions, positions, p=2), eps)
    rc = rcov.unsqueeze(-2) + rcov.unsqueeze(-1)
    dcf = torch.where(mask * (distances <= cutoff), dcounting_function(distances, rc, **kwargs), torch.tensor(0.0, **dd))
    rij = positions.unsqueeze(-2) - positions.unsqueeze(-3)
    dcf = (dcf / distances).unsqueeze(-1) * rij  # [i, j] = d cf_ij / d R_i
    # dCN_i/dR_j = -dcf_ij for j != i ; dCN_i/dR_i = sum_j dcf_ij
    out = -dcf
    diag = dcf.sum(-2)
    idx = torch.arange(numbers.shape[-1], device=positions.device)
    out[..., idx, idx, :] = diag
    return out


coordination_number = cn_d3
