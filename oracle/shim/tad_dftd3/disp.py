import torch

from tad_mctc import storch
from tad_mctc.batch import real_pairs
from tad_mctc.data.radii import COV_D3
from tad_mctc.ncoord import cn_d3, exp_count

from . import data, defaults
from .damping import rational_damping
from .model import atomic_c6, gaussian_weight, weight_references
from .reference import Reference


def dispersion2(numbers, positions, param, c6, r4r2, damping_function=rational_damping, cutoff=None, **kwargs):
    dd = {"device": positions.device, "dtype": positions.dtype}
    if cutoff is None:
        cutoff = torch.tensor(defaults.D3_DISP_CUTOFF, **dd)
    mask = real_pairs(numbers, mask_diagonal=True)
    eps = torch.tensor(torch.finfo(positions.dtype).eps, **dd)
    zero = torch.tensor(0.0, **dd)
    distances = torch.where(mask, storch.cdist(positions, positions, p=2), eps)
    qq = 3 * r4r2.unsqueeze(-1) * r4r2.unsqueeze(-2)
    ok = mask * (distances <= cutoff)
    t6 = torch.where(ok, damping_function(6, distances, qq, param, **kwargs), zero)
    t8 = torch.where(ok, damping_function(8, distances, qq, param, **kwargs), zero)
    e6 = -0.5 * torch.sum(c6 * t6, dim=-1)
    e8 = -0.5 * torch.sum(c6 * qq * t8, dim=-1)
    s6 = param.get("s6", torch.tensor(defaults.S6, **dd))
    s8 = param.get("s8", torch.tensor(defaults.S8, **dd))
    return s6 * e6 + s8 * e8


def dispersion(numbers, positions, param, c6, rvdw=None, r4r2=None, damping_function=rational_damping, cutoff=None, **kwargs):
    if r4r2 is None:
        r4r2 = data.R4R2(device=positions.device, dtype=positions.dtype)[numbers]
    energy = dispersion2(numbers, positions, param, c6, r4r2, damping_function, cutoff, **kwargs)
    s9 = param.get("s9", None)
    if s9 is not None and float(s9) != 0.0:
        raise NotImplementedError("the three-body ATM term is not provided by the oracle shim (GFN1-xTB uses s9 = 0)")
    return energy


def dftd3(numbers, positions, param, *, ref=None, rcov=None, rvdw=None, r4r2=None, cutoff=None, counting_function=exp_count,
          weighting_function=gaussian_weight, damping_function=rational_damping, **kwargs):
    dd = {"device": positions.device, "dtype": positions.dtype}
    ref = (ref or Reference()).to(**dd)
    if rcov is None:
        rcov = COV_D3(**dd)[numbers]
    if r4r2 is None:
        r4r2 = data.R4R2(**dd)[numbers]
    cn = cn_d3(numbers, positions, counting_function=counting_function, rcov=rcov)
    weights = weight_references(numbers, cn, ref, weighting_function)
    c6 = atomic_c6(numbers, weights, ref)
    return dispersion(numbers, positions, param, c6, rvdw, r4r2, damping_function, cutoff)
