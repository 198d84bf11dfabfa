"""Stand-in for tad-multicharge 0.5.0: ``get_eeq_charges`` (EEQ-2019 model, Caldeweyher et al., JCP 150, 154122).

Restated from the publication: charges minimise  sum_i [chi_i - kcn_i sqrt(CN_i)] q_i + 1/2 sum_i (eta_i + sqrt(2/pi)/rad_i) q_i^2
+ 1/2 sum_{i!=j} q_i q_j erf(gamma_ij R_ij)/R_ij  subject to  sum q = total charge  (gamma_ij = 1/sqrt(rad_i^2+rad_j^2)),
with the erf-counting CN capped at 8.  Parameters: dxtb_b200/data/gfn1_param.json "eeq2019" (restated, pinned for H/C
only).  Written with differentiable torch ops so that dxtb's autograd forces see the guess's position dependence."""
import math

import torch

from tad_mctc import storch
from tad_mctc.batch import real_atoms, real_pairs
from tad_mctc.data._blob import third_party
from tad_mctc.ncoord.eeq import cn_eeq

__version__ = "0.5.0"
__all__ = ["get_eeq_charges", "get_charges"]


def _param(name, device, dtype):
    return torch.tensor([0.0] + third_party()["eeq2019"][name], device=device, dtype=dtype)


def get_eeq_charges(numbers, positions, chrg, *, cutoff=None, **kwargs):
    dd = {"device": positions.device, "dtype": positions.dtype}
    eps = torch.tensor(torch.finfo(positions.dtype).eps, **dd)
    real = real_atoms(numbers)
    mask = real_pairs(numbers, mask_diagonal=True)
    chi, eta, kcn, rad = (_param(k, **dd)[numbers] for k in ("chi", "eta", "kcn", "rad"))
    cn = cn_eeq(numbers, positions, cutoff=cutoff)
    chrg = torch.as_tensor(chrg, **dd)
    rhs = torch.where(real, -chi + storch.sqrt(cn) * kcn, torch.tensor(0.0, **dd))
    rhs = torch.cat([rhs, chrg.reshape(*rhs.shape[:-1], 1)], dim=-1)

    dist = torch.where(mask, storch.cdist(positions, positions, p=2), eps)
    rad_safe = torch.where(real, rad, torch.tensor(1.0, **dd))
    gam = 1.0 / torch.sqrt(rad_safe.unsqueeze(-1) ** 2 + rad_safe.unsqueeze(-2) ** 2)
    coul = torch.where(mask, torch.erf(dist * gam) / dist, torch.tensor(0.0, **dd))
    diag = torch.where(real, eta + math.sqrt(2.0 / math.pi) / rad_safe, torch.tensor(1.0, **dd))  # padding rows: q = 0
    coul = coul + torch.diag_embed(diag)
    n = numbers.shape[-1]
    constraint = real.to(positions.dtype)
    top = torch.cat([coul, constraint.unsqueeze(-1)], dim=-1)
    bottom = torch.cat([constraint, torch.zeros(*constraint.shape[:-1], 1, **dd)], dim=-1).unsqueeze(-2)
    amat = torch.cat([top, bottom], dim=-2)
    x = torch.linalg.solve(amat, rhs.unsqueeze(-1)).squeeze(-1)
    return x[..., :n]


get_charges = get_eeq_charges
