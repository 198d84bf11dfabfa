import torch


def center_of_mass(masses, positions):
    return (masses.unsqueeze(-1) * positions).sum(-2) / masses.sum(-1, keepdim=True)


def positions_rel_com(masses, positions):
    return positions - center_of_mass(masses, positions).unsqueeze(-2)


def inertia_moment(masses, positions, center_pa: bool = True, pos_already_com: bool = False):
    r = positions if pos_already_com else positions_rel_com(masses, positions)
    r2 = (r * r).sum(-1)
    eye = torch.eye(3, dtype=positions.dtype, device=positions.device)
    return torch.einsum("...a,...aij->...ij", masses, r2[..., None, None] * eye - r.unsqueeze(-1) * r.unsqueeze(-2))
