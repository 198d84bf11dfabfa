"""Radii tables as callables ``TABLE(device=None, dtype=None) -> Tensor`` indexed by atomic number (0 = padding), in Bohr."""
import torch

from ._blob import third_party


def _aa2au():
    c = third_party()["codata2018"]
    return 1.0 / (c["bohr_m"] * 1e10)


def _table(values_aa, scale=1.0, zmax=118):
    vals = [0.0] + [v * _aa2au() * scale for v in values_aa]
    vals += [0.0] * (zmax + 1 - len(vals))

    def get(device=None, dtype=None):
        return torch.tensor(vals, device=device, dtype=dtype if dtype is not None else torch.get_default_dtype())

    return get


def ATOMIC_RADII(device=None, dtype=None):
    return _table(third_party()["atomic_radii_angstrom"])(device, dtype)


def COV_D3(device=None, dtype=None):
    """D3 covalent radii: Pyykko/Atsumi 2009 (metals scaled by 0.9) times 4/3."""
    return _table(third_party()["cov_2009_angstrom"], 4.0 / 3.0)(device, dtype)


def _absent(name):
    def get(device=None, dtype=None):
        raise NotImplementedError(f"tad_mctc.data.radii.{name} is third-party data that the oracle shim does not carry "
                                  "(it is not on the GFN1 single-point path)")

    return get


VDW_D3 = _absent("VDW_D3")


def VDW_PAIRWISE(device=None, dtype=None):
    """Pairwise vdW radii enter only the zero-damping / ATM variants of D3; GFN1 uses rational damping with s9 = 0.
    dxtb still indexes the table when it builds the D3 cache, so hand out zeros of the right shape."""
    return torch.zeros((104, 104), device=device, dtype=dtype if dtype is not None else torch.get_default_dtype())
