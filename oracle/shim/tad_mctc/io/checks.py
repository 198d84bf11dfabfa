"""Input validation used by dxtb's calculators (calculators/types/energy.py:98-106)."""
import torch

from ..data.pse import MAX_ELEMENT
from ..exceptions import MoleculeError


def shape_checks(numbers, positions, allow_batched: bool = True) -> bool:
    if numbers.shape != positions.shape[:-1]:
        raise ValueError(f"Shape of positions ({positions.shape[:-1]}) is not consistent with atomic numbers ({numbers.shape}).")
    if not allow_batched and numbers.ndim > 1:
        raise ValueError("Batched tensors are not allowed.")
    return True


def content_checks(numbers, positions, max_element: int = MAX_ELEMENT, allow_batched: bool = True) -> bool:
    if numbers.max() > max_element:
        raise ValueError(f"Atomic number larger than {max_element} found.")
    if numbers.min() < (0 if allow_batched else 1):
        raise ValueError("Atomic number smaller than 0 found.")
    return True


def deflatable_check(positions, fileinfo=None, **kwargs) -> bool:
    return True
