import torch

from ._table import table


def R4R2(device=None, dtype=None):
    """sqrt(0.5 * sqrt(Z) * <r4>/<r2>) per element (index 0 = padding)."""
    return torch.tensor(table()["r4r2"], device=device, dtype=dtype if dtype is not None else torch.get_default_dtype())
