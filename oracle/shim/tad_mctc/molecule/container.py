from __future__ import annotations

import torch

from ..typing import TensorLike


class Mol(TensorLike):
    """Numbers / positions / charge container (constructed by dxtb only in optional convenience paths)."""

    __slots__ = ["_numbers", "_positions", "_charge", "_name"]

    def __init__(self, numbers, positions, charge=0, name=None, device=None, dtype=None):
        super().__init__(device if device is not None else positions.device, dtype if dtype is not None else positions.dtype)
        self._numbers, self._positions, self._name = numbers, positions, name
        self._charge = torch.as_tensor(charge, device=self.device, dtype=self.dtype)

    numbers = property(lambda self: self._numbers)
    positions = property(lambda self: self._positions)
    charge = property(lambda self: self._charge)
    name = property(lambda self: self._name)
