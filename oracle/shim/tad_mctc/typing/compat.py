from __future__ import annotations

import os
import sys
from typing import Callable, Generator, Sequence, Union

import torch
from torch import Tensor
from typing_extensions import Self, TypeAlias, TypeGuard, override

__all__ = ["Callable", "CountingFunction", "DampingFunction", "Generator", "PathLike", "Self", "Sequence", "Size", "Sliceable",
           "TensorOrTensors", "TypeAlias", "TypeGuard", "override"]

CountingFunction = Callable[[Tensor, Tensor], Tensor]
DampingFunction = Callable[[int, Tensor, Tensor, dict], Tensor]
PathLike = Union[str, os.PathLike]
Size = Union[list, tuple, torch.Size]
Sliceable = Union[list, tuple]
TensorOrTensors = Union[list, tuple, Tensor]
