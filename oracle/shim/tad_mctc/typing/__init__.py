from typing import Any, Callable, Protocol

from torch import Tensor

from .builtin import *  # noqa: F401,F403
from .compat import *  # noqa: F401,F403
from .pytorch import DD, MockTensor, Molecule, TensorLike, get_default_device, get_default_dtype  # noqa: F401

CNFunc = Callable[..., Tensor]
CNFunction = Callable[..., Tensor]
CNGradFunction = Callable[..., Tensor]
