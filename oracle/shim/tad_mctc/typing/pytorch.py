"""Tensor-related typing helpers, most importantly the ``TensorLike`` base class (device / dtype bookkeeping with
``.to`` and ``.type`` that rebuild the object from its ``__slots__``)."""
from __future__ import annotations

from typing import Any, Protocol, TypedDict

import torch
from torch import Tensor

from ..exceptions import DtypeError

__all__ = ["DD", "MockTensor", "Molecule", "Tensor", "TensorLike", "get_default_device", "get_default_dtype"]


class DD(TypedDict):
    device: "torch.device | None"
    dtype: torch.dtype


class Molecule(TypedDict):
    numbers: Tensor
    positions: Tensor


class MockTensor(Tensor):
    @property
    def device(self) -> Any:
        return self._device

    @device.setter
    def device(self, value: Any) -> None:
        self._device = value


def get_default_device() -> torch.device:
    return torch.tensor(1.0).device


def get_default_dtype() -> torch.dtype:
    return torch.get_default_dtype()


class TensorLike:
    __device: "torch.device"
    __dtype: torch.dtype
    __dd: DD
    __slots__ = ["__device", "__dtype", "__dd"]

    def __init__(self, device: "torch.device | None" = None, dtype: "torch.dtype | None" = None):
        self.__device = device if device is not None else get_default_device()
        self.__dtype = dtype if dtype is not None else get_default_dtype()
        self.__dd = {"device": self.device, "dtype": self.dtype}

    @property
    def device(self) -> "torch.device":
        return self.__device

    @device.setter
    def device(self, *_: Any) -> None:
        raise AttributeError("Change object to device using the `.to` method")

    @property
    def dtype(self) -> torch.dtype:
        return self.__dtype

    @dtype.setter
    def dtype(self, *_: Any) -> None:
        raise AttributeError("Change object to dtype using the `.type` method")

    @property
    def dd(self) -> DD:
        return self.__dd

    @property
    def allowed_dtypes(self) -> tuple:
        return (torch.float16, torch.float32, torch.float64)

    def _rebuild(self, cast, **extra):
        if len(self.__slots__) == 0:
            raise RuntimeError(f"The `.type`/`.to` method requires setting `__slots__` in the '{self.__class__.__name__}' class.")
        args = {}
        for s in self.__slots__:
            if s.startswith("__"):
                continue
            attr = getattr(self, s)
            if isinstance(attr, Tensor) or issubclass(type(attr), TensorLike):
                attr = cast(attr)
            args[s] = attr
        return self.__class__(**args, **extra)

    def type(self, dtype: torch.dtype):
        if self.dtype == dtype:
            return self
        if dtype not in self.allowed_dtypes:
            raise DtypeError(f"Only '{self.allowed_dtypes}' allowed (received '{dtype}').")

        def cast(a):
            return a.type(dtype) if a.dtype in self.allowed_dtypes else a

        return self._rebuild(cast, dtype=dtype)

    def to(self, device: "torch.device"):
        if self.device == device:
            return self
        return self._rebuild(lambda a: a.to(device=device), device=device)
