from __future__ import annotations

import torch

from ._table import table


class Reference:
    """Reference systems: coordination numbers ``cn`` (Z+1, 7; -1 = absent) and C6 coefficients ``c6`` (Z+1, Z+1, 7, 7)."""

    __slots__ = ["cn", "c6", "__device", "__dtype"]

    def __init__(self, cn=None, c6=None, device=None, dtype=None):
        if cn is None or c6 is None:
            # kept in double unless a dtype is asked for: dxtb calls Reference().to(float64), and a detour through the
            # default float32 would round the supplied table
            t = table()
            cn = torch.tensor(t["cn"], device=device, dtype=dtype if dtype is not None else torch.float64)
            c6 = torch.tensor(t["c6"], device=device, dtype=dtype if dtype is not None else torch.float64)
        self.cn, self.c6 = cn, c6
        self.__device, self.__dtype = cn.device, cn.dtype

    device = property(lambda self: self.__device)
    dtype = property(lambda self: self.__dtype)

    def to(self, device=None, dtype=None) -> "Reference":
        return Reference(self.cn.to(device=device, dtype=dtype), self.c6.to(device=device, dtype=dtype))

    def type(self, dtype) -> "Reference":
        return Reference(self.cn.type(dtype), self.c6.type(dtype))
