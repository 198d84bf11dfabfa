"""Version of the package this shim stands in for, and the running torch version as a tuple."""
import torch

__version__ = "0.7.0"


def _tv(v: str):
    out = []
    for p in v.split("+")[0].split(".")[:3]:
        digits = "".join(ch for ch in p if ch.isdigit())
        out.append(int(digits) if digits else 0)
    return tuple(out)


__tversion__ = _tv(torch.__version__)
