"""Stand-in for tad-dftd4 0.8.0: dxtb imports it at module level (GFN2 / D4), nothing on the GFN1 path calls it.
Every entry point raises when used."""
from . import cutoff, damping, data, defaults, dispersion, model  # noqa: F401
from .cutoff import Cutoff  # noqa: F401

__version__ = "0.8.0"


class Param(dict):
    """Damping parameters (a plain mapping in the real package's typing)."""


def dftd4(*args, **kwargs):
    raise NotImplementedError("tad_dftd4.dftd4 is not provided by the oracle shim (D4 is outside the GFN1 hot path)")
