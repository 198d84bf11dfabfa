"""Minimal stand-in for tad-mctc 0.7.0 (test infrastructure; see oracle/shim/README.md)."""
from ._version import __version__
from . import exceptions, typing, math, storch, batch, convert, data, ncoord, units, autograd, molecule, io
from .io import read

__all__ = ["autograd", "batch", "convert", "data", "exceptions", "io", "math", "molecule", "ncoord", "read", "storch",
           "typing", "units", "__version__"]
