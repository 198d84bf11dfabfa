"""Stand-in for tad-dftd3 0.6.0 (D3(BJ), two-body; Grimme et al. JCP 132, 154104 and JCC 32, 1456).

The ARITHMETIC (Gaussian reference weighting, C6 interpolation, rational damping) is restated here; the reference
DATA (reference CNs, 32k reference C6 values, sqrt(Z) r4/r2) ships only with the real package and does not exist offline.
A table of the real shape must be supplied as an .npz (keys cn (Z+1,7), c6 (Z+1,Z+1,7,7), r4r2 (Z+1,)) through the
environment variable TAD_DFTD3_SHIM_TABLE; without it every use of the dispersion raises."""
from . import damping, data, defaults, disp, model, reference, typing  # noqa: F401
from .disp import dftd3  # noqa: F401

__version__ = "0.6.0"
