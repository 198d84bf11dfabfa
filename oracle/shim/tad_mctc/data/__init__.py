from . import getters, mass, pse, radii  # noqa: F401
from .radii import ATOMIC_RADII, COV_D3, VDW_D3, VDW_PAIRWISE  # noqa: F401
from .mass import ATOMIC_MASS  # noqa: F401


def PAULING(device=None, dtype=None):
    raise NotImplementedError("PAULING electronegativities are not carried by the oracle shim (D4 is outside the hot path)")
