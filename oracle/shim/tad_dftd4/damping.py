class Damping:
    """Base class of the damping functions (only instantiated as a default argument at import time)."""


class RationalDamping(Damping):
    def __call__(self, *args, **kwargs):
        raise NotImplementedError("tad_dftd4 damping is not provided by the oracle shim")
