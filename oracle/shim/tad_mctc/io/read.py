"""File readers are not part of the hot path: dxtb's CLI imports this module, nothing on the single-point path calls it."""


def __getattr__(name):
    def _missing(*a, **k):
        raise NotImplementedError(f"tad_mctc.io.read.{name} is not provided by the oracle shim")

    return _missing
