"""Exception / warning types with the names dxtb uses on this path
(reference: ``dxtb/_src/exlibs``-free subset of ``dxtb._src.constants`` / ``tad_mctc.exceptions``)."""


class DtypeError(ValueError):
    """Wrong dtype (reference: calculators/types/base.py:518-524, 811-842)."""


class DeviceError(RuntimeError):
    """Wrong device (reference: calculators/types/base.py:811-842)."""


class SCFConvergenceError(RuntimeError):
    """SCF did not converge and ``force_convergence`` is set (scf/unrolling/default.py:123-136)."""


class SCFConvergenceWarning(RuntimeWarning):
    """SCF did not converge (scf/unrolling/default.py:123-136, 324-353)."""


class MissingD3ReferenceError(NotImplementedError):
    """The D3 reference C6 table (third-party data of tad-dftd3) is not available."""
