"""Exception types dxtb re-exports (dxtb/_src/typing/exceptions)."""


class DeviceError(RuntimeError):
    """Tensors live on different devices."""


class DtypeError(ValueError):
    """Tensor has the wrong dtype."""


class FormatError(ValueError):
    pass


class FormatErrorORCA(FormatError):
    pass


class FormatErrorTM(FormatError):
    pass


class EmptyFileError(RuntimeError):
    pass


class MoleculeError(RuntimeError):
    pass


class MoleculeWarning(UserWarning):
    pass
