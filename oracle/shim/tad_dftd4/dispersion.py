def dispersion2(*args, **kwargs):
    raise NotImplementedError("tad_dftd4.dispersion.dispersion2 is not provided by the oracle shim")


def dispersion3(*args, **kwargs):
    raise NotImplementedError("tad_dftd4.dispersion.dispersion3 is not provided by the oracle shim")
