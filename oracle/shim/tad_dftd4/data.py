def R4R2(device=None, dtype=None):
    raise NotImplementedError("tad_dftd4.data.R4R2 is not provided by the oracle shim")
