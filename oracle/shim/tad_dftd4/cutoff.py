class Cutoff:
    def __init__(self, *args, **kwargs):
        raise NotImplementedError("tad_dftd4.Cutoff is not provided by the oracle shim")
