class D4Model:
    def __init__(self, *args, **kwargs):
        raise NotImplementedError("tad_dftd4.model.D4Model is not provided by the oracle shim (D4 is outside the GFN1 hot path)")
