def cn_d4(*args, **kwargs):
    raise NotImplementedError("cn_d4 is not provided by the oracle shim (D4 / GFN2 are outside the hot path)")
