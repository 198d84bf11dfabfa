def ATOMIC_MASS(device=None, dtype=None):
    raise NotImplementedError("ATOMIC_MASS is not carried by the oracle shim (vibrational analysis is outside the hot path)")
