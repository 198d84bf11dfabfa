import os

import numpy as np

_CACHE = {}


def table() -> dict:
    path = os.environ.get("TAD_DFTD3_SHIM_TABLE")
    if not path:
        raise RuntimeError(
            "tad-dftd3 stand-in: the D3 reference data (reference CNs, C6 table, r4r2) is third-party data that is not "
            "available offline. Point TAD_DFTD3_SHIM_TABLE to an .npz with keys cn, c6, r4r2, or exclude the dispersion "
            "(opts={'exclude': ['disp']})."
        )
    if path not in _CACHE:
        with np.load(path) as f:
            _CACHE[path] = {k: np.asarray(f[k], dtype=np.float64) for k in ("cn", "c6", "r4r2")}
    return _CACHE[path]
