"""Element tables restated from the published sources; one copy lives in dxtb_b200/data/gfn1_param.json
("third_party" section, written by tools/make_param_blob.py) and is read here so shim and product cannot diverge."""
import json
from pathlib import Path

_PATH = Path(__file__).resolve().parents[4] / "dxtb_b200" / "data" / "gfn1_param.json"
_BLOB = None


def third_party() -> dict:
    global _BLOB
    if _BLOB is None:
        _BLOB = json.loads(_PATH.read_text())["third_party"]
    return _BLOB
