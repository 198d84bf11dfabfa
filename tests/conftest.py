import json
import sys
from pathlib import Path

import numpy as np
import pytest

ROOT = Path(__file__).resolve().parent.parent
if str(ROOT) not in sys.path:
    sys.path.insert(0, str(ROOT))
if str(ROOT / "tests") not in sys.path:
    sys.path.insert(0, str(ROOT / "tests"))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (B200)")


def pytest_collection_modifyitems(config, items):
    """gpu-marked tests are skipped (not failed) on a box without a CUDA device."""
    import torch

    if torch.cuda.is_available():
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


@pytest.fixture(scope="session")
def mols():
    return json.load(open(ROOT / "tests" / "golden" / "molecules.json"))


@pytest.fixture(scope="session")
def goldens():
    return np.load(ROOT / "tests" / "golden" / "reference.npz")


@pytest.fixture(scope="session")
def energies():
    return json.load(open(ROOT / "tests" / "golden" / "energies.json"))
