import json
import sys
from pathlib import Path

import numpy as np
import pytest

ROOT = Path(__file__).resolve().parent.parent
if str(ROOT) not in sys.path:
    sys.path.insert(0, str(ROOT))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (B200)")


@pytest.fixture(scope="session")
def mols():
    return json.load(open(ROOT / "tests" / "golden" / "molecules.json"))


@pytest.fixture(scope="session")
def goldens():
    return np.load(ROOT / "tests" / "golden" / "reference.npz")


@pytest.fixture(scope="session")
def energies():
    return json.load(open(ROOT / "tests" / "golden" / "energies.json"))
