import sys
from pathlib import Path

import numpy as np
import pytest

REPO = Path(__file__).resolve().parents[1]
if str(REPO) not in sys.path:
    sys.path.insert(0, str(REPO))

GOLDEN = REPO / "tests" / "golden"


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def golden_extraction():
    return np.load(GOLDEN / "extraction.npz")


@pytest.fixture(scope="session")
def golden_constants():
    return np.load(GOLDEN / "constants.npz")


@pytest.fixture(scope="session")
def golden_cloak():
    return np.load(GOLDEN / "cloak.npz")


@pytest.fixture(scope="session")
def golden_norm():
    return np.load(GOLDEN / "norm.npz")

collect_ignore = ["ref_fixture"]          # verbatim reference sources (test fixtures), nothing to collect
