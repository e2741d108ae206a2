import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def golden():
    return np.load(os.path.join(GOLDEN, "golden_ref.npz"))


@pytest.fixture(scope="session")
def q100():
    from oracle.pyoracle import load_quality_csv
    return load_quality_csv(os.path.join(GOLDEN, "quality.csv"))


@pytest.fixture(scope="session", autouse=True)
def _built():
    """Build the checker (oracle port) and the product library once per session."""
    import __graft_entry__ as g
    g.build()


def golden_state(golden, prefix):
    from oracle.pyoracle import State
    return State(*(np.array(golden[f"{prefix}_{k}"]) for k in
                   ("coords", "dir", "speed", "speed_factor", "lag", "sample_rate", "food_intake")))
