import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a B200 (run with -m gpu on the GPU box)")


def load_golden(name):
    path = os.path.join(GOLDEN, name)
    if name.endswith(".npy"):
        return np.load(path)
    return np.load(path, allow_pickle=False)


@pytest.fixture(scope="session")
def golden():
    return load_golden


@pytest.fixture(autouse=True, scope="session")
def _fixed_f_table():
    """xraydb is not vendored: product and oracle share one fixed f'/f'' table."""
    from giwaxsim_b200.tools import utilities
    from giwaxsim_b200 import synth
    utilities.set_f1f2_provider(synth.fixed_f1f2)
    yield
