import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")
    config.addinivalue_line("markers", "slow: long-running")


def load_golden(name):
    return np.load(os.path.join(GOLDEN, name), allow_pickle=False)


@pytest.fixture(scope="session")
def golden_outputs():
    import json
    with open(os.path.join(GOLDEN, "workflow_outputs.json")) as fh:
        return json.load(fh)


@pytest.fixture(scope="session")
def small_panel():
    d = load_golden("small_panel.npz")
    return {k: d[k] for k in d.files}


@pytest.fixture(scope="session")
def sample_inbred():
    d = load_golden("sample_inbred.npz")
    return {k: d[k] for k in d.files}


@pytest.fixture(scope="session")
def sample_cross():
    d = load_golden("sample_cross.npz")
    return {k: d[k] for k in d.files}
