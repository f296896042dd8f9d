import json
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
def golden_dir():
    return GOLDEN


@pytest.fixture(scope="session")
def reference_vectors():
    """Vectors produced by the unmodified reference over shims (oracle/make_golden.py)."""
    with open(os.path.join(GOLDEN, "reference_vectors.json")) as fh:
        data = json.load(fh)
    for ds in data.values():
        for mode in ds["modes"].values():
            mode["values"] = np.array([float.fromhex(h) for h in mode["hex"]])
    return data


@pytest.fixture(scope="session")
def ref_golden_distmat():
    """The reference's own golden vector test/ref/test_distmat.npy (-n -ns 17)."""
    return np.load(os.path.join(GOLDEN, "ref_test_distmat.npy"))
