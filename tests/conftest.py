import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def fixtures():
    """The reference's two test images (tests/golden/make_golden.py), decoded RGBA8."""
    import numpy as np
    z = np.load(os.path.join(ROOT, "tests", "golden", "fixtures.npz"))
    return {k: z[k] for k in z.files}


@pytest.fixture(scope="session")
def golden_hashes():
    import json
    with open(os.path.join(ROOT, "tests", "golden", "reference_hashes.json")) as f:
        return json.load(f)
