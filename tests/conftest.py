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


def _load(name):
    with np.load(os.path.join(GOLDEN, name)) as z:
        flat = {k: z[k] for k in z.files}
    out = {}
    for k, v in flat.items():
        case, field = k.split("/", 1)
        out.setdefault(case, {})[field] = v
    return out


@pytest.fixture(scope="session")
def cg_golden():
    return _load("cg_golden.npz")


@pytest.fixture(scope="session")
def models_golden():
    return _load("models_golden.npz")
