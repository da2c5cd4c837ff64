import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run by the driver with -m gpu)")


def load_golden(name):
    """npz -> nested dict ('case.key' entries become d[case][key])."""
    z = np.load(os.path.join(GOLDEN, name))
    out = {}
    for k in z.files:
        v = z[k]
        if v.shape == ():
            v = v.item()
        if "." in k:
            c, kk = k.split(".", 1)
            out.setdefault(c, {})[kk] = v
        else:
            out[k] = v
    return out


@pytest.fixture(scope="session")
def golden():
    return load_golden
