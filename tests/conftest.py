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


class TapeRng:
    """Replays recorded *standard* normals through the RandomState.normal signature."""

    def __init__(self, z):
        self.z = np.asarray(z, dtype=np.float64).ravel()
        self.pos = 0

    def normal(self, loc=0.0, scale=1.0, size=None):
        n = int(np.prod(size)) if size is not None else 1
        out = self.z[self.pos:self.pos + n]
        assert out.size == n, "normal tape exhausted"
        self.pos += n
        return loc + scale * out.reshape(size if size is not None else ())


class replay_uniforms:
    """Context manager: np.random.rand() hands out the recorded global uniforms."""

    def __init__(self, u):
        self.u = list(np.atleast_1d(np.asarray(u, dtype=np.float64)))
        self.pos = 0

    def __enter__(self):
        self._orig = np.random.rand

        def rand(*a):
            assert not a
            v = self.u[self.pos] if self.pos < len(self.u) else 0.5
            self.pos += 1
            return v

        np.random.rand = rand
        return self

    def __exit__(self, *exc):
        np.random.rand = self._orig
        return False


def have_gpu():
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False
