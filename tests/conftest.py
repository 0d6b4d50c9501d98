import os

os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")   # several ranks share one process in the in-process peer tests
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


class Golden:
    """Access to tests/golden/<group>.npz written by tests/golden/make_golden.py (reference outputs)."""

    def __init__(self, group):
        self._d = np.load(os.path.join(GOLDEN, group + ".npz"))

    def cases(self):
        return sorted({k.split("__")[0] for k in self._d.files})

    def case(self, name):
        pre = name + "__"
        return {k[len(pre):]: self._d[k] for k in self._d.files if k.startswith(pre)}


@pytest.fixture(scope="session")
def golden():
    cache = {}

    def get(group):
        if group not in cache:
            cache[group] = Golden(group)
        return cache[group]
    return get


def rel_l2(a, b):
    a = np.asarray(a, dtype=np.float64).ravel(); b = np.asarray(b, dtype=np.float64).ravel()
    nb = np.linalg.norm(b)
    return np.linalg.norm(a - b) / (nb if nb > 0 else 1.0)


def rel_max(a, b):
    a = np.asarray(a, dtype=np.float64).ravel(); b = np.asarray(b, dtype=np.float64).ravel()
    s = np.abs(b).max() if b.size else 1.0
    return np.abs(a - b).max() / (s if s > 0 else 1.0) if b.size else 0.0
