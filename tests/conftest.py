import os
import sys

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def pytest_collection_modifyitems(config, items):
    if torch.cuda.is_available():
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


class Golden:
    """Read-only view of one tests/golden/*.npz file as nested case dictionaries."""

    def __init__(self, name):
        self.z = np.load(os.path.join(GOLDEN, name + ".npz"))

    def keys(self, prefix=""):
        return [k for k in self.z.files if k.startswith(prefix)]

    def arr(self, key):
        return self.z[key]

    def t(self, key, dtype=None):
        a = torch.from_numpy(np.array(self.z[key]))
        return a if dtype is None or not a.is_floating_point() else a.to(dtype)

    def group(self, prefix, dtype=None):
        L = len(prefix)
        return {k[L:]: self.t(k, dtype) for k in self.z.files if k.startswith(prefix)}


@pytest.fixture(scope="session")
def golden():
    cache = {}

    def get(name):
        if name not in cache:
            cache[name] = Golden(name)
        return cache[name]

    return get
