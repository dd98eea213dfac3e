import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = os.path.join(ROOT, "tests", "golden")
for path in (ROOT, GOLDEN):
    if path not in sys.path:
        sys.path.insert(0, path)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def load_ragged(npz, prefix):
    offsets, values = npz[prefix + "/offsets"], npz[prefix + "/values"]
    return [values[offsets[i] : offsets[i + 1]] for i in range(len(offsets) - 1)]


@pytest.fixture(scope="session")
def golden():
    """name -> lazily loaded npz of tests/golden"""
    cache = {}

    def get(name):
        if name not in cache:
            cache[name] = np.load(os.path.join(GOLDEN, name + ".npz"))
        return cache[name]

    return get


@pytest.fixture(scope="session")
def speech():
    import pydrobert_speech_b200

    return pydrobert_speech_b200
