import os
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def has_gpu():
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False


def pytest_collection_modifyitems(config, items):
    if has_gpu():
        return
    skip = pytest.mark.skip(reason="no CUDA device in this environment")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


@pytest.fixture(scope="session")
def apd_lib_path():
    """Builds libapd_b200.so in-tree if it is stale (nvcc cross-compiles without a GPU)."""
    from audio_pattern_discovery_b200 import build as b
    return b.build()


def random_sequences(rng, n, lo, hi, dim, integer=False):
    """integer=True gives tie-heavy small-integer frames (exercises the strict-< quirk)."""
    out = []
    for _ in range(n):
        t = int(rng.integers(lo, hi + 1))
        if integer:
            out.append(rng.integers(0, 3, size=(t, dim)).astype(np.float32))
        else:
            out.append(rng.normal(size=(t, dim)).astype(np.float32))
    return out
