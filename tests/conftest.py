import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")
REFERENCE = "/root/reference"


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def pytest_collection_modifyitems(config, items):
    try:
        import torch

        has_gpu = torch.cuda.is_available()
    except Exception:
        has_gpu = False
    if has_gpu:
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


@pytest.fixture(scope="session")
def golden_blur():
    return np.load(os.path.join(GOLDEN, "blur_torch.npz"))


@pytest.fixture(scope="session")
def golden_taps():
    return np.load(os.path.join(GOLDEN, "blur_torch_taps.npz"))


@pytest.fixture(scope="session")
def golden_swd():
    return np.load(os.path.join(GOLDEN, "swd.npz"))


def uniform_pm1(seed, shape):
    """The input generator of tests/golden/make_golden.py (bit-stable across machines)."""
    return (np.random.RandomState(seed).rand(*shape) * 2.0 - 1.0).astype(np.float32)


has_reference = os.path.isdir(REFERENCE)
needs_reference = pytest.mark.skipif(not has_reference, reason="/root/reference only exists in the build container")
