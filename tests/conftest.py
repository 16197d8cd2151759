import glob
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run with -m gpu on a B200)")


def golden_files(pattern="*.npz"):
    return sorted(glob.glob(os.path.join(GOLDEN, pattern)))


def load_golden(path):
    with np.load(path) as z:
        return {k: z[k] for k in z.files}


def pair_goldens():
    """Every fixture recording a (query, train) problem."""
    return [p for p in golden_files() if "collection" not in os.path.basename(p) and "sequence" not in os.path.basename(p)]


@pytest.fixture(scope="session")
def has_cuda():
    import torch
    return torch.cuda.is_available()
