import glob
import os
import sys

import numpy as np
import pytest

ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
sys.path.insert(0, ROOT)

TOL = 1e-9   # BASELINE.json north_star: fp64 agreement within 1e-9 relative (norm-wise)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def rel_err(a, b):
    """norm-wise relative error  max|a-b| / max(max|b|, 1e-300)  (SURVEY.md 8(d) parity metric)"""
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    assert a.shape == b.shape, (a.shape, b.shape)
    if a.size == 0:
        return 0.0
    return float(np.max(np.abs(a - b)) / max(float(np.max(np.abs(b))), 1e-300))


def golden_cases():
    return sorted(glob.glob(os.path.join(ROOT, "tests", "golden", "*.npz")))


def load_golden(path):
    z = np.load(path, allow_pickle=False)
    return {k: (z[k].item() if z[k].ndim == 0 else z[k]) for k in z.files}


@pytest.fixture(scope="session")
def cuda_lib():
    """The product library on a GPU box.  Fails (does not skip) if it cannot be used there."""
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    from multioutputihgp_b200 import _lib
    return _lib.load()
