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


def literal_smoother_err(a, b, big=1e100):
    """Worst per-latent norm-wise relative error of literal-smoother output `a` against the reference `b`, both [..., T, L, d].
    IHGP::backwardSmoother (ihgp.h:108-113) is a backward recursion with gain G; where rho(G) > 1 (default
    Matern-3/2, rho = 6.39) the reference itself overflows a few hundred steps from the end of the sequence.  For such a
    latent the comparison covers the trailing steps where the reference is still below `big`, normalised by that tail's
    own maximum; for a stable latent it covers the whole sequence.  Returns (worst error, fewest steps compared)."""
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    assert a.shape == b.shape and a.ndim >= 3
    T, L = a.shape[-3], a.shape[-2]
    a = a.reshape(-1, T, L, a.shape[-1])
    b = b.reshape(-1, T, L, b.shape[-1])
    worst, fewest = 0.0, T
    for l in range(L):
        with np.errstate(invalid="ignore"):
            m = np.max(np.abs(b[:, :, l, :]), axis=(0, 2))
        bad = ~(np.isfinite(m) & (m < big))
        t0 = int(np.max(np.nonzero(bad)[0])) + 1 if bad.any() else 0
        assert t0 < T, "reference not finite even at the last step"
        worst = max(worst, rel_err(a[:, t0:, l, :], b[:, t0:, l, :]))
        fewest = min(fewest, T - t0)
    return worst, fewest


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
