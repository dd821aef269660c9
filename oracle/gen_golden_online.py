"""TEST INFRASTRUCTURE: golden trajectory of the reference's OWN Python online learner.

Runs the UNMODIFIED reference files moihgp/{__init__,pywrapper,online_learning}.py (copied to a scratch directory outside
the repository, because pywrapper.py:22 loads lib/libmoihgp.so relative to its own location and /root/reference is
read-only) against the reference's C ABI compiled from its own sources (oracle/_ref/libmoihgp_ref_O2.so: src/wrapper.cpp +
headers, Eigen-API shim), and stores inputs + outputs as tests/golden_online/online_py_*.npz.  Only runs where
/root/reference exists (this container); the fixture travels to the GPU box.

    python oracle/gen_golden_online.py
"""
import importlib
import os
import shutil
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REF = "/root/reference/moihgp"


def load_reference_package():
    tmp = tempfile.mkdtemp(prefix="refpy_")
    pkg = os.path.join(tmp, "moihgp")
    os.makedirs(os.path.join(pkg, "lib"))
    for f in ("__init__.py", "pywrapper.py", "online_learning.py"):
        shutil.copy(os.path.join(REF, f), os.path.join(pkg, f))
    shutil.copy(os.path.join(HERE, "_ref", "libmoihgp_ref_O2.so"), os.path.join(pkg, "lib", "libmoihgp.so"))
    sys.path.insert(0, tmp)
    return importlib.import_module("moihgp"), tmp


def case(mod, p, L, window, steps, seed, nan_rate=0.0):
    rng = np.random.default_rng(seed)
    dt = 0.1
    gp = mod.MOIHGPOnlineLearning(dt, p, L, gamma=0.9, windowsize=window, threading=False)
    # the constructor's U is random (moihgp.h:105, Q11): inject deterministic hyper-parameters
    U0 = np.eye(p, L) + 0.2 * rng.standard_normal((p, L))
    tbl = [(1, 1, .1), (.5, .5, .1), (2, .3, .05), (.5, .3, .5)]
    params0 = np.concatenate([U0.ravel(), 0.5 + 0.25 * np.arange(L), [0.05], np.array([tbl[l % 4] for l in range(L)], dtype=float).ravel()])
    gp.moihgp.update(params0)
    t = np.arange(steps) * dt
    data = np.stack([np.sin((1 + i % 3) * t + 0.3 * i) for i in range(p)], axis=1) + 0.05 * rng.standard_normal((steps, p))
    if nan_rate > 0:
        data[rng.random(data.shape) < nan_rate] = np.nan
    yh, ps = [], []
    for y in data:
        yh.append(gp.step(y.copy()).copy())
        ps.append(gp.params.copy())
    return dict(dt=dt, p=p, L=L, window=window, gamma=0.9, params0=params0, data=data, yhat=np.array(yh), params=np.array(ps))


if __name__ == "__main__":
    mod, tmp = load_reference_package()
    out = os.path.join(HERE, "..", "tests", "golden_online")
    os.makedirs(out, exist_ok=True)
    np.savez(os.path.join(out, "online_py_p8L4_w2.npz"), **case(mod, 8, 4, 2, 40, 11))          # example.py's shape (8 outputs, 4 latents, window 2)
    np.savez(os.path.join(out, "online_py_p4L2_w1.npz"), **case(mod, 4, 2, 1, 30, 12))          # example_online_learning.cpp's window 1
    shutil.rmtree(tmp, ignore_errors=True)
    print("written:", sorted(os.listdir(out)))
