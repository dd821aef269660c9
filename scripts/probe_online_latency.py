"""Latency breakdown of one objective evaluation at the streaming learner's shape (config 2: p = 8, L = 4, Matern-3/2):
update(params), objective over a window of W observations (host buffers), per-observation step.  Diagnosis only."""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np

from multioutputihgp_b200 import MOIHGPSequences
from multioutputihgp_b200.pywrapper import MOIHGP

p, L = 8, 4
rng = np.random.default_rng(3)
seq = MOIHGPSequences(0.1, p, L, "Matern32")
params = seq.params.copy()
params[: p * L] = rng.standard_normal(p * L)


def timeit(f, n=300):
    for _ in range(20):
        f()
    t = time.perf_counter()
    for _ in range(n):
        f()
    return 1e6 * (time.perf_counter() - t) / n


print("update(params)            : %7.1f us" % timeit(lambda: seq.update(params)))
for W in (1, 64, 256, 1024):
    Y = rng.standard_normal((1, W, p))
    x0 = np.zeros((1, L, 2)); dx0 = np.zeros((1, L, 3, 2))
    print("objective  W=%-5d host    : %7.1f us" % (W, timeit(lambda: seq.objective(Y, x0, dx0, want_state=True))))
    seq.bind(Y)
    print("objective  W=%-5d bound   : %7.1f us" % (W, timeit(lambda: seq.objective_bound(x0, dx0))))
    if hasattr(seq, "online_eval"):
        print("online_eval W=%-5d        : %7.1f us" % (W, timeit(lambda: seq.online_eval(params, x0, dx0))))
gp = MOIHGP(0.1, p, L, "Matern32", False)
x = np.zeros((L, 2)); dx = np.zeros((L, 3, 2)); y = rng.standard_normal(p)
print("step (x,y,dx) legacy      : %7.1f us" % timeit(lambda: gp.step(x, y, dx)))
print("negLogLikelihood legacy   : %7.1f us" % timeit(lambda: gp.negLogLikelihood(x, y, dx)))
