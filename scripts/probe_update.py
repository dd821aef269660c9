"""update(params) at BASELINE config-5 shape (p = 256, L = 64): wall time per call; run under ncu for the kernel split."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from bench import model_params, DT
from multioutputihgp_b200 import MOIHGPSequences
p, L = int(os.environ.get("FIT_P", 256)), int(os.environ.get("FIT_L", 64))
params, _ = model_params(p, L, "Matern32", 1238)
rng = np.random.default_rng(1)
m = MOIHGPSequences(DT, p, L, "Matern32", threading=True)
raw = params.copy(); raw[:p * L] += 0.05 * rng.standard_normal(p * L)     # a raw (non-orthonormal) U block, as inside a line search
for _ in range(3): m.update(raw)
t = time.perf_counter()
for _ in range(20): m.update(raw)
m.synchronize()
print("update p=%d L=%d: %.3f ms per call" % (p, L, 1e3 * (time.perf_counter() - t) / 20))
