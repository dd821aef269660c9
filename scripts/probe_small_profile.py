"""Per-kernel device times of one objective evaluation at the streaming learner's shape (p = 8, L = 4, W = 64)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from multioutputihgp_b200 import MOIHGPSequences
p, L, W = 8, 4, 64
rng = np.random.default_rng(3)
seq = MOIHGPSequences(0.1, p, L, "Matern32")
params = seq.params.copy(); params[:p * L] = rng.standard_normal(p * L)
seq.update(params)
Y = rng.standard_normal((1, W, p)); x0 = np.zeros((1, L, 2)); dx0 = np.zeros((1, L, 3, 2))
for _ in range(5): seq.objective(Y, x0, dx0)
seq.profile(True)
for _ in range(3):
    seq.objective(Y, x0, dx0)
    print(seq.profile_read())
seq.profile(False)
import torch
ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
torch.cuda.synchronize()
t = time.perf_counter()
for _ in range(200): seq.update(params)
print("update wall %.1f us" % (1e6 * (time.perf_counter() - t) / 200))
