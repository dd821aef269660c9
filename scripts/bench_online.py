"""BASELINE config 2 (example_online_learning streaming update: p = 8, L = 4, Matern-3/2, window 1): per-sample latency of
MOIHGPOnlineLearning.step on the GPU library.  Latency-bound by construction (SURVEY 7.3 hard part 6): a handful of
L-BFGS-B objective evaluations per sample, each a few microseconds of arithmetic behind ~10 kernel launches.
Prints one line; not the headline benchmark."""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np

from multioutputihgp_b200 import MOIHGPOnlineLearning

steps = int(os.environ.get("ONLINE_STEPS", 300))
p, L, window = 8, 4, int(os.environ.get("ONLINE_WINDOW", 1))
rng = np.random.default_rng(7)
gp = MOIHGPOnlineLearning(0.1, p, L, gamma=0.9, windowsize=window, threading=False)
t = np.arange(steps + 20) * 0.1
data = np.stack([np.sin((1 + i % 3) * t + 0.3 * i) for i in range(p)], axis=1) + 0.05 * rng.standard_normal((steps + 20, p))
for y in data[:20]:
    gp.step(y.copy())
l0 = gp._seq.launch_count + gp.moihgp.launch_count if hasattr(gp.moihgp, "launch_count") else gp._seq.launch_count
tic = time.perf_counter()
for y in data[20:]:
    gp.step(y.copy())
dt = time.perf_counter() - tic
l1 = gp._seq.launch_count + gp.moihgp.launch_count if hasattr(gp.moihgp, "launch_count") else gp._seq.launch_count
print("online learner (config 2 shape p=%d L=%d window=%d): %.2f ms per streamed sample, %.0f latent-steps/s, %.1f kernel launches per sample"
      % (p, L, window, 1e3 * dt / steps, steps * L / dt, (l1 - l0) / steps))
