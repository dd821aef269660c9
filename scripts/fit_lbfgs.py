"""BASELINE config 5 as it is used: an L-BFGS-B hyper-parameter fit on ONE long sequence resident on the GPU
(p = 256, L = 64, Matern-3/2; T from FIT_T, default 1e6).  Each objective evaluation = update(params) (device polar
factor + K-setup) + one NLL/gradient pass over the bound data; SciPy's L-BFGS-B drives it on the host, as LBFGS++ does in
the reference (moihgp_regression.h:118-124).  Prints evaluations/s; not the headline benchmark."""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from scipy.optimize import minimize

from bench import model_params, DT
from multioutputihgp_b200 import MOIHGPSequences

p, L, T = int(os.environ.get("FIT_P", 256)), int(os.environ.get("FIT_L", 64)), int(float(os.environ.get("FIT_T", 1e6)))
params0, Hmix = model_params(p, L, "Matern32", 1238)
rng = np.random.default_rng(5)
t = np.arange(T) * DT
Y = np.sin(t[:, None] * (1.0 + 3.0 * np.arange(L) / max(L - 1, 1))[None, :]) @ Hmix.T + 0.1 * (2 * rng.random((T, p)) - 1)
m = MOIHGPSequences(DT, p, L, "Matern32", threading=True)
m.bind(Y[None])
evals, t_upd, t_obj = [0], [0.0], [0.0]


def fun(x):
    a = time.perf_counter()
    m.update(x)
    b = time.perf_counter()
    loss, grad = m.objective_bound()
    c = time.perf_counter()
    evals[0] += 1
    t_upd[0] += b - a
    t_obj[0] += c - b
    return loss, grad


bounds = [(-1e4, 1e4)] * (p * L) + [(1e-4, 1e4)] * L + [(1e-4, 1e2)] * (1 + 3 * L)      # moihgp_regression.h:91-98
fun(params0)
evals[0], t_upd[0], t_obj[0] = 0, 0.0, 0.0
tic = time.perf_counter()
res = minimize(fun, params0, jac=True, method="L-BFGS-B", bounds=bounds, options={"maxiter": int(os.environ.get("FIT_ITERS", 5)), "maxls": 20})
wall = time.perf_counter() - tic
print("L-BFGS-B fit, p=%d L=%d T=%d (%d parameters): %d iterations, %d objective evaluations in %.2f s -> %.1f ms per evaluation "
      "(update %.1f ms, objective %.1f ms) = %.3g latent-steps/s; loss %.6g -> %.6g"
      % (p, L, T, params0.size, res.nit, evals[0], wall, 1e3 * wall / max(evals[0], 1), 1e3 * t_upd[0] / max(evals[0], 1),
         1e3 * t_obj[0] / max(evals[0], 1), evals[0] * T * L / wall, fun(params0)[0], res.fun))
