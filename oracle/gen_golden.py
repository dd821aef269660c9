"""TEST INFRASTRUCTURE: generate tests/golden/*.npz from the REFERENCE ITSELF.

Runs only in the build container (needs /root/reference and oracle/_ref, see oracle/Makefile):
the unmodified reference sources, compiled against the Eigen-API shim, are driven through their
own entry points (gp32_* of src/wrapper.cpp; oracle/ref_probe.cpp for what wrapper.cpp cannot
reach) on seeded inputs, and inputs + outputs are frozen as small fixtures.  The fixtures travel
to the GPU box; /root/reference does not.

    python oracle/gen_golden.py
"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from oracle.binding import RefMOIHGP, build, ref_available  # noqa: E402

OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests", "golden")

# per-latent (magnitude, lengthscale, noise) tables, filter-stable under the reference's semantics
TABLE = {
    "Matern32": [(1, 1, .1), (.5, .5, .1), (2, .3, .05), (.5, .3, .5)],
    "Matern52": [(1, 1, .1), (.5, .5, .1), (.5, .3, .05), (.5, .5, .5)],
}
# the same, restricted to hyper-parameters where the LITERAL backward smoother (ihgp.h:105-113) is stable too:
# rho(G) = 0.32, 0.03, 0.21, 0.40 (Matern-3/2) and 0.35, 0.22, 0.60, 0.22 (Matern-5/2) at dt = 0.1, so that smoothed means can
# be compared over whole sequences of any length (the default Matern-3/2 entry (1, 1, .1) has rho(G) = 6.39 and overflows)
LITERAL_STABLE = {
    "Matern32": [(.5, .5, .1), (2, .3, .05), (.5, .3, .5), (10, 1, .1)],
    "Matern52": TABLE["Matern52"],
}


def make_params(rng, p, L, kernel, table=None):
    U = (np.eye(p, L) + 0.3 * rng.standard_normal((p, L))).ravel()
    S = rng.uniform(0.5, 2.0, L)
    sigma = 0.05
    tbl = (table or TABLE)[kernel]
    igp = np.array([tbl[l % len(tbl)] for l in range(L)], dtype=np.float64).ravel()
    return np.concatenate([U, S, [sigma], igp])


def make_data(rng, p, L, T, dt=0.1):
    """example_regression.cpp:18-28 style: sinusoidal latents mixed by a random matrix + uniform noise."""
    t = np.arange(T) * dt
    w = 1.0 + 3.0 * np.arange(L) / max(L - 1, 1)
    F = np.sin(np.outer(t, w))
    H = rng.standard_normal((p, L)) / np.sqrt(L)
    return F @ H.T + 0.1 * rng.uniform(-1, 1, (T, p))


def case(kernel, threading, p, L, T, seed, dt=0.1):
    rng = np.random.default_rng(seed)
    params = make_params(rng, p, L, kernel)
    Y = make_data(rng, p, L, T, dt)
    ref = RefMOIHGP(dt, p, L, kernel, threading)
    ref.update(params)
    out = {"kernel": kernel, "threading": int(threading), "p": p, "L": L, "T": T, "dt": dt, "params": params, "Y": Y}
    out["params_after_update"] = ref.params.copy()
    loss, grad, xT, dxT = ref.objective(Y)
    out.update(obj_loss=loss, obj_grad=grad, obj_xT=xT, obj_dxT=dxT)
    # carried (non-zero) initial state, as the online objective uses (moihgp_online.h:57-58)
    x0 = 0.3 * rng.standard_normal((L, ref.d))
    dx0 = 0.1 * rng.standard_normal((L, 3, ref.d))
    loss2, grad2, xT2, dxT2 = ref.objective(Y[: max(T // 3, 2)], x0, dx0)
    out.update(x0=x0, dx0=dx0, obj2_loss=loss2, obj2_grad=grad2, obj2_xT=xT2, obj2_dxT=dxT2)
    X, Yhat, nll = ref.filter_nll(Y)
    out.update(flt_X=X, flt_Yhat=Yhat, flt_nll=nll)
    d = ref.d
    for l in range(L):
        c = ref.ihgp_consts(params[p * L + L + 1 + 3 * l:][:3])
        for k, v in c.items():
            out["c%d_%s" % (l, k)] = np.asarray(v)
        Xs, P, G = ref.ihgp_smoother(params[p * L + L + 1 + 3 * l:][:3], X[: min(T, 40), l, :])
        out["sm%d_Xs" % l], out["sm%d_P" % l], out["sm%d_G" % l] = Xs, P, G
    # single observations with missing entries (moihgp.h:150-178) and predict-only (moihgp.h:381-428)
    xs = rng.standard_normal((L, d))
    dxs = rng.standard_normal((L, 3, d))
    masks = [[0], list(range(0, p, 2))[: max(p - L, 1)], list(range(p))]
    for i, m in enumerate(masks):
        y = rng.standard_normal(p)
        y[m] = np.nan
        xn, yh, dxn = ref.step(xs, y, dxs)
        out["nan%d_y" % i], out["nan%d_xn" % i], out["nan%d_yh" % i], out["nan%d_dxn" % i] = y, xn, yh, dxn
    # predict-only: with threading the reference joins never-created threads (moihgp.h:402 under NDEBUG),
    # so it is driven through a non-threaded instance (the result does not depend on the flag)
    ref_nt = ref if not threading else RefMOIHGP(dt, p, L, kernel, False)
    if threading:
        ref_nt.update(params)
    xn, yh = ref_nt.step(xs)
    out.update(one_x=xs, one_dx=dxs, pred_xn=xn, pred_yh=yh)
    y = rng.standard_normal(p)
    l1, g1 = ref.negLogLikelihood(xs, y, dxs)
    l2 = ref.negLogLikelihood(xs, y)
    out.update(one_y=y, one_lik1=l1, one_grad=g1, one_lik2=l2)
    return out


def smoother_case(kernel, p, L, T, seed, dt=0.1):
    """IHGP::backwardSmoother (ihgp.h:103-114) of the reference over a WHOLE sequence of its own filtered states, for
    hyper-parameters where its gain is stable (LITERAL_STABLE), so that the fixture pins smoothed means and the
    smoothed covariance at full length rather than on a 40-step prefix."""
    rng = np.random.default_rng(seed)
    params = make_params(rng, p, L, kernel, LITERAL_STABLE)
    Y = make_data(rng, p, L, T, dt)
    ref = RefMOIHGP(dt, p, L, kernel, False)
    ref.update(params)
    X, _, nll = ref.filter_nll(Y)
    out = {"kernel": kernel, "p": p, "L": L, "T": T, "dt": dt, "params": params, "Y": Y, "flt_X": X, "flt_nll": nll}
    Xs = np.zeros_like(X)
    P = np.zeros((L, ref.d, ref.d))
    G = np.zeros((L, ref.d, ref.d))
    for l in range(L):
        Xs[:, l, :], P[l], G[l] = ref.ihgp_smoother(params[p * L + L + 1 + 3 * l:][:3], X[:, l, :])
    out.update(sm_Xs=Xs, sm_P=P, sm_G=G)
    return out


SMOOTHER_CASES = [
    # name, kernel, p, L, T, seed
    ("lit_m32_p8L4_T700", "Matern32", 8, 4, 700, 21),
    ("lit_m52_p16L8_T600", "Matern52", 16, 8, 600, 22),
    ("lit_m32_p5L3_T300", "Matern32", 5, 3, 300, 23),
]

CASES = [
    # name, kernel, threading, p, L, T, seed
    ("c1_m32_p2L1_T63", "Matern32", True, 2, 1, 63, 1234),     # BASELINE config 1 shape (example_regression.cpp)
    ("m32_p5L3_T40_thr", "Matern32", True, 5, 3, 40, 11),
    ("m32_p5L3_T40_nothr", "Matern32", False, 5, 3, 40, 11),
    ("m32_p8L4_T300", "Matern32", True, 8, 4, 300, 12),        # BASELINE config 2 shape, > 1 scan chunk
    ("m52_p6L2_T50_nothr", "Matern52", False, 6, 2, 50, 13),
    ("m52_p16L8_T520", "Matern52", True, 16, 8, 520, 14),      # BASELINE config 3 shape, 3 scan chunks
]

if __name__ == "__main__":
    build()
    if not ref_available():
        sys.exit("oracle/_ref is not built (needs /root/reference): cannot generate fixtures")
    os.makedirs(OUT, exist_ok=True)
    for name, kernel, thr, p, L, T, seed in CASES:
        np.savez_compressed(os.path.join(OUT, name + ".npz"), **case(kernel, thr, p, L, T, seed))
        print("wrote", name)
    out2 = os.path.join(OUT, "..", "golden_smoother")
    os.makedirs(out2, exist_ok=True)
    for name, kernel, p, L, T, seed in SMOOTHER_CASES:
        np.savez_compressed(os.path.join(out2, name + ".npz"), **smoother_case(kernel, p, L, T, seed))
        print("wrote", name)
