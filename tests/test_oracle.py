"""CPU tests of the oracle (no GPU): the Eigen-free restatement against
  (a) the golden fixtures generated from the reference itself (tests/golden, oracle/gen_golden.py),
  (b) the reference library directly, where oracle/_ref exists (this container, not the GPU box),
  (c) its own internal identities (rank-1 dU form, literal vs simplified paths, quirk semantics)."""
import numpy as np
import pytest

from conftest import TOL, golden_cases, load_golden, rel_err
from oracle.binding import OracleMOIHGP, RefMOIHGP, ref_available

KNAME = {32: "Matern32", 52: "Matern52"}


def _oracle_for(g):
    o = OracleMOIHGP(float(g["dt"]), int(g["p"]), int(g["L"]), str(g["kernel"]), bool(g["threading"]))
    o.update(g["params"])
    return o


@pytest.mark.parametrize("path", golden_cases(), ids=lambda p: p.split("/")[-1][:-4])
def test_oracle_matches_reference_golden(path):
    g = load_golden(path)
    o = _oracle_for(g)
    p, L, T = int(g["p"]), int(g["L"]), int(g["T"])
    assert rel_err(o.params, g["params_after_update"]) < 1e-12
    # objective loop (moihgp_regression.h:42-50): both the literal O(p^3L^2) dU loop and its rank-1 form
    for literal in (False, True):
        loss, grad, xT, dxT = o.objective(g["Y"], literal=literal)
        assert abs(loss - g["obj_loss"]) <= TOL * abs(g["obj_loss"])
        assert rel_err(grad, g["obj_grad"]) < TOL
        assert rel_err(xT[0], g["obj_xT"]) < TOL and rel_err(dxT[0], g["obj_dxT"]) < TOL
    n2 = max(T // 3, 2)
    loss, grad, xT, dxT = o.objective(g["Y"][:n2], g["x0"][None], g["dx0"][None])
    assert abs(loss - g["obj2_loss"]) <= TOL * abs(g["obj2_loss"])
    assert rel_err(grad, g["obj2_grad"]) < TOL
    assert rel_err(xT[0], g["obj2_xT"]) < TOL and rel_err(dxT[0], g["obj2_dxT"]) < TOL
    # filter loop + NLL (moihgp_regression.h:127-139, moihgp.h:614-688)
    r = o.filter_smoother_nll(g["Y"], smoother_mode=0, want_yhat=True)
    assert rel_err(r["X"][0], g["flt_X"]) < TOL
    assert rel_err(r["Yhat"][0], g["flt_Yhat"]) < TOL
    assert abs(r["nll"][0] - g["flt_nll"]) <= TOL * abs(g["flt_nll"])
    # steady-state members (ihgp.h:243-254) and the literal smoother (ihgp.h:103-114)
    for l in range(L):
        c = o.ihgp_consts(l)
        for k, v in c.items():
            ref = g["c%d_%s" % (l, k)]
            if np.max(np.abs(ref)) > 0:
                assert rel_err(np.asarray(v), np.asarray(ref)) < TOL, (l, k)
            else:
                assert np.max(np.abs(v)) == 0
        G, P = o.smoother_consts(l, 0)
        Xs = o.ihgp_smooth(l, 0, g["flt_X"][: min(T, 40), l, :])
        assert rel_err(G, g["sm%d_G" % l]) < TOL
        # P is iterate #100 of the reference's un-converged DLyap map (Q2; 2.5e162 for the default Matern-3/2 latent) and the
        # literal recursion grows like 6.39^n there (Q3) - both are still well-defined numbers and agree far inside TOL
        assert np.all(np.isfinite(g["sm%d_P" % l])) and np.all(np.isfinite(g["sm%d_Xs" % l]))
        assert rel_err(P, g["sm%d_P" % l]) < TOL
        assert rel_err(Xs, g["sm%d_Xs" % l]) < TOL
    # missing observations (moihgp.h:150-178), predict-only (moihgp.h:381-428), single-step likelihoods
    for i in range(3):
        xn, yh, dxn = o.step(g["one_x"], g["nan%d_y" % i], g["one_dx"])
        assert rel_err(xn, g["nan%d_xn" % i]) < TOL and rel_err(yh, g["nan%d_yh" % i]) < TOL and rel_err(dxn, g["nan%d_dxn" % i]) < TOL
    xn, yh = o.step(g["one_x"])
    assert rel_err(xn, g["pred_xn"]) < TOL and rel_err(yh, g["pred_yh"]) < TOL
    l1, g1 = o.negLogLikelihood(g["one_x"], g["one_y"], g["one_dx"])
    assert abs(l1 - g["one_lik1"]) <= TOL * abs(g["one_lik1"]) and rel_err(g1, g["one_grad"]) < TOL
    assert abs(o.negLogLikelihood(g["one_x"], g["one_y"]) - g["one_lik2"]) <= TOL * abs(g["one_lik2"])


def smoother_cases():
    import glob
    import os
    from conftest import ROOT
    return sorted(glob.glob(os.path.join(ROOT, "tests", "golden_smoother", "*.npz")))


@pytest.mark.parametrize("path", smoother_cases(), ids=lambda p: p.split("/")[-1][:-4])
def test_oracle_matches_reference_smoother_over_whole_sequences(path):
    """tests/golden_smoother: the reference's IHGP::backwardSmoother (ihgp.h:103-114) run over WHOLE sequences of its own
    filtered states, hyper-parameters with rho(G) < 1 (oracle/gen_golden.py: LITERAL_STABLE) - means, covariance, gain."""
    g = load_golden(path)
    L = int(g["L"])
    o = OracleMOIHGP(float(g["dt"]), int(g["p"]), L, str(g["kernel"]), False)
    o.update(g["params"])
    r = o.filter_smoother_nll(g["Y"], smoother_mode=0)
    assert rel_err(r["X"][0], g["flt_X"]) < TOL and abs(r["nll"][0] - g["flt_nll"]) <= TOL * abs(g["flt_nll"])
    assert rel_err(r["Xs"][0], g["sm_Xs"]) < TOL
    for l in range(L):
        G, P = o.smoother_consts(l, 0)
        assert max(abs(np.linalg.eigvals(G))) < 1.0
        assert rel_err(G, g["sm_G"][l]) < TOL and rel_err(P, g["sm_P"][l]) < TOL
        assert rel_err(r["Xs"][0][:, l], g["sm_Xs"][:, l]) < TOL, l


def test_survey_anchor_values():
    """SURVEY.md section 10 sanity anchors (Matern-3/2, params (1, 1, 0.1), dt = 0.1)."""
    o = OracleMOIHGP(0.1, 2, 1, "Matern32")
    o.update(np.array([1.0, 0.0, 1.0, 0.01, 1.0, 1.0, 0.1]))
    c = o.ihgp_consts(0)
    assert abs(c["S"] - 0.3625238487309506) < 1e-12
    assert np.allclose(c["K"], [0.7241560786964509, -1.0039882276641663], atol=1e-12)
    assert np.allclose(c["A"], [[0.9866245648897065, 0.0840965131393047], [-0.2522895394179141, 0.6953056978963876]], atol=1e-13)
    it = o.ihgp_iters(0)
    assert it[0] == 13 and it[1:4] == [100, 100, 100]       # DARE stops at 13; all three DLyap hit the limit (Q1, Q2)
    G, _ = o.smoother_consts(0, 0)
    assert abs(np.max(np.abs(np.linalg.eigvals(G))) - 6.390) < 1e-2   # literal smoother is unstable here (Q3)
    o52 = OracleMOIHGP(0.1, 2, 1, "Matern52")
    o52.update(np.array([1.0, 0.0, 1.0, 0.01, 1.0, 1.0, 0.1]))
    assert abs(o52.ihgp_consts(0)["S"] - 2.6888023285229576) < 1e-10
    assert o52.ihgp_iters(0)[0] == 100                       # un-converged DARE


def test_threading_flag_changes_loss_only():
    """Q5: the gradient version adds the per-latent NLL terms only when `threading` (moihgp.h:588 vs :601)."""
    rng = np.random.default_rng(0)
    p, L, T = 5, 3, 40
    params = np.concatenate([(np.eye(p, L) + 0.3 * rng.standard_normal((p, L))).ravel(), [1.3, 0.8, 2.0, 0.05, 1, 1, .1, 2, .5, .2, .7, 3, .05]])
    Y = rng.standard_normal((T, p))
    res = {}
    for thr in (False, True):
        o = OracleMOIHGP(0.1, p, L, "Matern32", thr)
        o.update(params)
        res[thr] = o.objective(Y)
    assert abs(res[True][0] - 565.9793953332739) < 1e-7      # SURVEY section 10 anchors
    assert abs(res[False][0] - 403.68780815657374) < 1e-7
    assert rel_err(res[True][1], res[False][1]) == 0.0
    # L < 2 forces threading off (moihgp.h:128-135)
    a, b = OracleMOIHGP(0.1, 2, 1, "Matern32", True), OracleMOIHGP(0.1, 2, 1, "Matern32", False)
    Y1 = rng.standard_normal((10, 2))
    assert a.objective(Y1)[0] == b.objective(Y1)[0]


def test_update_is_idempotent_on_U():
    rng = np.random.default_rng(2)
    o = OracleMOIHGP(0.1, 6, 3, "Matern32")
    params = np.concatenate([rng.standard_normal(18), [1.0, 2.0, 0.5, 0.02], [1, 1, .1] * 3])
    o.update(params)
    U1 = o.U
    assert np.allclose(U1.T @ U1, np.eye(3), atol=1e-13)
    o.update(o.params)
    assert rel_err(o.U, U1) < 1e-13


def test_smoother_modes_on_fixed_interval():
    """rts_correct: last smoothed state equals the filtered one; literal follows ihgp.h:108-113."""
    o = OracleMOIHGP(0.1, 2, 1, "Matern52")
    o.update(np.array([1.0, 0.0, 1.0, 0.01, 0.5, 0.5, 0.1]))
    rng = np.random.default_rng(3)
    X = rng.standard_normal((25, 3))
    for mode in (0, 1):
        Xs = o.ihgp_smooth(0, mode, X)
        assert np.array_equal(Xs[-1], X[-1])
        G, _ = o.smoother_consts(0, mode)
        A = o.ihgp_consts(0)["A"]
        j = 10
        want = X[j + 1] + G @ Xs[j + 1] - A @ X[j + 1] if mode == 0 else X[j] + G @ (Xs[j + 1] - A @ X[j])
        assert rel_err(Xs[j], want) < 1e-13


@pytest.mark.skipif(not ref_available(), reason="oracle/_ref not built (no /root/reference on this box)")
@pytest.mark.parametrize("kernel", ["Matern32", "Matern52"])
@pytest.mark.parametrize("threading", [False, True])
def test_oracle_matches_reference_library(kernel, threading):
    """Direct comparison with the shim-compiled reference on fresh random inputs."""
    from oracle.gen_golden import make_data, make_params
    rng = np.random.default_rng(77)
    p, L, T = 7, 3, 30
    params = make_params(rng, p, L, kernel)
    Y = make_data(rng, p, L, T)
    o = OracleMOIHGP(0.1, p, L, kernel, threading)
    r = RefMOIHGP(0.1, p, L, kernel, threading)
    o.update(params)
    r.update(params)
    lo, go, xo, dxo = o.objective(Y, literal=True)
    lr, gr, xr, dxr = r.objective(Y)
    assert abs(lo - lr) <= TOL * abs(lr) and rel_err(go, gr) < TOL
    assert rel_err(xo[0], xr) < TOL and rel_err(dxo[0], dxr) < TOL
    Xr, Yhr, nr = r.filter_nll(Y)
    res = o.filter_smoother_nll(Y, smoother_mode=-1, want_yhat=True)
    assert rel_err(res["X"][0], Xr) < TOL and rel_err(res["Yhat"][0], Yhr) < TOL and abs(res["nll"][0] - nr) <= TOL * abs(nr)
