"""GPU parity tests (B200): the CUDA path, called through the C ABI / the reference-named host
mirror, against (a) the committed mpmath golden fixtures, (b) the numpy oracle on seeded inputs
at sizes it finishes in seconds, (c) size-independent properties at BASELINE.json's full sizes.

Tolerances are north_star's: NLML and gradients 1e-9 relative, predictions 1e-8."""
import json
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

from gpsurvival_b200 import gp, synth, _lib          # noqa: E402
from oracle import gp_oracle as orc                  # noqa: E402

GOLD = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "golden_small.json")))
TOL_NLML = 1e-9
TOL_GRAD = 1e-9
TOL_PRED = 1e-8


def relmax(a, b):
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    return float(np.max(np.abs(a - b)) / max(float(np.max(np.abs(b))), 1e-300))


@pytest.fixture(scope="module")
def fit0():
    f = gp.Fit(0)
    yield f
    f.close()


def _make_fit(X, y, ev, cov_form, mean_form, noise_corr, lnc=None, ard_mode="intended", extra_param=None):
    fit = gp.Fit(0)
    fit.set_options(cov_form, mean_form, noise_corr, ard_mode, extra_param)
    fit.set_data(X, y, ev)
    if noise_corr == "noiseCorrVec":
        fit.set_noise_corr(lnc)
    return fit


# ---- golden fixtures (mpmath, 50 digits) -------------------------------------------------------
@pytest.mark.parametrize("c", GOLD["cases"], ids=[c["name"] for c in GOLD["cases"]])
def test_golden_fixture(c):
    X, y, ev = np.array(c["X"]), np.array(c["y"]), np.array(c["events"])
    par = np.array(c["par"])
    fit = _make_fit(X, y, ev, c["cov_form"], c["mean_form"], c["noise_corr"], c["log_sc2_vec"],
                    extra_param=c.get("extra_param"))
    assert abs(fit.nlml(par) - c["nlml"]) <= TOL_NLML * abs(c["nlml"])
    if c["grad"] is None:                       # Matern: no gradient in the reference, none here
        with pytest.raises(gp.GpsError) as e:
            fit.nlml_grad(par)
        assert e.value.code == _lib.GPS_ERR_UNSUPPORTED
    else:
        f, g = fit.nlml_grad(par)
        assert abs(f - c["nlml"]) <= TOL_NLML * abs(c["nlml"])
        assert relmax(g, c["grad"]) <= TOL_GRAD
    mu, vf, vd = fit.predict(par, np.array(c["Xstar"]))
    assert relmax(mu, c["pred_mean"]) <= TOL_PRED
    sf2 = float(np.exp(par[1]))
    assert np.max(np.abs(vf - np.array(c["pred_varf"]))) <= TOL_PRED * sf2
    assert np.max(np.abs(vd - np.array(c["pred_vard"]))) <= TOL_PRED * sf2
    fit.close()


# ---- CovFunc ---------------------------------------------------------------------------------
@pytest.mark.parametrize("n1,n2,d", [(100, 0, 1), (301, 0, 20), (285, 0, 3000), (130, 77, 5), (64, 200, 700),
                                     (1, 1, 3), (2, 0, 1)])
@pytest.mark.parametrize("cov_form", ["SqExp", "ARD"])
@pytest.mark.parametrize("ard_mode", ["intended", "as_written"])
def test_cov_matches_oracle(fit0, n1, n2, d, cov_form, ard_mode):
    rng = np.random.default_rng(n1 * 7 + d)
    x1 = rng.standard_normal((n1, d))
    x2 = None if n2 == 0 else rng.standard_normal((n2, d))
    ell = np.array([1.7 * np.sqrt(d)]) if cov_form == "SqExp" else 1.7 * np.sqrt(d) * (1 + 0.3 * rng.random(d))
    K = fit0.cov(x1, x2, 0.8, ell, cov_form, ard_mode)
    Ko = orc.cov_func(x1, x1 if x2 is None else x2, 0.8, ell, cov_form, ard_mode)
    assert np.max(np.abs(K - Ko)) <= 1e-12
    if x2 is None:
        assert np.array_equal(K, K.T) and np.all(np.diag(K) == 0.8)      # exact symmetry and diagonal


def test_cov_reference_named_entry_point():
    rng = np.random.default_rng(5)
    x = rng.standard_normal((50, 3))
    xs = rng.standard_normal((20, 3))
    K = gp.CovFunc(x, x, x, None, None, 0.5, 1.3, "SqExp")
    Ks = gp.CovFunc(x, xs, x, None, None, 0.5, 1.3, "SqExp")
    assert np.max(np.abs(K - orc.cov_func(x, x, 0.5, [1.3]))) <= 1e-12
    assert np.max(np.abs(Ks - orc.cov_func(x, xs, 0.5, [1.3]))) <= 1e-12
    Km = gp.CovFunc(x, x, x, None, 3, 0.5, 1.3, "Matern")
    assert np.max(np.abs(Km - orc.cov_func(x, x, 0.5, [1.3], "Matern", extra_param=3))) <= 1e-11
    with pytest.raises(gp.GpsError):
        gp.CovFunc(x, x, x, None, None, 0.5, 1.3, "RF")


# ---- NLML / gradient / finalize / predict against the oracle -----------------------------------
CASES = [
    # name, n, m, d, c, cov, mean, noise_corr, seed
    ("C1 SqExp GPS1", 100, 50, 1, 40, "SqExp", "Zero", False, 11),
    ("SqExp GPS2 odd n", 257, 40, 4, 100, "SqExp", "Zero", "noiseCorrLearned", 12),
    ("C2-shaped ARD GPS3 n=300", 300, 64, 20, 150, "ARD", "Zero", "noiseCorrVec", 13),
    ("ARD Linear GPS2", 201, 33, 6, 90, "ARD", "Linear", "noiseCorrLearned", 14),
    ("SqExp Linear GPS1", 150, 20, 3, 60, "SqExp", "Linear", False, 15),
    ("C3-shaped ARD GPS3 d=1500", 285, 57, 1500, 142, "ARD", "Zero", "noiseCorrVec", 16),
    ("ARD GPS3 n=700", 700, 120, 20, 350, "ARD", "Zero", "noiseCorrVec", 17),
    ("tiny n=3", 3, 2, 2, 1, "ARD", "Zero", "noiseCorrVec", 18),
    ("one leaf n=64", 64, 5, 2, 20, "SqExp", "Zero", False, 19),
    ("n=65", 65, 5, 2, 20, "SqExp", "Zero", "noiseCorrLearned", 20),
    ("no censored rows GPS3", 90, 9, 3, 0, "ARD", "Zero", "noiseCorrVec", 21),
]


@pytest.mark.parametrize("case", CASES, ids=[c[0] for c in CASES])
def test_eval_matches_oracle(case):
    name, n, m, d, c, cov_form, mean_form, noise_corr, seed = case
    pb = synth.make_problem(n, m, d, c, seed)
    X, y, ev, Xs = pb["trainingData"], np.log(pb["trainingTargets"]), pb["events"], pb["testData"]
    rng = np.random.default_rng(seed)
    p = d if cov_form == "ARD" else 1
    par = [np.log(0.2), np.log(0.8)] + list(np.log(1.7 * np.sqrt(d)) + 0.2 * rng.standard_normal(p))
    if mean_form == "Linear":
        par += list(0.1 * rng.standard_normal(d + 1))
    lnc = None
    if noise_corr == "noiseCorrLearned":
        par += [np.log(0.3)]
    if noise_corr == "noiseCorrVec":
        lnc = np.log(0.2) + 0.5 * rng.standard_normal(int(np.sum(ev == 0)))
    par = np.array(par)
    kw = dict(cov_form=cov_form, mean_form=mean_form, noise_corr=noise_corr, log_hyp_noise_corr=lnc)
    fit = _make_fit(X, y, ev, cov_form, mean_form, noise_corr, lnc)
    f_o = orc.nlml(par, X, y, ev, **kw)
    g_o = orc.nlml_grad(par, X, y, ev, **kw)
    for _ in range(3):                       # direct launch, graph capture, graph replay
        assert abs(fit.nlml(par) - f_o) <= TOL_NLML * abs(f_o)
    for _ in range(3):                       # cached factorisation, then graphs
        f, g = fit.nlml_grad(par)
        assert abs(f - f_o) <= TOL_NLML * abs(f_o)
        assert relmax(g, g_o) <= TOL_GRAD
    f2, g2 = fit.nlml_grad(par + 0.01)       # uncached
    assert relmax(g2, orc.nlml_grad(par + 0.01, X, y, ev, **kw)) <= TOL_GRAD
    fin = fit.finalize(par, want_K=True, want_L=True)
    reg = orc.regress_finalize(par, X, y, ev, cov_form, mean_form, noise_corr, lnc)
    assert np.max(np.abs(fin["K"] - reg["K"])) <= 1e-12
    assert relmax(fin["L"], reg["L"]) <= 1e-9
    assert relmax(fin["alpha"], reg["alpha"]) <= TOL_PRED
    assert relmax(fin["meanTraining"], reg["meanTraining"]) <= 1e-12 or np.all(reg["meanTraining"] == 0)
    assert abs(fin["logMarLik"] - reg["logMarLik"]) <= TOL_NLML * abs(reg["logMarLik"])
    assert np.array_equal(np.triu(fin["L"], 1), np.zeros_like(fin["L"]))
    mu, vf, vd = fit.predict(par, Xs)
    reg["logHypChosen"] = {"noise": par[0], "func": par[1], "length": par[2:2 + p],
                           "mean": par[2 + p:2 + p + d + 1] if mean_form == "Linear" else np.zeros(d + 1)}
    pr = orc.predict(reg, X, y, ev, Xs, cov_form, mean_form, noise_corr,
                     par[-1:] if noise_corr == "noiseCorrLearned" else lnc)
    assert relmax(mu, pr["funcMeanPred"]) <= TOL_PRED
    assert np.max(np.abs(vf - pr["varFunction"])) <= TOL_PRED * 0.8
    assert np.max(np.abs(vd - pr["varData"])) <= TOL_PRED * 0.8
    fit.close()


def test_as_written_ard_nlml_matches_oracle():
    pb = synth.make_problem(120, 10, 7, 50, 33)
    X, y, ev = pb["trainingData"], np.log(pb["trainingTargets"]), pb["events"]
    par = np.concatenate([[np.log(0.2), np.log(0.8)], np.log(1.7 * np.sqrt(7)) + 0.1 * np.arange(7)])
    fit = _make_fit(X, y, ev, "ARD", "Zero", False, ard_mode="as_written")
    f_o = orc.nlml(par, X, y, ev, cov_form="ARD", ard_mode="as_written")
    assert abs(fit.nlml(par) - f_o) <= TOL_NLML * abs(f_o)
    with pytest.raises(gp.GpsError) as e:
        fit.nlml_grad(par)
    assert e.value.code == _lib.GPS_ERR_UNSUPPORTED
    fit.close()


# ---- reference-named stateless entry points ------------------------------------------------------
def test_for_optimisation_log_signature_and_caching():
    pb = synth.make_problem(80, 10, 2, 30, 5)
    X, y, ev = pb["trainingData"], np.log(pb["trainingTargets"]), pb["events"]
    par = np.log([0.2, 0.8, 2.0])
    args = (X, y, ev, "Zero", 2, None, "SqExp", 80, False, None, False)
    v = gp.ForOptimisationLog(par, *args)
    assert v.shape == (1, 1) and abs(v[0, 0] - orc.nlml(par, X, y, ev)) <= TOL_NLML * abs(v[0, 0])
    g = gp.OptimisationGradientLog(par, *args)
    assert g.shape == (3, 1) and relmax(g.ravel(), orc.nlml_grad(par, X, y, ev)) <= TOL_GRAD
    y2 = y + 0.1                                     # new targets: same X stays resident
    v2 = gp.ForOptimisationLog(par, X, y2, ev, "Zero", 2, None, "SqExp", 80, False, None, False)
    assert abs(v2[0, 0] - orc.nlml(par, X, y2, ev)) <= TOL_NLML * abs(v2[0, 0])
    with pytest.raises(gp.GpsError):
        gp.ForOptimisationLog(par, X, y, ev, "Zero", 2, None, "SqExp", 80, False, None, True)   # imposePriors


# ---- Adjust ----------------------------------------------------------------------------------
def test_adjust_matches_oracle_and_fixture(fit0):
    rng = np.random.default_rng(3)
    c = 20000
    mu = rng.standard_normal(c)
    s = 0.05 + rng.random(c)
    alpha = rng.uniform(-8, 7.5, c)
    a = mu + alpha * s
    em, ev = fit0.adjust(a, mu, s)
    eo, vo = orc.adjust(a, mu, s)
    al = (a - mu) / s
    lo = al < 5.0                                    # parity bar 1e-8 where 1 - Phi is well conditioned
    assert np.max(np.abs(em[lo] - eo[lo]) / np.maximum(np.abs(eo[lo]), 1.0)) <= 1e-8
    assert np.max(np.abs(ev[lo] - vo[lo]) / np.maximum(np.abs(vo[lo]), 1e-3)) <= 1e-8
    mid = (al >= 5.0) & (al < 6.5)                    # reported separately
    assert np.max(np.abs(em[mid] - eo[mid]) / np.maximum(np.abs(eo[mid]), 1.0)) <= 1e-5
    assert np.mean((ev == 0) != (vo == 0)) <= 0.002   # the hard (a, 0) branch may flip within 1 ulp of 1e-12
    assert np.all(em >= a - 1e-9 * np.maximum(1.0, np.abs(a)))
    rows = GOLD["adjust"]
    em, ev = fit0.adjust([r["a"] for r in rows], [r["mu"] for r in rows], [r["s"] for r in rows])
    for r, m_, v_ in zip(rows, em, ev):
        if r["alpha"] < 5:
            assert abs(m_ - r["as_written_mean"]) <= 1e-8 * max(1.0, abs(r["as_written_mean"]))
            assert abs(v_ - r["as_written_var"]) <= 1e-8 * max(1e-3, abs(r["as_written_var"]))
    out = gp.AdjustTrainingSurvivalMeanVariance(0.5, 0.0, 1.0)
    assert set(out) == {"mu", "sigma"} and abs(out["mu"] - orc.adjust_one(0.5, 0.0, 1.0)[0]) < 1e-12
    assert fit0.adjust([], [], [])[0].size == 0       # empty input


# ---- whole fits: GPS loop trajectory ---------------------------------------------------------------
@pytest.mark.parametrize("model,nc", [("GPS1", False), ("GPS2", "noiseCorrLearned"), ("GPS3", "noiseCorrVec"),
                                      ("GPR", None)])
def test_apply_gp_c1_matches_oracle(model, nc):
    cfg = synth.CONFIGS["C1"]
    pb = synth.make_config_problem(cfg)
    params = dict(modelType="survival" if nc is not None else "non-survival", noiseCorr=nc if nc is not None else False,
                  optimType="Nelder-Mead", maxit=300, maxitSurvival=10, meanFuncForm="Zero", covFuncForm="SqExp",
                  tolerance=1e-5 * 100 * 0.4, maxCount=20, logHypStart=synth.start_log_hyp(1, "SqExp"),
                  burnIn=False, imposePriors=False)
    tts = pb
    if nc is None:                                    # RemoveCensored (runSyntheticExp1.R:195)
        keep = pb["events"] == 1
        tts = dict(pb, trainingData=pb["trainingData"][keep], trainingTargets=pb["trainingTargets"][keep],
                   events=pb["events"][keep])
    r = gp.ApplyGP(tts, {"dataSource": "Generate"}, params)
    ro = orc.apply_gp(tts, params)
    if "count" in ro:
        assert r["count"] == ro["count"]
    # whole-fit agreement is a secondary, looser check (simplex comparisons are 1-ulp sensitive)
    assert relmax(r["funcMeanPred"].ravel(), ro["funcMeanPred"]) <= 1e-6
    assert relmax(r["varData"], ro["varData"]) <= 1e-6
    assert abs(r["objective"] - ro["objective"]) <= 1e-6 * abs(ro["objective"])


def test_gp_regression_bfgs_ard():
    pb = synth.make_problem(150, 20, 4, 60, 8)
    y = np.log(pb["trainingTargets"])
    params = dict(modelType="survival", noiseCorr="noiseCorrVec", optimType="BFGS", maxit=50, maxitSurvival=8,
                  meanFuncForm="Zero", covFuncForm="ARD", imposePriors=False)
    lh = synth.start_log_hyp(4, "ARD")
    lnc = np.full(int(np.sum(pb["events"] == 0)), lh["noise"])
    data = {"trainingData": pb["trainingData"], "trainingTargets": y, "events": pb["events"]}
    reg = gp.GPRegression(lh, params, data, lnc, want_matrices=True)
    rego = orc.gp_regression(lh, params, data, lnc)
    assert abs(reg["objective"] - rego["objective"]) <= 1e-7 * abs(rego["objective"])
    assert relmax(reg["alpha"].ravel(), rego["alpha"]) <= 1e-5
    assert reg["objective"] < orc.nlml(orc.flatten_log_hyp(lh, 4, "Zero", "noiseCorrVec", lnc), pb["trainingData"], y,
                                       pb["events"], cov_form="ARD", noise_corr="noiseCorrVec", log_hyp_noise_corr=lnc)
    assert set(("logHyp", "meanTraining", "K", "L", "alpha", "logHypChosen", "logMarLik", "objective",
                "logHypNoiseCorrection")) <= set(reg)


# ---- batches -------------------------------------------------------------------------------------
def test_batched_fits_equal_single_fits():
    B, n, m, d = 5, 90, 12, 6
    pbs = [synth.make_problem(n, m, d, 40, 100 + b) for b in range(B)]
    # the number of censored rows must be equal across a batch for GPS3; use GPS2 here
    X = np.stack([p["trainingData"] for p in pbs])
    y = np.stack([np.log(p["trainingTargets"]) for p in pbs])
    ev = np.stack([p["events"] for p in pbs])
    Xs = np.stack([p["testData"] for p in pbs])
    rng = np.random.default_rng(0)
    par = np.tile(np.concatenate([[np.log(0.2), np.log(0.8)], np.full(d, np.log(1.7 * np.sqrt(d))), [np.log(0.3)]]),
                  (B, 1)) + 0.2 * rng.standard_normal((B, d + 3))
    bf = gp.Fit(0, batch=B)
    bf.set_options("ARD", "Zero", "noiseCorrLearned", "intended")
    bf.set_data(X, y, ev)
    f, g = bf.nlml_grad(par)
    f_again, g_again = bf.nlml_grad(par + 0.0)
    assert np.array_equal(f, f_again) and np.array_equal(g, g_again)       # run-to-run reproducible
    mu, vf, vd = bf.predict(par, Xs)
    for b in range(B):
        kw = dict(cov_form="ARD", noise_corr="noiseCorrLearned")
        fo = orc.nlml(par[b], X[b], y[b], ev[b], **kw)
        assert abs(f[b] - fo) <= TOL_NLML * abs(fo)
        assert relmax(g[b], orc.nlml_grad(par[b], X[b], y[b], ev[b], **kw)) <= TOL_GRAD
        reg = orc.regress_finalize(par[b], X[b], y[b], ev[b], "ARD", "Zero", "noiseCorrLearned")
        reg["logHypChosen"] = {"noise": par[b, 0], "func": par[b, 1], "length": par[b, 2:2 + d], "mean": np.zeros(d + 1)}
        pr = orc.predict(reg, X[b], y[b], ev[b], Xs[b], "ARD", "Zero", "noiseCorrLearned", par[b, -1:])
        assert relmax(mu[b], pr["funcMeanPred"]) <= TOL_PRED
        assert np.max(np.abs(vd[b] - pr["varData"])) <= TOL_PRED
    bf.close()


# ---- error behaviour -------------------------------------------------------------------------------
def test_error_paths():
    pb = synth.make_problem(40, 5, 2, 10, 1)
    X, y, ev = pb["trainingData"], np.log(pb["trainingTargets"]), pb["events"]
    fit = gp.Fit(0)
    with pytest.raises(gp.GpsError) as e:
        fit.nlml(np.zeros(3))                          # no data yet
    assert e.value.code == _lib.GPS_ERR_STATE
    fit.set_data(X, y, ev)
    with pytest.raises(gp.GpsError) as e:
        fit.nlml(np.zeros(5))                          # wrong length
    assert e.value.code == _lib.GPS_ERR_ARG
    with pytest.raises(gp.GpsError) as e:
        fit.nlml(np.array([0.0, np.nan, 0.0]))
    assert e.value.code == _lib.GPS_ERR_ARG
    with pytest.raises(gp.GpsError):
        fit.set_options("Matern")
    # duplicate rows + vanishing noise => numerically singular Ky => NOT_PD with a pivot index
    Xd = np.vstack([X, X])
    fit.set_data(Xd, np.concatenate([y, y]), np.concatenate([ev, ev]))
    fit.set_near_pd(False)                              # without the nearPD repair (tests/test_gpu_round2.py covers it)
    with pytest.raises(gp.NotPositiveDefinite) as e:
        fit.nlml(np.array([-800.0, 0.0, 0.0]))
    assert e.value.code == _lib.GPS_ERR_NOT_PD and 1 <= e.value.pivot <= 80
    assert abs(fit.nlml(np.log([0.2, 0.8, 2.0])) - orc.nlml(np.log([0.2, 0.8, 2.0]), Xd, np.concatenate([y, y]),
                                                             np.concatenate([ev, ev]))) < 1e-7   # handle still usable
    fit.set_options("SqExp", "Zero", "noiseCorrVec")
    with pytest.raises(gp.GpsError):
        fit.set_noise_corr(np.zeros(3))                # wrong number of censored rows
    fit.close()


# ---- size-independent properties at BASELINE.json's full sizes ---------------------------------------
@pytest.mark.parametrize("name", ["C2", "C3"])
def test_full_size_properties(name):
    cfg = synth.CONFIGS[name]
    pb = synth.make_config_problem(cfg)
    X, y, ev, Xs = pb["trainingData"], np.log(pb["trainingTargets"]), pb["events"], pb["testData"]
    n, d = X.shape
    lnc = np.full(int(np.sum(ev == 0)), np.log(0.2))
    fit = _make_fit(X, y, ev, cfg.cov_form, "Zero", cfg.noise_corr, lnc)
    par = synth.start_par(cfg, d=d)
    f, g = fit.nlml_grad(par)
    fin = fit.finalize(par, want_K=True, want_L=True)
    K, L, alpha = fin["K"], fin["L"], fin["alpha"]
    noise = np.full(n, 0.2)
    noise[ev == 0] += np.exp(lnc)
    Ky = K + np.diag(noise)
    assert np.max(np.abs(L @ L.T - Ky)) <= 1e-12 * n                   # L L' = Ky
    assert np.max(np.abs(Ky @ alpha - y)) <= 1e-9 * np.max(np.abs(y)) * n   # Ky alpha = y
    assert abs(-fin["logMarLik"] - f) <= 1e-12 * abs(f)
    logdet = 2 * np.sum(np.log(np.diag(L)))
    assert abs(f - (0.5 * y @ alpha + 0.5 * logdet + 0.5 * n * np.log(2 * np.pi))) <= 1e-10 * abs(f)
    # directional derivative of the GPU NLML along a random direction vs the analytic gradient
    rng = np.random.default_rng(0)
    v = rng.standard_normal(par.size)
    v /= np.linalg.norm(v)
    h = 1e-4
    fd = (fit.nlml(par + h * v) - fit.nlml(par - h * v)) / (2 * h)
    assert abs(fd - g @ v) <= 1e-6 * max(1.0, abs(g @ v))
    mu, vf, vd = fit.predict(par, Xs)
    assert np.all(vf >= -1e-9) and np.all(vf <= 0.8 + 1e-9) and np.allclose(vd - vf, 0.2, atol=1e-12)
    # against the oracle's K-space formulas evaluated from the GPU's own K (cheap at any size)
    Ks = gp.CovFunc(X, Xs, X, None, None, 0.8, np.exp(par[2:2 + d]), cfg.cov_form)
    assert relmax(mu, Ks.T @ alpha) <= TOL_PRED
    fit.close()


# ---- batched whole fits (restarts / folds in lock-step on one GPU) -------------------------------------
@pytest.mark.parametrize("model,nc,method", [("GPS1", False, "Nelder-Mead"), ("GPS2", "noiseCorrLearned", "Nelder-Mead"),
                                             ("GPS3", "noiseCorrVec", "Nelder-Mead"), ("GPS3", "noiseCorrVec", "BFGS")])
def test_batched_fits_follow_sequential_apply_gp(model, nc, method):
    from gpsurvival_b200 import batchfit
    B = 4
    pbs = [synth.make_problem(70, 15, 2, 28, 900 + b) for b in range(B)]
    params = dict(modelType="survival", noiseCorr=nc, optimType=method, maxit=100, maxitSurvival=10 if method == "Nelder-Mead" else 4,
                  meanFuncForm="Zero", covFuncForm="ARD" if method == "BFGS" else "SqExp", tolerance=0.02, maxCount=9,
                  logHypStart=synth.start_log_hyp(2, "ARD" if method == "BFGS" else "SqExp"), burnIn=False, imposePriors=False)
    res = batchfit.apply_gp_batch(pbs, params)
    counts = set()
    for b in range(B):
        gp.SESSION.reset()
        ref = gp.ApplyGP(pbs[b], {"dataSource": "Generate"}, params)
        counts.add(ref["count"])
        assert res[b]["count"] == ref["count"]
        assert relmax(res[b]["objectiveTable"], ref["objectiveTable"]) <= 1e-7
        assert relmax(res[b]["funcMeanPred"], ref["funcMeanPred"]) <= 1e-6
        assert relmax(res[b]["varData"], ref["varData"]) <= 1e-6
        assert relmax(res[b]["trainingSetPredictions"], ref["trainingSetPredictions"]) <= 1e-6
    gp.SESSION.reset()


def test_not_pd_is_per_fit_and_non_fatal_for_optimisers():
    pb = synth.make_problem(40, 5, 2, 10, 1)
    X, y, ev = pb["trainingData"], np.log(pb["trainingTargets"]), pb["events"]
    Xd = np.vstack([X, X])                       # duplicate rows: singular without noise
    yd, evd = np.concatenate([y, y]), np.concatenate([ev, ev])
    bf = gp.Fit(0, batch=2)
    bf.set_near_pd(False)
    bf.set_data(np.stack([Xd, Xd]), np.stack([yd, yd]), np.stack([evd, evd]))
    par = np.array([[-800.0, 0.0, 0.0], np.log([0.2, 0.8, 2.0])])
    with pytest.raises(gp.NotPositiveDefinite):
        bf.nlml(par)
    f = bf.nlml(par, not_pd="inf")
    assert np.isinf(f[0]) and abs(f[1] - orc.nlml(par[1], Xd, yd, evd)) <= 1e-9 * abs(f[1])
    piv = bf.pivots()
    assert piv[0] >= 1 and piv[1] == 0
    f2, g2 = bf.nlml_grad(par, not_pd="inf")
    assert np.isinf(f2[0]) and np.all(np.isnan(g2[0])) and relmax(g2[1], orc.nlml_grad(par[1], Xd, yd, evd)) <= 1e-9
    bf.close()


# ---- Matern covariances (CovFunc.R:33-37): cov / NLML / finalize / predict / whole Nelder-Mead fit ------------
@pytest.mark.parametrize("nu", [1, 3, 5])
def test_matern_matches_oracle(fit0, nu):
    rng = np.random.default_rng(nu)
    x1, x2 = rng.standard_normal((150, 4)), rng.standard_normal((60, 4))
    K = fit0.cov(x1, None, 0.8, [2.0], "Matern", extra_param=nu)
    Ks = fit0.cov(x1, x2, 0.8, [2.0], "Matern", extra_param=nu)
    # r = sqrt(r2): the Gram form's absolute error on r2 (~1e-16 * s) becomes ~eps*s/(2r) on r
    assert np.max(np.abs(K - orc.cov_func(x1, x1, 0.8, [2.0], "Matern", extra_param=nu))) <= 1e-11
    assert np.max(np.abs(Ks - orc.cov_func(x1, x2, 0.8, [2.0], "Matern", extra_param=nu))) <= 1e-11
    assert np.all(np.diag(K) == 0.8)
    pb = synth.make_problem(130, 30, 3, 50, 70 + nu)
    X, y, ev, Xs = pb["trainingData"], np.log(pb["trainingTargets"]), pb["events"], pb["testData"]
    lnc = np.full(50, np.log(0.3))
    par = np.log([0.2, 0.8, 2.5])
    fit = _make_fit(X, y, ev, "Matern", "Zero", "noiseCorrVec", lnc, extra_param=nu)
    kw = dict(cov_form="Matern", noise_corr="noiseCorrVec", log_hyp_noise_corr=lnc, extra_param=nu)
    fo = orc.nlml(par, X, y, ev, **kw)
    assert abs(fit.nlml(par) - fo) <= TOL_NLML * abs(fo)
    # gradient: derived (the reference has none, OptimisationGradientLog.R:6) — against the oracle's W-form and
    # against central differences of the oracle NLML
    f, g = fit.nlml_grad(par)
    go = orc.nlml_grad(par, X, y, ev, **kw)
    assert relmax(g, go) <= TOL_GRAD
    for q in range(3):
        hq = np.zeros(3)
        hq[q] = 1e-5
        fd = (orc.nlml(par + hq, X, y, ev, **kw) - orc.nlml(par - hq, X, y, ev, **kw)) / 2e-5
        assert abs(fd - g[q]) <= 1e-6 * max(1.0, abs(fd))
    fin = fit.finalize(par, want_K=True, want_L=True)
    reg = orc.regress_finalize(par, X, y, ev, "Matern", "Zero", "noiseCorrVec", lnc, extra_param=nu)
    assert relmax(fin["L"], reg["L"]) <= 1e-9 and relmax(fin["alpha"], reg["alpha"]) <= TOL_PRED
    mu, vf, vd = fit.predict(par, Xs)
    reg["logHypChosen"] = {"noise": par[0], "func": par[1], "length": par[2:], "mean": np.zeros(4)}
    pr = orc.predict(reg, X, y, ev, Xs, "Matern", "Zero", "noiseCorrVec", lnc, extra_param=nu)
    assert relmax(mu, pr["funcMeanPred"]) <= TOL_PRED and np.max(np.abs(vd - pr["varData"])) <= TOL_PRED
    fit.close()


def test_apply_gp_matern_nelder_mead():
    pb = synth.make_config_problem(synth.CONFIGS["C1"], scale=0.6)
    params = dict(modelType="survival", noiseCorr="noiseCorrVec", optimType="Nelder-Mead", maxit=100, maxitSurvival=10,
                  meanFuncForm="Zero", covFuncForm="Matern", extraParam=3, tolerance=1e-3, maxCount=8,
                  logHypStart=synth.start_log_hyp(1, "SqExp"), burnIn=False, imposePriors=False)
    gp.SESSION.reset()
    r = gp.ApplyGP(pb, {"dataSource": "Generate"}, params)
    ro = orc.apply_gp(pb, params)
    assert r["count"] == ro["count"]
    assert relmax(r["funcMeanPred"].ravel(), ro["funcMeanPred"]) <= 1e-6
    gp.SESSION.reset()


def test_full_size_c4_batch_properties():
    """C4 shape (n=400, d=2000, ARD GPS3) as a batch of 8 fits: batch entries equal lone fits,
    gradient agrees with a directional finite difference of the batched NLML."""
    cfg = synth.CONFIGS["C4"]
    B = 8
    pbs = [synth.make_config_problem(cfg, fit_index=i) for i in range(2)]
    X = np.stack([pbs[i % 2]["trainingData"] for i in range(B)])
    y = np.stack([np.log(pbs[i % 2]["trainingTargets"]) for i in range(B)])
    ev = np.stack([pbs[i % 2]["events"] for i in range(B)])
    d = X.shape[2]
    c = int(np.sum(ev[0] == 0))
    lnc = np.full((B, c), np.log(0.2))
    rng = np.random.default_rng(2)
    par = np.tile(synth.start_par(cfg, d=d), (B, 1)) + 0.05 * rng.standard_normal((B, 2 + d))
    bf = gp.Fit(0, batch=B)
    bf.set_options("ARD", "Zero", "noiseCorrVec")
    bf.set_data(X, y, ev)
    bf.set_noise_corr(lnc)
    f, g = bf.nlml_grad(par)
    v = rng.standard_normal((B, 2 + d))
    v /= np.linalg.norm(v, axis=1, keepdims=True)
    h = 1e-4
    fd = (bf.nlml(par + h * v) - bf.nlml(par - h * v)) / (2 * h)
    an = np.sum(g * v, axis=1)
    assert np.max(np.abs(fd - an) / np.maximum(1.0, np.abs(an))) <= 1e-6
    one = _make_fit(X[3], y[3], ev[3], "ARD", "Zero", "noiseCorrVec", lnc[3])
    f1, g1 = one.nlml_grad(par[3])
    assert abs(f1 - f[3]) <= 1e-12 * abs(f1) and relmax(g[3], g1) <= 1e-10
    # small enough for the oracle too (n = 400): the batch entry against the reference restatement
    kw = dict(cov_form="ARD", noise_corr="noiseCorrVec", log_hyp_noise_corr=lnc[3])
    fo = orc.nlml(par[3], X[3], y[3], ev[3], as_written_ops=False, **kw)
    assert abs(f[3] - fo) <= TOL_NLML * abs(fo)
    assert relmax(g[3], orc.nlml_grad(par[3], X[3], y[3], ev[3], **kw)) <= TOL_GRAD
    one.close()
    bf.close()


def test_full_size_c5_properties():
    """C5 shape (n=16384, d=50, SqExp, GPS1): too large for the oracle, so size-independent
    properties — directional derivative, Ky alpha = y on a row sample with the library's own
    covariance, prediction against K*' alpha, variance bounds."""
    n, d, m = 16384, 50, 1024
    rng = np.random.default_rng(4)
    X = np.asfortranarray(rng.standard_normal((n, d)))
    Xs = np.asfortranarray(rng.standard_normal((m, d)))
    y = rng.standard_normal(n)
    ev = (rng.random(n) > 0.5).astype(np.float64)
    par = np.log([0.2, 0.8, 1.7 * np.sqrt(d)])
    fit = gp.Fit(0)
    fit.set_options("SqExp", "Zero", False)
    fit.set_data(X, y, ev)
    f, g = fit.nlml_grad(par)
    v = np.array([0.6, -0.48, 0.64])
    h = 1e-4
    fd = (fit.nlml(par + h * v) - fit.nlml(par - h * v)) / (2 * h)
    assert abs(fd - g @ v) <= 1e-6 * abs(fd)
    fin = fit.finalize(par)
    assert abs(-fin["logMarLik"] - f) <= 1e-12 * abs(f)
    alpha = fin["alpha"]
    idx = rng.choice(n, 48, replace=False)
    Krows = fit.cov(X[idx], X, 0.8, [1.7 * np.sqrt(d)], "SqExp")
    assert np.max(np.abs(Krows @ alpha + 0.2 * alpha[idx] - y[idx])) <= 1e-9
    mu, vf, vd = fit.predict(par, Xs)
    Ks = fit.cov(X, Xs[:24], 0.8, [1.7 * np.sqrt(d)], "SqExp")
    assert np.max(np.abs(mu[:24] - Ks.T @ alpha)) <= TOL_PRED * max(1.0, np.max(np.abs(mu)))
    assert np.all(vf >= -1e-9) and np.all(vf <= 0.8 + 1e-9) and np.allclose(vd - vf, 0.2, atol=1e-12)
    fit.close()


# ---- randomized sweep: shapes around every tile / block / padding edge, all option combinations -----------
def _random_case(seed):
    rng = np.random.default_rng(seed)
    n = int(rng.choice([1, 2, 3, 5, 31, 32, 33, 63, 64, 65, 66, 95, 127, 128, 129, 130, 191, 193, 255, 257, 300]))
    d = int(rng.choice([1, 2, 3, 7, 15, 16, 17, 33, 70]))
    m = int(rng.choice([1, 2, 7, 64, 65, 130]))
    cov_form = str(rng.choice(["SqExp", "ARD", "Matern"]))
    mean_form = str(rng.choice(["Zero", "Linear"])) if n > d + 3 else "Zero"
    noise_corr = [False, "noiseCorrLearned", "noiseCorrVec"][int(rng.integers(3))]
    extra = int(rng.choice([1, 3, 5])) if cov_form == "Matern" else None
    c = int(rng.integers(0, n)) if n > 1 else 0
    return n, m, d, c, cov_form, mean_form, noise_corr, extra


@pytest.mark.parametrize("seed", range(36))
def test_randomized_shapes_and_options(seed):
    n, m, d, c, cov_form, mean_form, noise_corr, extra = _random_case(1000 + seed)
    rng = np.random.default_rng(seed)
    X = np.asfortranarray(rng.standard_normal((n, d)))
    Xs = np.asfortranarray(rng.standard_normal((m, d)))
    y = rng.standard_normal(n)
    ev = np.ones(n)
    ev[rng.permutation(n)[:c]] = 0
    p = d if cov_form == "ARD" else 1
    par = [np.log(0.3), np.log(0.8)] + list(np.log(1.5 * np.sqrt(d)) + 0.2 * rng.standard_normal(p))
    if mean_form == "Linear":
        par += list(0.1 * rng.standard_normal(d + 1))
    lnc = None
    if noise_corr == "noiseCorrLearned":
        par += [np.log(0.3)]
    if noise_corr == "noiseCorrVec":
        lnc = np.log(0.2) + 0.5 * rng.standard_normal(c)
    par = np.array(par)
    kw = dict(cov_form=cov_form, mean_form=mean_form, noise_corr=noise_corr, log_hyp_noise_corr=lnc)
    ex = {"extra_param": extra} if extra is not None else {}
    fit = _make_fit(X, y, ev, cov_form, mean_form, noise_corr, lnc, extra_param=extra)
    fo = orc.nlml(par, X, y, ev, **kw, **ex)
    assert abs(fit.nlml(par) - fo) <= TOL_NLML * max(1.0, abs(fo))
    f, g = fit.nlml_grad(par)
    go = orc.nlml_grad(par, X, y, ev, **kw, **ex)
    assert np.max(np.abs(g - go)) <= TOL_GRAD * max(1.0, float(np.max(np.abs(go))))
    fin = fit.finalize(par, want_K=True, want_L=True)
    reg = orc.regress_finalize(par, X, y, ev, cov_form, mean_form, noise_corr, lnc, **ex)
    assert np.max(np.abs(fin["K"] - reg["K"])) <= 1e-11 and relmax(fin["L"], reg["L"]) <= 1e-9
    mu, vf, vd = fit.predict(par, Xs)
    reg["logHypChosen"] = {"noise": par[0], "func": par[1], "length": par[2:2 + p],
                           "mean": par[2 + p:2 + p + d + 1] if mean_form == "Linear" else np.zeros(d + 1)}
    pr = orc.predict(reg, X, y, ev, Xs, cov_form, mean_form, noise_corr,
                     par[-1:] if noise_corr == "noiseCorrLearned" else lnc, **ex)
    scale = max(1.0, float(np.max(np.abs(pr["funcMeanPred"]))))
    assert np.max(np.abs(mu - pr["funcMeanPred"])) <= TOL_PRED * scale
    assert np.max(np.abs(vd - pr["varData"])) <= TOL_PRED
    fit.close()


def test_handle_lifecycle_and_interleaving():
    """One handle across shape changes; two live handles used alternately (separate streams, graphs, caches)."""
    rng = np.random.default_rng(9)
    fa, fb = gp.Fit(0), gp.Fit(0)
    data = {}
    for tag, fit, n, d in (("a", fa, 90, 3), ("b", fb, 140, 5)):
        X, y, ev = rng.standard_normal((n, d)), rng.standard_normal(n), (rng.random(n) > 0.4).astype(float)
        fit.set_options("ARD", "Zero", "noiseCorrLearned")
        fit.set_data(X, y, ev)
        data[tag] = (X, y, ev, np.concatenate([[np.log(0.3), np.log(0.8)], np.full(d, np.log(2.0)), [np.log(0.2)]]))
    kw = dict(cov_form="ARD", noise_corr="noiseCorrLearned")
    for rep in range(4):                      # interleave: a.nlml, b.grad, a.grad (cached), b.nlml ...
        for tag, fit in (("a", fa), ("b", fb)):
            X, y, ev, par = data[tag]
            par = par + 0.01 * rep
            assert abs(fit.nlml(par) - orc.nlml(par, X, y, ev, **kw)) <= 1e-9 * 300
            other = fb if fit is fa else fa
            Xo, yo, evo, paro = data["b" if tag == "a" else "a"]
            _, go = other.nlml_grad(paro)
            assert relmax(go, orc.nlml_grad(paro, Xo, yo, evo, **kw)) <= TOL_GRAD
            _, g = fit.nlml_grad(par)         # must reuse this handle's own cached factorisation
            assert relmax(g, orc.nlml_grad(par, X, y, ev, **kw)) <= TOL_GRAD
    # shape change on a live handle: reallocation, graphs rebuilt, old model dropped
    X2, y2, ev2 = rng.standard_normal((33, 2)), rng.standard_normal(33), np.ones(33)
    fa.set_options("SqExp", "Zero", False)
    fa.set_data(X2, y2, ev2)
    par2 = np.log([0.3, 0.8, 1.5])
    assert abs(fa.nlml(par2) - orc.nlml(par2, X2, y2, ev2)) <= 1e-9 * 100
    with pytest.raises(gp.GpsError):
        fa.predict(par2, X2[:3], reuse_model=True)     # no resident model after set_data
    mu, _, _ = fa.predict(par2, X2[:3])
    assert np.all(np.isfinite(mu))
    fa.close()
    fb.close()
    fa.close()                                          # double close is harmless


# ---- C-side whole fits (gps_fit_batch) against the Python lock-step driver and the sequential mirror ----------
@pytest.mark.parametrize("model,nc,method,cov", [("GPS1", False, "Nelder-Mead", "SqExp"), ("GPS2", "noiseCorrLearned", "Nelder-Mead", "SqExp"),
                                                 ("GPS3", "noiseCorrVec", "Nelder-Mead", "ARD"), ("GPS3", "noiseCorrVec", "BFGS", "ARD"),
                                                 ("GPS2", "noiseCorrLearned", "BFGS", "ARD")])
def test_c_side_fit_batch_matches_python_driver(model, nc, method, cov):
    from gpsurvival_b200 import batchfit
    B, n, d = 4, 70, 2
    pbs = [synth.make_problem(n, 15, d, 28, 900 + b) for b in range(B)]
    rng = np.random.default_rng(3)
    starts = [synth.start_log_hyp(d, cov, rng) for _ in range(B)]
    maxit = 10 if method == "Nelder-Mead" else 4
    params = dict(modelType="survival", noiseCorr=nc, optimType=method, maxit=100, maxitSurvival=maxit, meanFuncForm="Zero",
                  covFuncForm=cov, tolerance=0.02, maxCount=9, logHypStart=starts, burnIn=False, imposePriors=False)
    ref = batchfit.apply_gp_batch(pbs, params)
    X = np.stack([p["trainingData"] for p in pbs])
    Xt = np.stack([p["testData"] for p in pbs])
    y = np.log(np.stack([p["trainingTargets"] for p in pbs]))
    ev = np.stack([p["events"] for p in pbs])
    par0 = np.stack([gp._flatten(s, d, "Zero", nc, s["noise"] if nc == "noiseCorrLearned" else None) for s in starts])
    fit = gp.Fit(0, batch=B)
    fit.set_options(cov, "Zero", nc)
    out = fit.fit_batch(X, y, ev, par0, Xt, method=method, maxit=maxit, tolerance=0.02, max_count=9, survival=True)
    for b in range(B):
        assert out["count"][b] == ref[b]["count"]
        assert abs(out["objective"][b] - ref[b]["objective"]) <= 1e-7 * max(1.0, abs(ref[b]["objective"]))
        assert relmax(out["funcMeanPred"][b], ref[b]["funcMeanPred"].ravel()) <= 1e-6
        assert relmax(out["varData"][b], ref[b]["varData"]) <= 1e-6
        assert relmax(out["trainingSetPredictions"][b], ref[b]["trainingSetPredictions"].ravel()) <= 1e-6
        assert relmax(out["trainingTargetsLearned"][b], ref[b]["trainingTargetsLearned"]) <= 1e-6
    assert out["ticks"] == ref[0]["ticks"]
    fit.close()


def test_c_side_fit_batch_non_survival_and_errors():
    pb = synth.make_problem(60, 10, 1, 0, 77)
    X, y, ev = pb["trainingData"], np.log(pb["trainingTargets"]), pb["events"]
    fit = gp.Fit(0)
    fit.set_options("SqExp", "Zero", False)
    par0 = np.log([0.2, 0.8, 1.7])
    out = fit.fit_batch(X, y, ev, par0, pb["testData"], method="Nelder-Mead", maxit=200, survival=False)
    params = dict(modelType="non-survival", noiseCorr=False, optimType="Nelder-Mead", maxit=200, maxitSurvival=10, meanFuncForm="Zero",
                  covFuncForm="SqExp", tolerance=1e-3, maxCount=10, logHypStart=synth.start_log_hyp(1, "SqExp"))
    ro = orc.apply_gp(pb, params)
    assert abs(out["objective"][0] - ro["objective"]) <= 1e-7 * abs(ro["objective"])
    assert relmax(out["funcMeanPred"][0], ro["funcMeanPred"]) <= 1e-6
    # Matern + BFGS: the derived gradient drives vmmin like any other (same optimum as the oracle's BFGS run)
    fit.set_options("Matern", "Zero", False, "intended", 3)
    outm = fit.fit_batch(X, y, ev, par0, pb["testData"], method="BFGS", maxit=100, survival=False)
    pm = dict(params, covFuncForm="Matern", extraParam=3, optimType="BFGS", maxit=100)
    rm = orc.apply_gp(pb, pm)
    assert abs(outm["objective"][0] - rm["objective"]) <= 1e-6 * abs(rm["objective"])
    with pytest.raises(gp.GpsError):
        fit.fit_batch(X, y, ev, np.zeros(5), pb["testData"], survival=False)              # wrong par length
    fit.close()


def test_chol_matvec_and_gpu_sampler():
    """MakeSyntheticData.R:41,148: t(chol(K)) %*% d1 on the device against numpy."""
    rng = np.random.default_rng(8)
    n, d = 333, 6
    X = rng.standard_normal((n, d))
    par = np.log([1e-6, 0.5, 1.1 * np.sqrt(d)])
    fit = gp.Fit(0)
    fit.set_options("SqExp", "Zero", False)
    fit.set_data(X, np.zeros(n), None)
    v = rng.standard_normal(n)
    out = fit.chol_matvec(par, v)
    K = orc.cov_func(X, X, 0.5, [1.1 * np.sqrt(d)]) + 1e-6 * np.eye(n)
    ref = np.linalg.cholesky(K) @ v
    assert relmax(out, ref) <= 1e-9
    fit.close()
    a = synth.make_problem(200, 40, 3, 80, 5)
    b = synth.make_problem(200, 40, 3, 80, 5, sampler=synth.gpu_sampler(0))
    assert np.array_equal(a["trainingData"], b["trainingData"])
    assert relmax(np.log(b["trainingTargetsPreCensoring"]), np.log(a["trainingTargetsPreCensoring"])) <= 1e-5


# ---- kernel routes: every route of the GEMM dispatcher gives the same numbers -------------------------
def _batched_eval(n, d, B, seed, env):
    """NLML + gradient of B random ARD / GPS3 fits with the library knobs in `env` (read at gps_create)."""
    old = {k: os.environ.get(k) for k in env}
    os.environ.update(env)
    try:
        rng = np.random.default_rng(seed)
        X = np.stack([np.asfortranarray(rng.standard_normal((n, d))) for _ in range(B)])
        y = rng.standard_normal((B, n))
        c = n // 2                                     # a batch shares the censored count (gps_set_noise_corr)
        ev = np.ones((B, n))
        for b in range(B):
            ev[b, rng.choice(n, c, replace=False)] = 0.0
        fit = gp.Fit(0, batch=B)
        fit.set_options("ARD", "Zero", "noiseCorrVec", "intended")
        fit.set_data(X, y, ev)
        fit.set_noise_corr(np.full((B, c), np.log(0.3)))
        par = np.concatenate([[np.log(0.2), np.log(0.8)], np.full(d, np.log(1.7 * np.sqrt(d)))])
        par = par[None, :] + 0.05 * rng.standard_normal((B, par.size))
        f, g = fit.nlml_grad(par)
        f0 = fit.nlml(par + 0.01)
        fit.close()
        return X, y, ev, par, f, g, f0
    finally:
        for k, v in old.items():
            if v is None:
                os.environ.pop(k, None)
            else:
                os.environ[k] = v


@pytest.mark.parametrize("n,d,B", [(285, 1500, 12), (400, 700, 20), (777, 130, 24), (1100, 40, 16)])
def test_dispatcher_routes_agree(n, d, B):
    """Warp-specialised persistent kernels (64 / 80 / 96 / 96x64 / 128x32 tiles, weighted Gram, dynamic tile
    scheduler, stream groups) against the plain cp.async kernel on the pre-scale route: the k-order inside a
    tile is the same, so NLML and gradient agree to rounding of the differently grouped partial sums, and both
    match the oracle."""
    ref = _batched_eval(n, d, B, 5, {"GPS_WS": "0", "GPS_GROUPS": "1"})
    for env in ({"GPS_WS": "1", "GPS_GROUPS": "1"}, {"GPS_WS": "1", "GPS_GROUPS": "2"}, {"GPS_WS": "1", "GPS_GROUPS": "3"},
                {"GPS_WS": "1", "GPS_WS_MIN_N": "256", "GPS_WS_MIN": "8"}):
        out = _batched_eval(n, d, B, 5, env)
        assert relmax(out[4], ref[4]) <= 1e-12, env
        assert relmax(out[5], ref[5]) <= 1e-10, env
        assert relmax(out[6], ref[6]) <= 1e-12, env
    X, y, ev, par, f, g, _ = ref
    for b in (0, B - 1):
        lnc = np.full(int(np.sum(ev[b] == 0)), np.log(0.3))
        kw = dict(cov_form="ARD", noise_corr="noiseCorrVec", log_hyp_noise_corr=lnc)
        fo = orc.nlml(par[b], X[b], y[b], ev[b], **kw)
        go = orc.nlml_grad(par[b], X[b], y[b], ev[b], **kw)
        assert abs(f[b] - fo) <= TOL_NLML * abs(fo)
        assert relmax(g[b], go) <= TOL_GRAD


def test_stream_groups_are_bitwise_neutral():
    """A fit's result must not depend on which stream group it lands in (40 fits in 1 or 4 groups: every launch
    stays on the batched side of the dispatcher's thresholds, so the routes are the same)."""
    a = _batched_eval(520, 60, 40, 9, {"GPS_GROUPS": "1"})
    b = _batched_eval(520, 60, 40, 9, {"GPS_GROUPS": "4"})
    assert np.array_equal(a[4], b[4]) and np.array_equal(a[5], b[5])
    c = _batched_eval(520, 60, 40, 9, {"GPS_GROUPS": "4", "GPS_NBIG": "64"})      # one-level Cholesky: rounding only
    assert relmax(c[4], a[4]) <= 1e-12 and relmax(c[5], a[5]) <= 1e-10


def test_fused_front_end_agrees_with_gram_route():
    """Few features: small_front_kernel (direct differences, one launch) against the DMMA Gram route
    (pre-scale + Gram trick + fused covariance epilogue + fill_aug) — rounding-level agreement, both vs the oracle."""
    a = _batched_eval(300, 7, 6, 3, {"GPS_FRONT": "1"})
    b = _batched_eval(300, 7, 6, 3, {"GPS_FRONT": "0"})
    assert relmax(a[4], b[4]) <= 1e-12 and relmax(a[5], b[5]) <= 1e-10 and relmax(a[6], b[6]) <= 1e-12
    X, y, ev, par, f, g, _ = a
    lnc = np.full(int(np.sum(ev[0] == 0)), np.log(0.3))
    kw = dict(cov_form="ARD", noise_corr="noiseCorrVec", log_hyp_noise_corr=lnc)
    fo = orc.nlml(par[0], X[0], y[0], ev[0], **kw)
    assert abs(f[0] - fo) <= TOL_NLML * abs(fo)
    assert relmax(g[0], orc.nlml_grad(par[0], X[0], y[0], ev[0], **kw)) <= TOL_GRAD


def test_bench_line_carries_the_contract_keys():
    """python bench.py (small batch, few steps): one JSON line on stdout with every key of the measurement contract."""
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    res = subprocess.run([sys.executable, os.path.join(root, "bench.py"), "--steps", "3", "--warmup", "3", "--batch", "4", "--no-extras"],
                         capture_output=True, text=True, timeout=900)
    assert res.returncode == 0, res.stderr[-2000:]
    lines = [ln for ln in res.stdout.strip().splitlines() if ln.startswith("{")]
    assert len(lines) == 1
    b = json.loads(lines[0])
    for k in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline",
              "dtype", "data", "config", "e2e", "gpu_launches", "clocks", "roofline", "cpu_baseline"):
        assert k in b, k
    assert b["n_gpus"] == 1 and b["steps"] == 3 and b["dtype"] == "f64" and b["scaling"] == "weak" and b["vs_baseline"] is None
    assert b["gpu_launches"] > 0 and b["value"] > 0 and b["e2e"]["value"] > 0
    assert b["e2e"]["h2d_bytes_per_step"] > 0 and b["e2e"]["d2h_bytes_per_step"] > 0
    for k in ("bound", "achieved", "peak", "unit", "frac", "traffic"):
        assert k in b["roofline"], k
    for k in ("value", "unit", "cores", "kind", "sample"):
        assert k in b["cpu_baseline"], k
    assert b["parity_vs_oracle"]["nlml_rel"] <= TOL_NLML and b["parity_vs_oracle"]["grad_rel"] <= TOL_GRAD
