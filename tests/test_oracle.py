"""CPU tests: pin the numpy oracle (oracle/gp_oracle.py).

The reference ships no tests or golden vectors and R is absent, so the pins
are: the committed mpmath fixtures, finite differences, closed forms, algebraic invariants and
the four scratch-probe facts of SURVEY.md §0 (F3, F4, the Adjust cancellation, R's documented
optim example)."""
import json
import math
import os

import numpy as np
import pytest

from oracle import gp_oracle as orc
from oracle import mp_oracle as mpo
from gpsurvival_b200 import synth

GOLD = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "golden_small.json")))


def _case_kw(c):
    return dict(cov_form=c["cov_form"], mean_form=c["mean_form"], noise_corr=c["noise_corr"],
                log_hyp_noise_corr=c["log_sc2_vec"])


def _extra(c):
    return {"extra_param": c.get("extra_param")} if c.get("extra_param") is not None else {}


@pytest.mark.parametrize("c", GOLD["cases"], ids=[c["name"] for c in GOLD["cases"]])
def test_oracle_matches_mpmath_golden(c):
    X, y, ev = np.array(c["X"]), np.array(c["y"]), np.array(c["events"])
    par = np.array(c["par"])
    kw = _case_kw(c)
    for as_written in (True, False):
        f = orc.nlml(par, X, y, ev, as_written_ops=as_written, **kw, **_extra(c))
        assert abs(f - c["nlml"]) <= 1e-11 * abs(c["nlml"])
    if c["grad"] is not None:
        g = orc.nlml_grad(par, X, y, ev, **kw, **_extra(c))
        assert np.max(np.abs(g - np.array(c["grad"]))) <= 1e-10 * np.max(np.abs(c["grad"]))
    reg = orc.regress_finalize(par, X, y, ev, **kw, **_extra(c))
    assert abs(reg["logMarLik"] + c["nlml"]) <= 1e-11 * abs(c["nlml"])
    d = X.shape[1]
    p = d if c["cov_form"] == "ARD" else 1
    reg["logHypChosen"] = {"noise": par[0], "func": par[1], "length": par[2:2 + p],
                           "mean": par[2 + p:2 + p + d + 1] if c["mean_form"] == "Linear" else np.zeros(d + 1)}
    lnc = par[-1:] if c["noise_corr"] == "noiseCorrLearned" else c["log_sc2_vec"]
    for as_written in (True, False):
        pr = orc.predict(reg, X, y, ev, np.array(c["Xstar"]), c["cov_form"], c["mean_form"], c["noise_corr"], lnc,
                         as_written_ops=as_written, **_extra(c))
        assert np.allclose(pr["funcMeanPred"], c["pred_mean"], rtol=1e-10, atol=1e-12)
        assert np.allclose(pr["varFunction"], c["pred_varf"], rtol=0, atol=1e-11)
        assert np.allclose(pr["varData"], c["pred_vard"], rtol=0, atol=1e-11)


GOLD_MID = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "golden_mid.json")))


@pytest.mark.parametrize("c", GOLD_MID["cases"], ids=[c["name"] for c in GOLD_MID["cases"]])
def test_oracle_matches_mpmath_golden_mid(c):
    """n = 65 ... 600: the sizes that cross the CUDA Cholesky's leaf / leaf pair / 256-wide block (golden_mid.json)."""
    X, y, ev = np.array(c["X"]), np.array(c["y"]), np.array(c["events"])
    par = np.array(c["par"])
    kw = _case_kw(c)
    f = orc.nlml(par, X, y, ev, **kw)
    assert abs(f - c["nlml"]) <= 1e-11 * abs(c["nlml"])
    g = orc.nlml_grad(par, X, y, ev, **kw)
    assert np.max(np.abs(g - np.array(c["grad"]))) <= 1e-10 * np.max(np.abs(c["grad"]))
    reg = orc.regress_finalize(par, X, y, ev, **kw)
    d = X.shape[1]
    p = d if c["cov_form"] == "ARD" else 1
    reg["logHypChosen"] = {"noise": par[0], "func": par[1], "length": par[2:2 + p], "mean": np.zeros(d + 1)}
    pr = orc.predict(reg, X, y, ev, np.array(c["Xstar"]), c["cov_form"], "Zero", c["noise_corr"], c["log_sc2_vec"])
    assert np.allclose(pr["funcMeanPred"], c["pred_mean"], rtol=1e-9, atol=1e-11)
    assert np.allclose(pr["varFunction"], c["pred_varf"], rtol=0, atol=1e-10)


def test_matern_closed_forms():
    """Matern 1/2 is the exponential kernel; all three equal sf2 at r = 0 and decrease in r."""
    x = np.array([[0.0], [0.7], [2.0]])
    for nu in (1, 3, 5):
        K = orc.cov_func(x, x, 0.9, [1.3], "Matern", extra_param=nu)
        assert np.allclose(np.diag(K), 0.9) and K[0, 1] > K[0, 2] > 0
    K1 = orc.cov_func(x, x, 0.9, [1.3], "Matern", extra_param=1)
    assert abs(K1[0, 1] - 0.9 * math.exp(-0.7 / 1.3)) < 1e-15
    with pytest.raises(ValueError):
        orc.cov_func(x, x, 0.9, [1.3], "Matern", extra_param=2)


def test_golden_fixture_is_reproducible():
    """The committed fixture is what make_golden.py produces (first case recomputed)."""
    c = GOLD["cases"][0]
    f = float(mpo.nlml(c["par"], c["X"], c["y"], c["events"]))
    assert f == c["nlml"]


def _fd(f, p, h=1e-5):
    g = np.zeros_like(p)
    for i in range(p.size):
        e = np.zeros_like(p)
        e[i] = h
        g[i] = (f(p + e) - f(p - e)) / (2 * h)
    return g


@pytest.mark.parametrize("cov_form,mean_form,noise_corr", [
    ("SqExp", "Zero", False), ("SqExp", "Zero", "noiseCorrLearned"), ("ARD", "Zero", "noiseCorrVec"),
    ("ARD", "Zero", "noiseCorrLearned")])
def test_gradient_matches_finite_differences(cov_form, mean_form, noise_corr):
    pb = synth.make_problem(40, 5, 3, 15, 42)
    X, y, ev = pb["trainingData"], np.log(pb["trainingTargets"]), pb["events"]
    p = 3 if cov_form == "ARD" else 1
    par = np.concatenate([[math.log(0.2), math.log(0.8)], np.log(1.7 * math.sqrt(3)) + 0.1 * np.arange(p)])
    lnc = None
    if noise_corr == "noiseCorrLearned":
        par = np.append(par, math.log(0.3))
    if noise_corr == "noiseCorrVec":
        lnc = np.full(int(np.sum(ev == 0)), math.log(0.25))
    kw = dict(cov_form=cov_form, mean_form=mean_form, noise_corr=noise_corr, log_hyp_noise_corr=lnc)
    g = orc.nlml_grad(par, X, y, ev, **kw)
    gfd = _fd(lambda q: orc.nlml(q, X, y, ev, as_written_ops=False, **kw), par)
    assert np.max(np.abs(g - gfd)) <= 2e-7 * max(1.0, np.max(np.abs(g)))


def test_closed_form_two_points():
    """n=2, d=1 NLML by hand."""
    X = np.array([[0.0], [1.0]])
    y = np.array([0.3, -0.2])
    sn2, sf2, ell = 0.1, 0.7, 1.3
    k = sf2 * math.exp(-0.5 / ell ** 2)
    a, b = sf2 + sn2, k
    det = a * a - b * b
    quad = (a * y[0] ** 2 - 2 * b * y[0] * y[1] + a * y[1] ** 2) / det
    expect = 0.5 * quad + 0.5 * math.log(det) + math.log(2 * math.pi)
    got = orc.nlml(np.log([sn2, sf2, ell]), X, y, np.ones(2))
    assert abs(got - expect) < 1e-13


def test_cov_properties_and_F3_recycling():
    rng = np.random.default_rng(0)
    X = rng.standard_normal((60, 7))
    ell = 1.0 + rng.random(7)
    K = orc.cov_func(X, X, 0.5, ell, "ARD", "intended")
    assert np.allclose(K, K.T, atol=0) and np.all(np.diag(K) == 0.5)
    Kw = orc.cov_func(X, X, 0.5, ell, "ARD", "as_written")
    assert np.max(np.abs(K - Kw)) > 1e-2            # F3: as written is not per-dimension scaling
    X1 = X[:, :1]
    assert np.array_equal(orc.cov_func(X1, X1, 0.5, ell[:1], "ARD", "intended"),
                          orc.cov_func(X1, X1, 0.5, ell[:1], "ARD", "as_written"))   # exact at d=1
    # recycling rule: element (i, j) of the n x d matrix is divided by ell[(i + j*n) mod d]
    Z = orc._scale_inputs(X, ell, "as_written")
    i, j = 13, 4
    assert Z[i, j] == X[i, j] / ell[(i + j * 60) % 7]


def test_F4_as_written_ard_gradient_is_not_the_derivative():
    pb = synth.make_problem(30, 4, 3, 10, 7)
    X, y, ev = pb["trainingData"], np.log(pb["trainingTargets"]), pb["events"]
    par = np.log([0.2, 0.8, 2.0, 2.5, 3.0])
    ga = orc.nlml_grad(par, X, y, ev, cov_form="ARD", grad_mode="analytic")
    gw = orc.nlml_grad(par, X, y, ev, cov_form="ARD", grad_mode="as_written")
    gfd = _fd(lambda q: orc.nlml(q, X, y, ev, cov_form="ARD", as_written_ops=False), par)
    assert np.max(np.abs(ga - gfd)) < 1e-6 * max(1, np.max(np.abs(ga)))
    assert np.max(np.abs(gw[2:] - gfd[2:])) > 1e-2


def test_sqexp_reference_branch_equals_W_form():
    """The SqExp branch restated literally (:50-57) equals the W-form the CUDA path uses."""
    pb = synth.make_problem(25, 4, 2, 8, 9)
    X, y, ev = pb["trainingData"], np.log(pb["trainingTargets"]), pb["events"]
    par = np.log([0.2, 0.8, 2.2, 0.3])
    g = orc.nlml_grad(par, X, y, ev, cov_form="SqExp", noise_corr="noiseCorrLearned")
    gm = [float(v) for v in mpo.nlml_grad(par, X.tolist(), y, ev, gps2=True)]
    assert np.allclose(g, gm, rtol=1e-10, atol=1e-12)


def test_factor_invariants():
    pb = synth.make_problem(50, 10, 4, 20, 3)
    X, y, ev = pb["trainingData"], np.log(pb["trainingTargets"]), pb["events"]
    par = np.log([0.2, 0.8, 3.0])
    reg = orc.regress_finalize(par, X, y, ev)
    Ky = reg["K"] + 0.2 * np.eye(50)
    assert np.allclose(reg["L"] @ reg["L"].T, Ky, atol=1e-13)
    assert np.allclose(Ky @ reg["alpha"], y, atol=1e-11)
    reg["logHypChosen"] = {"noise": par[0], "func": par[1], "length": par[2:], "mean": np.zeros(5)}
    pr = orc.predict(reg, X, y, ev, pb["testData"])
    assert np.all(pr["varFunction"] >= -1e-12) and np.all(pr["varFunction"] <= 0.8 + 1e-12)


# ---- Adjust ---------------------------------------------------------------------------------
def test_cody_pnorm_and_dnorm_accuracy():
    import mpmath as mp
    mp.mp.dps = 50
    for x in np.linspace(-37, 8.2, 1500):
        ref = float(mp.erfc(-mp.mpf(float(x)) / mp.sqrt(2)) / 2)
        # exp() of a large argument amplifies its own rounding in the far tail
        assert abs(orc.pnorm_std(float(x)) - ref) <= (1e-15 if abs(x) < 6 else 1e-13) * ref
    for x in np.linspace(-37, 37, 1500):           # beyond |x| ~ 37.5 the density is denormal
        ref = float(mp.npdf(mp.mpf(float(x))))
        assert abs(orc.dnorm_std(float(x)) - ref) <= (1e-15 if abs(x) < 6 else 1e-13) * ref
    assert orc.pnorm_std(0.0) == 0.5 and orc.pnorm_std(9.0) == 1.0 and orc.pnorm_std(-40.0) == 0.0


@pytest.mark.parametrize("row", GOLD["adjust"], ids=[f"alpha={r['alpha']:.2f}" for r in GOLD["adjust"]])
def test_adjust_against_fixture(row):
    em, ev = orc.adjust_one(row["a"], row["mu"], row["s"])
    assert em == row["as_written_mean"] and ev == row["as_written_var"]
    if row["alpha"] < 5:      # where 1 - Phi is well conditioned the lossy form agrees with the exact one
        assert abs(em - row["exact_mean"]) <= 1e-8 * max(1, abs(row["exact_mean"]))
        assert abs(ev - row["exact_var"]) <= 1e-7 * max(1e-3, abs(row["exact_var"]))
    assert em >= row["a"] - 1e-12       # adjusted mean is never below the censoring time


def test_adjust_cancellation_probe():
    """lossy 1 - pnorm vs exact tail; and the hard (a, 0) branch."""
    import mpmath as mp
    for al, lo, hi in ((5.0, 1e-11, 1e-9), (6.0, 1e-9, 1e-6)):
        lossy = 1.0 - orc.pnorm_std(al)
        exact = float(mp.erfc(mp.mpf(al) / mp.sqrt(2)) / 2)
        assert lo < abs(lossy - exact) / exact < hi
    assert orc.adjust_one(7.2, 0.0, 1.0) == (7.2, 0.0)
    em, ev = orc.adjust_one(0.5, 0.0, 1.0)
    import scipy.stats as st
    tn = st.truncnorm(0.5, np.inf)
    assert abs(em - tn.mean()) < 1e-12 and abs(ev - tn.var()) < 1e-12


# ---- optim ----------------------------------------------------------------------------------
def _rosen(x):
    return 100.0 * (x[1] - x[0] ** 2) ** 2 + (1.0 - x[0]) ** 2


def _rosen_g(x):
    return np.array([-400.0 * x[0] * (x[1] - x[0] ** 2) - 2.0 * (1.0 - x[0]), 200.0 * (x[1] - x[0] ** 2)])


def test_optim_reproduces_R_documented_example():
    """?optim's own example, `optim(c(-1.2,1), fr)` and `optim(c(-1.2,1), fr, grr, method="BFGS")`,
    prints value 8.825241e-08 with 195 function evaluations, and value 9.594956e-18 with counts
    110 / 43 — the only published outputs of stats::optim we can pin against without R."""
    from gpsurvival_b200 import optim as popt
    for mod in (orc, popt):
        r = mod.nelder_mead(_rosen, [-1.2, 1.0], maxit=500)
        assert r["counts"] == 195 and abs(r["value"] - 8.825241e-08) < 5e-15
        assert np.allclose(r["par"], [1.000260, 1.000506], atol=5e-7)
        b = mod.bfgs(_rosen, _rosen_g, [-1.2, 1.0], maxit=100)
        assert tuple(b["counts"]) == (110, 43) and abs(b["value"] - 9.594956e-18) < 5e-24


# ---- ApplyGP loop ---------------------------------------------------------------------------
def test_apply_gp_loop_runs_minimum_six_rounds_and_updates_targets():
    cfg = synth.CONFIGS["C1"]
    pb = synth.make_config_problem(cfg, scale=0.4)
    params = dict(modelType="survival", noiseCorr="noiseCorrVec", optimType="Nelder-Mead", maxit=200,
                  maxitSurvival=10, meanFuncForm="Zero", covFuncForm="SqExp", tolerance=1e9, maxCount=50,
                  logHypStart=synth.start_log_hyp(1, "SqExp"))
    r = orc.apply_gp(pb, params)
    assert r["count"] == 6                      # `| count <= 5` (ApplyGP.R:162)
    cens = pb["events"] == 0
    ylog = np.log(pb["trainingTargets"])
    assert np.all(r["trainingTargetsLearned"][cens] >= ylog[cens] - 1e-12)
    assert np.array_equal(r["trainingTargetsLearned"][~cens], ylog[~cens])
    assert r["funcMeanPred"].shape == (pb["testData"].shape[0],)


def test_lbfgs_twin_minimises_and_agrees_with_vmmin():
    """optim.lbfgs (the host twin of csrc/lbfgs.cuh): Rosenbrock to machine precision, and on a GP likelihood the same
    optimum as the vmmin restatement (both stop on vmmin's relative-reduction test)."""
    from gpsurvival_b200 import optim

    def fr(x):
        return 100.0 * (x[1] - x[0] ** 2) ** 2 + (1.0 - x[0]) ** 2

    def gr(x):
        return np.array([-400.0 * x[0] * (x[1] - x[0] ** 2) - 2.0 * (1.0 - x[0]), 200.0 * (x[1] - x[0] ** 2)])
    r = optim.lbfgs(fr, gr, [-1.2, 1.0], maxit=200)
    assert r["convergence"] == 0 and r["value"] < 1e-15 and np.allclose(r["par"], 1.0, atol=1e-7)
    pb = synth.make_problem(80, 10, 3, 30, 5)
    X, y, ev = pb["trainingData"], np.log(pb["trainingTargets"]), pb["events"]
    kw = dict(cov_form="ARD", mean_form="Zero", noise_corr=False)
    p0 = np.concatenate([np.log([0.2, 0.8]), np.log(1.7 * np.sqrt(3)) * np.ones(3)])
    a = optim.lbfgs(lambda p: orc.nlml(p, X, y, ev, **kw), lambda p: orc.nlml_grad(p, X, y, ev, **kw), p0, maxit=200)
    b = optim.bfgs(lambda p: orc.nlml(p, X, y, ev, **kw), lambda p: orc.nlml_grad(p, X, y, ev, **kw), p0, maxit=200)
    assert a["convergence"] == 0 and abs(a["value"] - b["value"]) <= 1e-5 * abs(b["value"])
    assert a["value"] <= b["value"] + 1e-6 * abs(b["value"])
    # maxit is a cap on gradient evaluations, as in vmmin
    c = optim.lbfgs(lambda p: orc.nlml(p, X, y, ev, **kw), lambda p: orc.nlml_grad(p, X, y, ev, **kw), p0, maxit=3)
    assert c["convergence"] == 1 and c["counts"][1] == 3


def test_concordance_index_restatement_closed_forms():
    """c-index semantics (survcomp defaults, restated): perfect risk ordering -> 1, reversed -> 0, censored short times
    make pairs unusable, ties in x are excluded, ties in time are unusable."""
    t = np.array([1.0, 2.0, 3.0, 4.0, 5.0])
    e = np.ones(5)
    assert orc.concordance_index(-t, t, e)[0] == 1.0            # higher risk = shorter time
    assert orc.concordance_index(t, t, e)[0] == 0.0
    c, conc, disc, tied = orc.concordance_index(-t, t, np.array([0.0, 1.0, 1.0, 1.0, 1.0]))
    assert (conc, disc, tied) == (6, 0, 0) and c == 1.0          # the censored shortest time is never the "earlier event"
    c, conc, disc, tied = orc.concordance_index(np.array([3.0, 3.0, 1.0]), np.array([1.0, 2.0, 3.0]), np.ones(3))
    assert (conc, disc, tied) == (2, 0, 1) and c == 1.0
    c, conc, disc, tied = orc.concordance_index(np.array([2.0, 1.0]), np.array([1.0, 1.0]), np.ones(2))
    assert conc + disc + tied == 0 and np.isnan(c)
    assert abs(orc.rmse([1.0, 2.0], [2.0, 4.0]) - np.sqrt(2.5)) < 1e-15


def test_philox_known_answers_and_synth_recipe():
    """The synthetic-data restatement (oracle/synth_oracle.py): Philox4x32-10 against Random123's published known-answer
    vectors, the normal stream's moments, and the censoring loop's contract (CensorData.R:36-48)."""
    from oracle import synth_oracle as so
    assert so.philox4x32_10((0, 0, 0, 0), (0, 0)) == (0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8)
    assert so.philox4x32_10((0xffffffff,) * 4, (0xffffffff,) * 2) == (0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd)
    assert so.philox4x32_10((0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344), (0xa4093822, 0x299f31d0)) == \
        (0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1)
    z = so.normals(20000, 0, 0, 1234)
    assert abs(z.mean()) < 0.03 and abs(z.std() - 1.0) < 0.03
    X, y0, y, ev, done = so.generate(7, 0, 60, 2, 0.01, 0.5, 1.5, 1e-6, 24, 0.0, 50.0)
    assert done == 24 and int(np.sum(ev == 0)) == 24 and np.all(y[ev == 0] < y0[ev == 0]) and np.array_equal(y[ev == 1], y0[ev == 1])


def test_near_pd_restatement_properties_and_mpmath_pin():
    """gp_oracle.near_pd (Matrix::nearPD's published algorithm with the reference's defaults, ForOptimisationLog.R:43-51):
    independent 50-digit restatement on small indefinite / singular matrices, and the properties the algorithm guarantees."""
    from oracle import mp_oracle as mpo
    rng = np.random.default_rng(12)
    cases = []
    A = rng.standard_normal((6, 6)); A = A + A.T                          # indefinite
    cases.append(A)
    B = rng.standard_normal((5, 3)); B = B @ B.T                           # rank-deficient PSD
    cases.append(np.vstack([np.hstack([B, B]), np.hstack([B, B])])[:7, :7])   # duplicated rows, as a singular Ky has them
    C = np.eye(4) + 0.9 * (np.ones((4, 4)) - np.eye(4)); C[0, 1] = C[1, 0] = -0.9   # a broken correlation matrix
    cases.append(C)
    for x in cases:
        got, conv, it = orc.near_pd(x)
        ref, conv_mp, it_mp = mpo.near_pd(x.tolist())
        ref = np.array([[float(v) for v in row] for row in ref])
        assert conv and conv_mp and it == it_mp
        assert np.max(np.abs(got - ref)) <= 1e-9 * np.max(np.abs(ref))
        w = np.linalg.eigvalsh(got)
        assert np.allclose(got, got.T, atol=1e-14 * np.max(np.abs(got)))
        assert w[0] >= 0.99e-8 * w[-1]                                      # posd.tol floor
        np.linalg.cholesky(got)                                             # what ForOptimisationLog.R:45 does next
    # a well-conditioned positive definite matrix is a fixed point (one projection, nothing dropped, nothing raised)
    P = B[:3, :3] + 3.0 * np.eye(3)
    got, conv, it = orc.near_pd(P)
    assert conv and it == 1 and np.max(np.abs(got - P)) <= 1e-13 * np.max(np.abs(P))
