"""Developer check run on the GPU box: compares every piece of the CUDA path with the oracle
and prints one line per check (does not stop at the first failure, unlike pytest -x)."""
from __future__ import annotations

import os
import sys
import time
import traceback

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

from gpsurvival_b200 import gp, synth   # noqa: E402
from oracle import gp_oracle as orc     # noqa: E402


def rel(a, b):
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    return float(np.max(np.abs(a - b)) / max(np.max(np.abs(b)), 1e-300))


def report(name, err, tol):
    print(f"{'OK  ' if err <= tol else 'FAIL'} {name:58s} err={err:.3e} tol={tol:.0e}", flush=True)


def problem(n, m, d, c, seed):
    pb = synth.make_problem(n, m, d, c, seed)
    return pb["trainingData"], np.log(pb["trainingTargets"]), pb["events"], pb["testData"]


def check_cov(fit):
    rng = np.random.default_rng(1)
    for (n1, n2, d) in [(100, 0, 1), (301, 0, 20), (285, 0, 3000), (130, 77, 5), (64, 200, 700)]:
        x1 = rng.standard_normal((n1, d))
        x2 = None if n2 == 0 else rng.standard_normal((n2, d))
        for form, ell in (("SqExp", np.array([1.7 * np.sqrt(d)])), ("ARD", 1.7 * np.sqrt(d) * (1 + 0.3 * rng.random(d)))):
            for mode in ("intended", "as_written"):
                K = fit.cov(x1, x2, 0.8, ell, form, mode)
                Ko = orc.cov_func(x1, x1 if x2 is None else x2, 0.8, ell, form, mode)
                report(f"cov {form} {mode} n1={n1} n2={n2} d={d}", float(np.max(np.abs(K - Ko))), 1e-12)


def check_eval(name, n, m, d, c, cov_form, mean_form, noise_corr, seed):
    X, y, ev, Xs = problem(n, m, d, c, seed)
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
    fit = gp.Fit(0)
    fit.set_options(cov_form, mean_form, noise_corr, "intended")
    fit.set_data(X, y, ev)
    if lnc is not None:
        fit.set_noise_corr(lnc)
    kw = dict(cov_form=cov_form, mean_form=mean_form, noise_corr=noise_corr, log_hyp_noise_corr=lnc)
    f_o = orc.nlml(par, X, y, ev, **kw)
    g_o = orc.nlml_grad(par, X, y, ev, **kw)
    for rep in range(3):     # direct, graph capture, graph replay
        f = fit.nlml(par)
        report(f"{name} nlml (call {rep})", abs(f - f_o) / abs(f_o), 1e-9)
    for rep in range(3):
        f2, g = fit.nlml_grad(par)
        report(f"{name} grad (call {rep})", float(np.max(np.abs(g - g_o)) / np.max(np.abs(g_o))), 1e-9)
    par2 = par + 0.01
    f2, g2 = fit.nlml_grad(par2)     # uncached path
    report(f"{name} grad uncached", rel(g2, orc.nlml_grad(par2, X, y, ev, **kw)), 1e-9)
    fin = fit.finalize(par, want_K=True, want_L=True)
    reg = orc.regress_finalize(par, X, y, ev, cov_form, mean_form, noise_corr, lnc)
    report(f"{name} finalize K", float(np.max(np.abs(fin["K"] - reg["K"]))), 1e-12)
    report(f"{name} finalize L", rel(fin["L"], reg["L"]), 1e-9)
    report(f"{name} finalize alpha", rel(fin["alpha"], reg["alpha"]), 1e-8)
    report(f"{name} finalize logMarLik", abs(fin["logMarLik"] - reg["logMarLik"]) / abs(reg["logMarLik"]), 1e-9)
    mu, vf, vd = fit.predict(par, Xs)
    reg["logHypChosen"] = {"noise": par[0], "func": par[1], "length": par[2:2 + p],
                           "mean": par[2 + p:2 + p + d + 1] if mean_form == "Linear" else np.zeros(d + 1)}
    pr = orc.predict(reg, X, y, ev, Xs, cov_form, mean_form, noise_corr, lnc if noise_corr != "noiseCorrLearned" else par[-1:])
    report(f"{name} predict mean", rel(mu, pr["funcMeanPred"]), 1e-8)
    report(f"{name} predict varFunction", float(np.max(np.abs(vf - pr["varFunction"])) / 0.8), 1e-8)
    report(f"{name} predict varData", float(np.max(np.abs(vd - pr["varData"])) / 0.8), 1e-8)
    fit.close()


def check_adjust(fit):
    rng = np.random.default_rng(3)
    c = 5000
    mu = rng.standard_normal(c)
    s = 0.05 + rng.random(c)
    alpha = rng.uniform(-8, 7.5, c)
    a = mu + alpha * s
    em, ev = fit.adjust(a, mu, s)
    eo, vo = orc.adjust(a, mu, s)
    al = (a - mu) / s
    for lo, hi, tol in ((-9, 5, 1e-8), (5, 6.5, 1e-5), (6.5, 9, 1.0)):
        msk = (al >= lo) & (al < hi)
        report(f"adjust mean alpha in [{lo},{hi})", float(np.max(np.abs(em[msk] - eo[msk]) / np.maximum(np.abs(eo[msk]), 1))), tol)
        report(f"adjust var  alpha in [{lo},{hi})", float(np.max(np.abs(ev[msk] - vo[msk]) / np.maximum(np.abs(vo[msk]), 1e-3))), tol)
    report("adjust branch (a,0) agreement", float(np.mean((ev == 0) != (vo == 0))), 0.002)


def check_apply_gp():
    cfg = synth.CONFIGS["C1"]
    pb = synth.make_config_problem(cfg)
    for model, nc in (("GPS1", False), ("GPS2", "noiseCorrLearned"), ("GPS3", "noiseCorrVec"), ("GPR", None)):
        params = dict(modelType="survival" if nc is not None else "non-survival", noiseCorr=nc if nc is not None else False,
                      optimType="Nelder-Mead", maxit=300, maxitSurvival=10, meanFuncForm="Zero", covFuncForm="SqExp",
                      tolerance=1e-5 * 100 * 0.4, maxCount=20, logHypStart=synth.start_log_hyp(1, "SqExp"),
                      burnIn=False, imposePriors=False)
        tts = pb
        if nc is None:
            keep = pb["events"] == 1
            tts = dict(pb, trainingData=pb["trainingData"][keep], trainingTargets=pb["trainingTargets"][keep],
                       events=pb["events"][keep])
        t = time.time()
        r = gp.ApplyGP(tts, {"dataSource": "Generate"}, params)
        tg = time.time() - t
        t = time.time()
        ro = orc.apply_gp(tts, params)
        to = time.time() - t
        report(f"ApplyGP {model} funcMeanPred (gpu {tg:.2f}s, oracle {to:.2f}s)", rel(r["funcMeanPred"].ravel(), ro["funcMeanPred"]), 1e-6)
        report(f"ApplyGP {model} objective", abs(r["objective"] - ro["objective"]) / abs(ro["objective"]), 1e-6)
        if "count" in r:
            report(f"ApplyGP {model} rounds {r['count']} vs {ro['count']}", abs(r["count"] - ro["count"]), 0)


def main():
    fit = gp.Fit(0)
    for fn in (lambda: check_cov(fit), lambda: check_adjust(fit)):
        try:
            fn()
        except Exception:
            traceback.print_exc()
    cases = [
        ("C1 SqExp GPS1 n=100 d=1", 100, 50, 1, 40, "SqExp", "Zero", False, 11),
        ("SqExp GPS2 n=257 d=4", 257, 40, 4, 100, "SqExp", "Zero", "noiseCorrLearned", 12),
        ("ARD GPS3 n=300 d=20", 300, 64, 20, 150, "ARD", "Zero", "noiseCorrVec", 13),
        ("ARD Linear GPS2 n=201 d=6", 201, 33, 6, 90, "ARD", "Linear", "noiseCorrLearned", 14),
        ("SqExp Linear GPS1 n=150 d=3", 150, 20, 3, 60, "SqExp", "Linear", False, 15),
        ("ARD GPS3 n=285 d=1500", 285, 57, 1500, 142, "ARD", "Zero", "noiseCorrVec", 16),
        ("ARD GPS3 n=700 d=20", 700, 120, 20, 350, "ARD", "Zero", "noiseCorrVec", 17),
    ]
    for c in cases:
        try:
            check_eval(*c)
        except Exception:
            traceback.print_exc()
    try:
        check_apply_gp()
    except Exception:
        traceback.print_exc()


if __name__ == "__main__":
    main()
