"""Generate tests/golden/golden_small.json.

The reference ships no golden vectors and cannot run here (no R) — SURVEY.md F1/F2 — so the
committed expectations come from the independent 50-digit mpmath restatement
(oracle/mp_oracle.py), rounded to FP64.  Both the numpy oracle (-m "not gpu") and the CUDA
path (-m gpu) are checked against this file.  Re-run:  python tests/golden/make_golden.py
"""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import mp_oracle as mpo          # noqa: E402
from oracle import gp_oracle as orc          # noqa: E402
from gpsurvival_b200 import synth            # noqa: E402

CASES = [
    # name, n, m, d, c, cov, mean, noise_corr
    ("sqexp_zero_gps1", 10, 4, 1, 4, "SqExp", "Zero", False),
    ("sqexp_zero_gps2", 11, 3, 2, 5, "SqExp", "Zero", "noiseCorrLearned"),
    ("ard_zero_gps3", 12, 4, 3, 5, "ARD", "Zero", "noiseCorrVec"),
    ("ard_linear_gps2", 9, 3, 2, 4, "ARD", "Linear", "noiseCorrLearned"),
    ("sqexp_linear_none", 8, 3, 3, 3, "SqExp", "Linear", False),
    ("ard_zero_gps1_d5", 13, 5, 5, 6, "ARD", "Zero", False),
    ("matern1_zero_gps1", 9, 3, 2, 3, "Matern1", "Zero", False),
    ("matern3_zero_gps3", 10, 4, 2, 4, "Matern3", "Zero", "noiseCorrVec"),
    ("matern5_zero_gps2", 11, 3, 3, 5, "Matern5", "Zero", "noiseCorrLearned"),
]


def main():
    out = {"note": "expected values: 50-digit mpmath restatement rounded to FP64 (no R available; parity unpinned "
                   "by the reference itself)", "cases": []}
    for idx, (name, n, m, d, c, cov, mean, nc) in enumerate(CASES):
        pb = synth.make_problem(n, m, d, c, 777 + idx)
        X, y, ev, Xs = pb["trainingData"], np.log(pb["trainingTargets"]), pb["events"], pb["testData"]
        rng = np.random.default_rng(100 + idx)
        p = d if cov == "ARD" else 1
        par = [np.log(0.2), np.log(0.8)] + list(np.log(1.7 * np.sqrt(d)) + 0.3 * rng.standard_normal(p))
        if mean == "Linear":
            par += list(0.2 * rng.standard_normal(d + 1))
        lnc = None
        if nc == "noiseCorrLearned":
            par += [np.log(0.3)]
        if nc == "noiseCorrVec":
            lnc = (np.log(0.2) + 0.5 * rng.standard_normal(int(np.sum(ev == 0)))).tolist()
        kw = dict(ard=(cov == "ARD"), linear_mean=(mean == "Linear"), gps2=(nc == "noiseCorrLearned"),
                  log_sc2_vec=lnc)
        Xl = X.tolist()
        extra = None
        if cov.startswith("Matern"):
            extra = int(cov[-1])
            cov = "Matern"
        mpo.MATERN = extra
        nl = float(mpo.nlml(par, Xl, y, ev, **kw))
        # the reference has no Matern gradient (OptimisationGradientLog.R:6); ours is derived (mp_oracle._cov_d)
        # and checked here against a 50-digit central difference of the mpmath NLML before it is pinned
        grm = mpo.nlml_grad(par, Xl, y, ev, **kw)
        if extra is not None:
            import mpmath as mp
            hstep = mp.mpf(10) ** -20
            for q in range(len(par)):
                pp = [mp.mpf(float(v)) for v in par]
                pm = list(pp)
                pp[q] += hstep
                pm[q] -= hstep
                fd = (mpo.nlml(pp, Xl, y, ev, **kw) - mpo.nlml(pm, Xl, y, ev, **kw)) / (2 * hstep)
                assert abs(fd - grm[q]) <= mp.mpf(10) ** -25 * max(1, abs(fd)), (name, q, fd, grm[q])
        gr = [float(v) for v in grm]
        mu, vf, vd = mpo.predict(par, Xl, y, ev, Xs.tolist(), **kw)
        mpo.MATERN = None
        out["cases"].append({
            "name": name, "cov_form": cov, "extra_param": extra, "mean_form": mean, "noise_corr": nc, "X": X.tolist(), "y": y.tolist(),
            "events": ev.tolist(), "Xstar": Xs.tolist(), "par": [float(v) for v in par], "log_sc2_vec": lnc,
            "nlml": nl, "grad": gr, "pred_mean": [float(v) for v in mu], "pred_varf": [float(v) for v in vf],
            "pred_vard": [float(v) for v in vd]})
    # truncated-normal moments: exact (mpmath) and as the reference computes them (lossy 1 - pnorm)
    adj = []
    rng = np.random.default_rng(5)
    for al in list(np.linspace(-6, 4.9, 23)) + [5.5, 6.0, 6.5, 7.0, 7.1, 7.5, 9.0]:
        mu_, s_ = float(rng.standard_normal()), float(0.1 + rng.random())
        a_ = mu_ + float(al) * s_
        em, ev_ = mpo.truncated_moments(a_, mu_, s_)
        eo, vo = orc.adjust_one(a_, mu_, s_)
        adj.append({"a": a_, "mu": mu_, "s": s_, "alpha": float(al), "exact_mean": float(em), "exact_var": float(ev_),
                    "as_written_mean": eo, "as_written_var": vo})
    out["adjust"] = adj
    with open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden_small.json"), "w") as f:
        json.dump(out, f)
    print("wrote", len(out["cases"]), "cases,", len(adj), "adjust rows")


if __name__ == "__main__":
    main()
