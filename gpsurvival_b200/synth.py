"""Synthetic benchmark / test inputs.

This is the *input recipe* only — a restatement of what the reference's data layer does
before the hot path starts, so that every implementation (CUDA path, oracle, CPU
baseline) is fed identical arrays:

* X ~ N(0,1) iid, y = exp(chol(K)' d1 + sn d2)       (MakeSyntheticData.R:40-41,145-149)
* 'NormalLoopSample' right-censoring by |N(0, 50^2)|  (CensorData.R:36-48)
* train/test split                                   (GenerateData.R:131-152)
* column z-score with training mean / sd             (NormaliseExpressionData.R:18-31)

The log of the targets (ApplyGP.R:79) is NOT taken here: ApplyGP does it, as in the
reference.  Generating hyper-parameters sn2=0.01, sf2=0.5 (runSyntheticExp1.R:97) with
ell_gen = 1.1*sqrt(d); start hyper-parameters sn2=0.2, sf2=0.8, ell=1.7*sqrt(d)
(runSyntheticExp1.R:139, scaled by sqrt(d) so r^2 = O(1) at every d).

numpy only by default (runs anywhere); `gpu_sampler` optionally moves the one O(n^3) step of the recipe,
t(chol(K)) %*% d1, to the GPU through gps_chol_matvec.  No oracle imports.
"""
from __future__ import annotations

import dataclasses
import math

import numpy as np

BASE_SEED = 20160916


@dataclasses.dataclass(frozen=True)
class Config:
    name: str
    index: int
    n: int            # training rows
    m: int            # test rows
    d: int            # features
    c: int            # censored training rows (target)
    cov_form: str     # 'SqExp' | 'ARD'
    noise_corr: object  # False | 'noiseCorrLearned' | 'noiseCorrVec'
    n_fits: int = 1

    @property
    def n_hyp(self) -> int:
        p = self.d if self.cov_form == "ARD" else 1
        return 2 + p + (1 if self.noise_corr == "noiseCorrLearned" else 0)


# BASELINE.json configs[0..4]
CONFIGS = {
    "C1": Config("C1", 0, n=100, m=50, d=1, c=40, cov_form="SqExp", noise_corr=False),
    "C2": Config("C2", 1, n=2000, m=500, d=20, c=1000, cov_form="ARD", noise_corr="noiseCorrVec"),
    "C3": Config("C3", 2, n=285, m=57, d=10000, c=142, cov_form="ARD", noise_corr="noiseCorrVec"),
    "C4": Config("C4", 3, n=400, m=100, d=2000, c=200, cov_form="ARD", noise_corr="noiseCorrVec",
                 n_fits=512),
    "C5": Config("C5", 4, n=16384, m=4096, d=50, c=8192, cov_form="SqExp", noise_corr=False),
}


def _sqexp_gram(X, sf2, ell):
    Z = X / ell
    s = np.einsum("ij,ij->i", Z, Z)
    r2 = np.maximum(s[:, None] + s[None, :] - 2.0 * (Z @ Z.T), 0.0)
    np.fill_diagonal(r2, 0.0)
    return sf2 * np.exp(-0.5 * r2)


def gpu_sampler(device=0):
    """Returns sampler(X, sf2, ell, d1) -> t(chol(K)) %*% d1 computed on the GPU (gps_chol_matvec), for
    input generation at sizes where the host Cholesky would dominate (C5: 20 480 rows)."""
    def sample(X, sf2, ell, d1):
        from . import gp
        fit = gp.Fit(device)
        fit.set_options("SqExp", "Zero", False)
        fit.set_data(X, np.zeros(X.shape[0]), None)
        jitter = 1e-8
        while True:
            try:
                out = fit.chol_matvec(np.log([jitter, sf2, ell]), d1)
                break
            except gp.NotPositiveDefinite:
                jitter *= 10.0
        fit.close()
        return out
    return sample


def make_problem(n, m, d, c, seed, sn2_gen=0.01, sf2_gen=0.5, sampler=None):
    """One synthetic right-censored data set in the shape of trainingTestStructure.

    Returns a dict with trainingData (n x d), trainingTargets (n,), events (n,, 0=censored),
    testData (m x d), testTargets (m,), testTargetsPreCensoring, trainingTargetsPreCensoring —
    all on the *time* scale (ApplyGP takes the log)."""
    rng = np.random.Generator(np.random.PCG64(seed))
    nt = n + m
    X = rng.standard_normal((nt, d))
    ell = 1.1 * math.sqrt(d)
    d1 = rng.standard_normal(nt)
    if sampler is not None:
        f = sampler(X, sf2_gen, ell, d1)
    else:
        K = _sqexp_gram(X, sf2_gen, ell)
        jitter = 1e-8
        while True:
            try:
                L = np.linalg.cholesky(K + jitter * np.eye(nt))
                break
            except np.linalg.LinAlgError:
                jitter *= 10.0
        f = L @ d1
    f = f + math.sqrt(sn2_gen) * rng.standard_normal(nt)
    y_pre = np.exp(f)
    # the first n rows are the training set (rows are iid, so a fixed split is a random split)
    y = y_pre.copy()
    events = np.ones(nt)
    # censor training and test pools separately so the training count is exactly c
    for lo, hi, target in ((0, n, c), (n, nt, int(round(m * (c / max(n, 1)))))):
        pool = list(range(lo, hi))
        done = 0
        guard = 0
        while done < target and pool:
            k = int(rng.integers(len(pool)))
            t = abs(rng.normal(0.0, 50.0))
            guard += 1
            if guard > 200000 * max(target, 1):
                raise RuntimeError("censoring loop did not terminate")
            if y[pool[k]] > t:
                y[pool[k]] = t
                events[pool[k]] = 0
                pool.pop(k)
                done += 1
    Xtr, Xte = X[:n], X[n:]
    mean = Xtr.mean(axis=0)
    sd = Xtr.std(axis=0, ddof=1)
    keep = sd != 0
    Xtr = ((Xtr - mean) / np.where(keep, sd, 1.0))[:, keep]
    Xte = ((Xte - mean) / np.where(keep, sd, 1.0))[:, keep]
    return {
        "trainingData": np.asfortranarray(Xtr), "trainingTargets": y[:n].copy(),
        "events": events[:n].copy(), "testData": np.asfortranarray(Xte),
        "testTargets": y[n:].copy(), "testEvents": events[n:].copy(),
        "trainingTargetsPreCensoring": y_pre[:n].copy(), "testTargetsPreCensoring": y_pre[n:].copy(),
        "nTraining": n, "nTest": m, "dimension": int(Xtr.shape[1]),
    }


def make_config_problem(cfg: Config, fit_index: int = 0, scale: float = 1.0):
    """Problem for one of the BASELINE configs; ``scale`` < 1 shrinks n, m, c (tests)."""
    n = max(4, int(round(cfg.n * scale)))
    m = max(2, int(round(cfg.m * scale)))
    c = min(n - 2, int(round(cfg.c * scale)))
    return make_problem(n, m, cfg.d, c, BASE_SEED + cfg.index + 1000 * fit_index)


def start_log_hyp(d, cov_form, perturb_rng=None):
    """logHypStart (runSyntheticExp1.R:139) with ell scaled by sqrt(d); optional restart
    perturbation N(0, 0.25^2) in log space."""
    p = d if cov_form == "ARD" else 1
    lh = {"noise": math.log(0.2), "func": math.log(0.8),
          "length": np.full(p, math.log(1.7 * math.sqrt(d))), "mean": np.zeros(d + 1)}
    if perturb_rng is not None:
        lh["noise"] += 0.25 * perturb_rng.standard_normal()
        lh["func"] += 0.25 * perturb_rng.standard_normal()
        lh["length"] = lh["length"] + 0.25 * perturb_rng.standard_normal(p)
    return lh


def start_par(cfg: Config, d=None, perturb_rng=None):
    """Flat optimiser vector for ``cfg`` (Zero mean): [log sn2, log sf2, log ell.., (log sc2)]."""
    d = cfg.d if d is None else d
    lh = start_log_hyp(d, cfg.cov_form, perturb_rng)
    vec = [lh["noise"], lh["func"], *lh["length"]]
    if cfg.noise_corr == "noiseCorrLearned":
        vec.append(lh["noise"])
    return np.array(vec, dtype=np.float64)
