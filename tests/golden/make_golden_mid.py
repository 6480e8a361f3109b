"""Generate tests/golden/golden_mid.json: mpmath (30-digit) expectations at n = 65, 130, 300 and 600.

golden_small.json stops at n <= 13, i.e. inside one 64-column leaf of the CUDA Cholesky.  These three sizes cross
the leaf boundary (65), a leaf pair (130) and a 256-wide block of the two-level scheme (300), so something that is
independent of the numpy oracle pins those routes too (the reference ships no vectors and R is absent — SURVEY.md
F1 / F2).  The arithmetic is oracle/mp_oracle.py's formulas with the O(n^3) steps written once (one Cholesky, one
explicit inverse) so that n = 300 finishes in minutes.  Re-run:  python tests/golden/make_golden_mid.py
"""
import json
import os
import sys
import time

import mpmath as mp
import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from gpsurvival_b200 import synth            # noqa: E402

mp.mp.dps = 30

CASES = [
    # name, n, m, d, c, cov, noise_corr
    ("n65_sqexp_gps1", 65, 3, 2, 26, "SqExp", False),
    ("n130_ard_gps3", 130, 3, 3, 52, "ARD", "noiseCorrVec"),
    ("n300_ard_gps3", 300, 3, 4, 120, "ARD", "noiseCorrVec"),
    ("n600_ard_gps3", 600, 3, 3, 240, "ARD", "noiseCorrVec"),      # three 256-wide blocks: the recursion above the blocks
]


def chol_lower(A, n):
    L = [[mp.mpf(0)] * n for _ in range(n)]
    for j in range(n):
        s = A[j][j] - mp.fsum(L[j][k] * L[j][k] for k in range(j))
        L[j][j] = mp.sqrt(s)
        inv = 1 / L[j][j]
        for i in range(j + 1, n):
            L[i][j] = (A[i][j] - mp.fdot(L[i][:j], L[j][:j])) * inv
    return L


def lower_inverse(L, n):
    X = [[mp.mpf(0)] * n for _ in range(n)]
    for j in range(n):
        X[j][j] = 1 / L[j][j]
        for i in range(j + 1, n):
            X[i][j] = -mp.fdot(L[i][j:i], [X[k][j] for k in range(j, i)]) / L[i][i]
    return X


def main():
    out = {"note": "30-digit mpmath restatement rounded to FP64; sizes that cross the CUDA Cholesky's 64-column leaf, a leaf "
                   "pair and a 256-wide diagonal block (parity is unpinned by the reference itself: no R, no vectors)",
           "cases": []}
    for idx, (name, n, m, d, c, cov, nc) in enumerate(CASES):
        t0 = time.time()
        pb = synth.make_problem(n, m, d, c, 4100 + idx)
        X, y, ev, Xs = pb["trainingData"], np.log(pb["trainingTargets"]), pb["events"], pb["testData"]
        rng = np.random.default_rng(300 + idx)
        p = d if cov == "ARD" else 1
        par = [np.log(0.2), np.log(0.8)] + list(np.log(1.7 * np.sqrt(d)) + 0.3 * rng.standard_normal(p))
        lnc = None
        if nc == "noiseCorrVec":
            lnc = (np.log(0.2) + 0.5 * rng.standard_normal(int(np.sum(ev == 0)))).tolist()
        sn2, sf2 = mp.exp(mp.mpf(par[0])), mp.exp(mp.mpf(par[1]))
        ell = [mp.exp(mp.mpf(v)) for v in par[2:2 + p]]
        Xm = [[mp.mpf(float(v)) for v in row] for row in X]
        Xsm = [[mp.mpf(float(v)) for v in row] for row in Xs]
        yv = [mp.mpf(float(v)) for v in y]

        def lk(k):
            return ell[k] if p > 1 else ell[0]

        def kfun(a, b):
            r2 = mp.fsum(((a[k] - b[k]) / lk(k)) ** 2 for k in range(d))
            return sf2 * mp.exp(-r2 / 2)
        K = [[None] * n for _ in range(n)]
        for i in range(n):
            for j in range(i + 1):
                K[i][j] = K[j][i] = kfun(Xm[i], Xm[j])
        Ky = [row[:] for row in K]
        cc = 0
        for i in range(n):
            Ky[i][i] += sn2
            if ev[i] == 0:
                if lnc is not None:
                    Ky[i][i] += mp.exp(mp.mpf(float(lnc[cc])))
                cc += 1
        L = chol_lower(Ky, n)
        Li = lower_inverse(L, n)
        # Kinv = Li' Li (symmetric)
        LiT = [[Li[k][i] for k in range(n)] for i in range(n)]      # LiT[i] = column i of Li
        Kinv = [[None] * n for _ in range(n)]
        for i in range(n):
            for j in range(i + 1):
                Kinv[i][j] = Kinv[j][i] = mp.fdot(LiT[i][i:], LiT[j][i:])
        alpha = [mp.fdot(Kinv[i], yv) for i in range(n)]
        nlml = mp.fdot(yv, alpha) / 2 + mp.fsum(mp.log(L[i][i]) for i in range(n)) + mp.mpf(n) / 2 * mp.log(2 * mp.pi)
        # gradient (zero mean: y - m = y), W = a a' - Kinv
        g = [-mp.fsum(alpha[i] * alpha[i] - Kinv[i][i] for i in range(n)) * sn2 / 2]
        M = [[(alpha[i] * alpha[j] - Kinv[i][j]) * K[i][j] for j in range(n)] for i in range(n)]
        g.append(-mp.fsum(mp.fsum(M[i]) for i in range(n)) / 2)
        for k in range(p):
            s = mp.mpf(0)
            for i in range(n):
                for j in range(i):
                    if p > 1:
                        r2 = ((Xm[i][k] - Xm[j][k]) / ell[k]) ** 2
                    else:
                        r2 = mp.fsum(((Xm[i][q] - Xm[j][q]) / ell[0]) ** 2 for q in range(d))
                    s += 2 * M[i][j] * r2
            g.append(-s / 2)
        mu, vf = [], []
        for j in range(m):
            col = [kfun(Xm[i], Xsm[j]) for i in range(n)]
            mu.append(mp.fdot(col, alpha))
            v = [mp.fdot(Li[i][:i + 1], col[:i + 1]) for i in range(n)]
            vf.append(sf2 - mp.fdot(v, v))
        out["cases"].append({
            "name": name, "cov_form": cov, "mean_form": "Zero", "noise_corr": nc, "X": X.tolist(), "y": y.tolist(),
            "events": ev.tolist(), "Xstar": Xs.tolist(), "par": [float(v) for v in par], "log_sc2_vec": lnc,
            "nlml": float(nlml), "grad": [float(v) for v in g], "pred_mean": [float(v) for v in mu],
            "pred_varf": [float(v) for v in vf], "pred_vard": [float(v + sn2) for v in vf]})
        print(name, "done in %.0f s" % (time.time() - t0), flush=True)
    with open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden_mid.json"), "w") as f:
        json.dump(out, f)
    print("wrote", len(out["cases"]), "cases")


if __name__ == "__main__":
    main()
