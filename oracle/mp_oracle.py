"""TEST INFRASTRUCTURE ONLY — independent 50-digit mpmath restatement of the same maths
as oracle/gp_oracle.py, for tiny cases (n <= ~16).  Its purpose is to bound the FP64
oracle's own error, because the reference ships nothing to pin against (SURVEY.md F2,
§8c "What pins results instead" item 2).  Written from the formulas, not from
gp_oracle.py, and deliberately in a different style (explicit loops, exact inverse).

Formulas: ForOptimisationLog.R:34-40,63 (NLML), OptimisationGradientLog.R:50-57 in
W-form (gradient), GPPredict.R:49-56 (prediction), AdjustTrainingSurvivalMeanVariance.R:
10,21-26 (truncated-normal moments, here with the *exact* upper tail).
"""
from __future__ import annotations

import mpmath as mp

mp.mp.dps = 50


MATERN = None   # module-level switch: None (SqExp / ARD) or 1, 3, 5 (CovFunc.R:33-37)


def _cov(x1, x2, sf2, ell):
    n1, n2, d = len(x1), len(x2), len(x1[0])
    K = mp.zeros(n1, n2)
    for i in range(n1):
        for j in range(n2):
            r2 = mp.mpf(0)
            for k in range(d):
                lk = ell[k] if len(ell) > 1 else ell[0]
                r2 += ((mp.mpf(x1[i][k]) - mp.mpf(x2[j][k])) / lk) ** 2
            if MATERN is None:
                K[i, j] = sf2 * mp.exp(-r2 / 2)
            else:
                r = mp.sqrt(r2)
                if MATERN == 1:
                    K[i, j] = sf2 * mp.exp(-r)
                elif MATERN == 3:
                    K[i, j] = sf2 * (1 + mp.sqrt(3) * r) * mp.exp(-mp.sqrt(3) * r)
                else:
                    K[i, j] = sf2 * (1 + mp.sqrt(5) * r + 5 * r2 / 3) * mp.exp(-mp.sqrt(5) * r)
    return K


def _cov_d(x, sf2, ell):
    """Kd_ij = -sf2 f'(r)/r, the factor in dK_ij / dlog(ell_k) = Kd_ij * (x_ik - x_jk)^2 / ell_k^2 for
    K = sf2 f(r), r = |x_i/ell - x_j/ell|.  Derived here (the reference has no Matern gradient,
    OptimisationGradientLog.R:6): SqExp f = exp(-r^2/2) gives Kd = K; Matern 1/2: sf2 exp(-r)/r (0 at r = 0,
    where the distance factor vanishes); 3/2: 3 sf2 exp(-sqrt3 r); 5/2: 5/3 sf2 (1 + sqrt5 r) exp(-sqrt5 r)."""
    n, d = len(x), len(x[0])
    Kd = mp.zeros(n, n)
    for i in range(n):
        for j in range(n):
            r2 = mp.mpf(0)
            for k in range(d):
                lk = ell[k] if len(ell) > 1 else ell[0]
                r2 += ((mp.mpf(x[i][k]) - mp.mpf(x[j][k])) / lk) ** 2
            r = mp.sqrt(r2)
            if MATERN is None:
                Kd[i, j] = sf2 * mp.exp(-r2 / 2)
            elif MATERN == 1:
                Kd[i, j] = sf2 * mp.exp(-r) / r if r > 0 else mp.mpf(0)
            elif MATERN == 3:
                Kd[i, j] = 3 * sf2 * mp.exp(-mp.sqrt(3) * r)
            else:
                Kd[i, j] = mp.mpf(5) / 3 * sf2 * (1 + mp.sqrt(5) * r) * mp.exp(-mp.sqrt(5) * r)
    return Kd


def _unpack(par, d, ard, linear_mean, gps2):
    par = [p if isinstance(p, mp.mpf) else mp.mpf(float(p)) for p in par]
    sc2 = None
    if gps2:
        sc2 = mp.exp(par[-1])
        par = par[:-1]
    p = d if ard else 1
    sn2, sf2 = mp.exp(par[0]), mp.exp(par[1])
    ell = [mp.exp(v) for v in par[2:2 + p]]
    w = par[len(par) - d - 1:] if linear_mean else [mp.mpf(0)] * (d + 1)
    return sn2, sf2, ell, w, sc2


def _mean(w, x):
    d = len(x[0])
    return [sum(mp.mpf(x[i][k]) * w[k] for k in range(d)) + w[d] for i in range(len(x))]


def _ky(par, X, events, ard, linear_mean, gps2, sc2_vec):
    n, d = len(X), len(X[0])
    sn2, sf2, ell, w, sc2 = _unpack(par, d, ard, linear_mean, gps2)
    K = _cov(X, X, sf2, ell)
    Ky = K.copy()
    c = 0
    for i in range(n):
        Ky[i, i] += sn2
        if events[i] == 0:
            if gps2:
                Ky[i, i] += sc2
            elif sc2_vec is not None:
                Ky[i, i] += mp.exp(mp.mpf(float(sc2_vec[c])))
            c += 1
    return K, Ky, (sn2, sf2, ell, w, sc2)


def nlml(par, X, y, events, ard=False, linear_mean=False, gps2=False, log_sc2_vec=None):
    n = len(X)
    K, Ky, (sn2, sf2, ell, w, sc2) = _ky(par, X, events, ard, linear_mean, gps2, log_sc2_vec)
    m = _mean(w, X)
    r = mp.matrix([mp.mpf(float(y[i])) - m[i] for i in range(n)])
    L = mp.cholesky(Ky)
    alpha = mp.lu_solve(Ky, r)
    quad = sum(r[i] * alpha[i] for i in range(n))
    return quad / 2 + sum(mp.log(L[i, i]) for i in range(n)) + mp.mpf(n) / 2 * mp.log(2 * mp.pi)


def nlml_grad(par, X, y, events, ard=False, linear_mean=False, gps2=False, log_sc2_vec=None):
    """True derivative of nlml() w.r.t. par, with the reference's quirk that the W
    quadratic forms use y rather than y - m (OptimisationGradientLog.R:51-53,57)."""
    n, d = len(X), len(X[0])
    K, Ky, (sn2, sf2, ell, w, sc2) = _ky(par, X, events, ard, linear_mean, gps2, log_sc2_vec)
    Kinv = Ky ** -1
    yv = mp.matrix([mp.mpf(float(v)) for v in y])
    a = Kinv * yv
    W = a * a.T - Kinv
    g = [-(sum(W[i, i] for i in range(n))) * sn2 / 2,
         -sum(W[i, j] * K[i, j] for i in range(n) for j in range(n)) / 2]
    p = d if ard else 1
    Kd = K if MATERN is None else _cov_d(X, sf2, ell)
    for k in range(p):
        s = mp.mpf(0)
        for i in range(n):
            for j in range(n):
                if ard:
                    r2 = ((mp.mpf(X[i][k]) - mp.mpf(X[j][k])) / ell[k]) ** 2
                else:
                    r2 = sum(((mp.mpf(X[i][q]) - mp.mpf(X[j][q])) / ell[0]) ** 2 for q in range(d))
                s += W[i, j] * Kd[i, j] * r2
        g.append(-s / 2)
    if linear_mean:
        m = _mean(w, X)
        ac = Kinv * mp.matrix([yv[i] - m[i] for i in range(n)])
        for k in range(d):
            g.append(-sum(mp.mpf(X[i][k]) * ac[i] for i in range(n)))
        g.append(-sum(ac[i] for i in range(n)))
    if gps2:
        g.append(-sum(W[i, i] for i in range(n) if events[i] == 0) * sc2 / 2)
    return g


def predict(par, X, y, events, Xstar, ard=False, linear_mean=False, gps2=False, log_sc2_vec=None):
    n = len(X)
    K, Ky, (sn2, sf2, ell, w, sc2) = _ky(par, X, events, ard, linear_mean, gps2, log_sc2_vec)
    m = _mean(w, X)
    ms = _mean(w, Xstar)
    Kinv = Ky ** -1
    alpha = Kinv * mp.matrix([mp.mpf(float(y[i])) - m[i] for i in range(n)])
    ks = _cov(X, Xstar, sf2, ell)
    mu, varf = [], []
    for j in range(len(Xstar)):
        col = ks[:, j]
        mu.append(ms[j] + sum(col[i] * alpha[i] for i in range(n)))
        varf.append(sf2 - (col.T * Kinv * col)[0, 0])
    return mu, varf, [v + sn2 for v in varf]


def truncated_moments(a, mu, s):
    """Lower-truncated normal mean / variance with the exact upper tail (no 1-Phi loss)."""
    a, mu, s = mp.mpf(float(a)), mp.mpf(float(mu)), mp.mpf(float(s))
    al = (a - mu) / s
    lam = mp.npdf(al) / (mp.erfc(al / mp.sqrt(2)) / 2)
    return mu + s * lam, s ** 2 * (1 - lam * (lam - al))


def near_pd(x, eig_tol="1e-6", conv_tol="1e-7", posd_tol="1e-8", maxit=100):
    """Matrix::nearPD with the defaults the reference uses (call site ForOptimisationLog.R:43-44) in 50-digit arithmetic:
    Higham (2002) alternating projections with Dykstra's correction, eigenvalues below eig.tol * lambda_max dropped, relative
    infinity-norm convergence, then the posd.tol eigenvalue floor and the diagonal-preserving rescale.  Written from the
    published algorithm with explicit loops (mp.eigsy for the symmetric eigenproblem) — an independent pin of the FP64
    restatement gp_oracle.near_pd for tiny matrices.  Returns (matrix as nested lists of mpf, converged, iterations)."""
    eig_tol, conv_tol, posd_tol = mp.mpf(eig_tol), mp.mpf(conv_tol), mp.mpf(posd_tol)
    n = len(x)
    X = mp.matrix(n, n)
    for i in range(n):
        for j in range(n):
            X[i, j] = (mp.mpf(x[i][j]) + mp.mpf(x[j][i])) / 2

    def inf_norm(A):
        return max(sum(abs(A[i, j]) for j in range(n)) for i in range(n))

    def rebuild(vals, Q, keep):
        out = mp.zeros(n, n)
        for k in range(n):
            if not keep[k]:
                continue
            for i in range(n):
                qik = Q[i, k] * vals[k]
                for j in range(n):
                    out[i, j] += qik * Q[j, k]
        return out

    D_S = mp.zeros(n, n)
    it, converged = 0, False
    while it < maxit and not converged:
        Y = X.copy()
        R = Y - D_S
        vals, Q = mp.eigsy(R)
        vals = [vals[k] for k in range(n)]
        lam = max(vals)
        keep = [v > eig_tol * lam for v in vals]
        if not any(keep):
            raise ArithmeticError("Matrix seems negative semi-definite")
        X = rebuild(vals, Q, keep)
        D_S = X - R
        conv = inf_norm(Y - X) / inf_norm(Y)
        it += 1
        converged = conv <= conv_tol
    vals, Q = mp.eigsy(X)
    vals = [vals[k] for k in range(n)]
    eps = posd_tol * abs(max(vals))
    if min(vals) < eps:
        o_diag = [X[i, i] for i in range(n)]
        X = rebuild([v if v >= eps else eps for v in vals], Q, [True] * n)
        D = [mp.sqrt(max(eps, o_diag[i]) / X[i, i]) for i in range(n)]
        for i in range(n):
            for j in range(n):
                X[i, j] = D[i] * X[i, j] * D[j]
    return [[X[i, j] for j in range(n)] for i in range(n)], converged, it
