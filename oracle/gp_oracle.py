"""TEST INFRASTRUCTURE ONLY (see oracle/__init__.py) — numpy FP64 restatement of the
GP training / prediction hot path of kllloyd/GPSurvival.  PARITY UNPINNED (no reference
golden vectors exist; R is absent from the image).

Every function cites the reference lines it follows (paths relative to
/root/reference/code/).  The restatement keeps the reference's *operation order*
where it matters numerically (direct-difference distance then square, ``1 - pnorm``,
general ``solve`` on the triangular factor) so that the same file doubles as the CPU
baseline ("reference algorithm restated in numpy; R unavailable in image").

Conventions: matrices are numpy arrays in R's logical layout (row i, column j);
``events`` is 0/1 (0 = censored); hyper-parameter vector layout is
``[log sn2, log sf2, log ell (1 | d), (w_1..w_d, b — Linear mean only), (log sc2 — GPS2 only)]``
(ForOptimisationLog.R:9-31, GPRegression.R:22-25).
"""
from __future__ import annotations

import math

import numpy as np
import scipy.linalg as sla

# --------------------------------------------------------------------------------------
# option strings, exactly as the reference spells them (SURVEY.md §5 "Config")
# --------------------------------------------------------------------------------------
COV_FORMS = ("SqExp", "ARD", "Matern")
MEAN_FORMS = ("Zero", "Linear")
NOISE_CORRS = (False, "none", "noiseCorrLearned", "noiseCorrVec")


def _has_corr(noise_corr) -> bool:
    return noise_corr in ("noiseCorrLearned", "noiseCorrVec")


# --------------------------------------------------------------------------------------
# CovFunc.R
# --------------------------------------------------------------------------------------
def _scale_inputs(x, ell, ard_mode):
    """``t(x1)/as.numeric(lengthHyp)`` (CovFunc.R:25,27).

    ``ard_mode='intended'``: column k divided by ell[k] (north_star item 1).
    ``ard_mode='as_written'``: R recycles the length-p vector down the column-major
    n x d matrix, so element (i, j) is divided by ell[(i + j*n) mod p].
    With p == 1 both modes coincide.
    """
    x = np.asarray(x, dtype=np.float64)
    ell = np.atleast_1d(np.asarray(ell, dtype=np.float64)).ravel()
    n, d = x.shape
    p = ell.size
    if p == 1:
        return x / ell[0]
    if ard_mode == "intended":
        if p != d:
            raise ValueError("ARD needs one length-scale per column")
        return x / ell[None, :]
    if ard_mode == "as_written":
        i = np.arange(n)[:, None]
        j = np.arange(d)[None, :]
        return x / ell[(i + j * n) % p]
    raise ValueError(f"ard_mode {ard_mode!r}")


def rdist(z1, z2):
    """fields::rdist — Euclidean distance by direct differences (call sites CovFunc.R:25,27).

    Accumulates (a-b)^2 over the feature axis in index order, then sqrt; memory is
    O(n1*n2) regardless of d.
    """
    n1, d = z1.shape
    n2 = z2.shape[0]
    acc = np.zeros((n1, n2))
    for k in range(d):
        diff = z1[:, k][:, None] - z2[:, k][None, :]
        acc += diff * diff
    return np.sqrt(acc)


def cov_func(x1, x2, sf2, ell, cov_form="SqExp", ard_mode="intended", extra_param=None):
    """CovFunc.R:21-29,33-39 — SqExp / ARD: K_ij = sf2 * exp(-(rdist(x1/ell, x2/ell)_ij)^2 / 2);
    Matern with maternParamFlag = as.character(extraParam) in '1', '3', '5' (:31-37).

    Noise is NOT added here (callers add it, ForOptimisationLog.R:35).
    """
    if cov_form not in COV_FORMS:
        raise NotImplementedError(cov_form)  # RF/InformedARD: SURVEY.md §2
    distance = rdist(_scale_inputs(x1, ell, ard_mode), _scale_inputs(x2, ell, ard_mode))
    if cov_form == "Matern":
        flag = str(int(extra_param)) if extra_param is not None else None
        if flag == "1":
            return float(sf2) * np.exp(-distance)                                                   # :34
        if flag == "3":
            return float(sf2) * (1 + (math.sqrt(3) * distance)) * np.exp(-(math.sqrt(3) * distance))   # :35
        if flag == "5":
            return float(sf2) * (1 + (math.sqrt(5) * distance) + (5 * distance ** 2) / 3) * np.exp(-(math.sqrt(5) * distance))  # :36
        raise ValueError(f"maternParam {extra_param!r}")
    return float(sf2) * np.exp(-(distance ** 2) / 2.0)   # CovFunc.R:38-39


# --------------------------------------------------------------------------------------
# MeanFunc.R
# --------------------------------------------------------------------------------------
def mean_func(mean_hyp, x, mean_form):
    """MeanFunc.R:8-13 — Zero -> 0_n ; Linear -> x %*% w + b."""
    x = np.asarray(x, dtype=np.float64)
    n, d = x.shape
    if mean_form == "Zero":
        return np.zeros(n)
    if mean_form == "Linear":
        mean_hyp = np.asarray(mean_hyp, dtype=np.float64).ravel()
        return x @ mean_hyp[:d] + mean_hyp[d]
    raise ValueError(mean_form)


# --------------------------------------------------------------------------------------
# hyper-parameter unpacking shared by ForOptimisationLog.R:9-31 and
# OptimisationGradientLog.R:10-32
# --------------------------------------------------------------------------------------
def unpack_log_hyp(par, d, cov_form, mean_form, noise_corr, log_hyp_noise_corr=None):
    par = np.asarray(par, dtype=np.float64).ravel()
    if noise_corr == "noiseCorrLearned":            # ForOptimisationLog.R:10-13
        log_hyp_noise_corr = par[-1:]
        par = par[:-1]
    p = d if cov_form == "ARD" else 1               # :14-21 (SqExp | Matern share one length-scale)
    log_sn2 = par[0]
    log_sf2 = par[1]
    log_ell = par[2:2 + p]
    if mean_form == "Zero":                         # :31
        mean_hyp = np.zeros(d + 1)
    else:
        mean_hyp = par[par.size - d - 1:]
    return log_sn2, log_sf2, log_ell, mean_hyp, log_hyp_noise_corr


def _noise_diag(n, log_sn2, events, noise_corr, log_hyp_noise_corr):
    """sn2*I [+ diag(sc2 on censored rows)] (ForOptimisationLog.R:35-40)."""
    diag = np.full(n, math.exp(log_sn2))
    if _has_corr(noise_corr):
        corr = np.zeros(n)
        cens = np.asarray(events).ravel() == 0
        corr[cens] = np.exp(np.asarray(log_hyp_noise_corr, dtype=np.float64).ravel())
        diag = diag + corr
    return diag


def _chol_lower(Ky):
    """chol(K) then t() (ForOptimisationLog.R:41-42,54). Raises LinAlgError if not PD (the caller decides about the
    nearPD fallback of :43-51)."""
    return np.linalg.cholesky(Ky)


def near_pd(x, eig_tol=1e-6, conv_tol=1e-7, posd_tol=1e-8, maxit=100):
    """Matrix::nearPD(x) with its defaults (corr=FALSE, keepDiag=FALSE, do2eigen=TRUE, doSym=FALSE, doDykstra=TRUE,
    conv.norm.type="I") — third-party (Matrix is not under /root/reference, version unpinned; call site
    ForOptimisationLog.R:43-44), restated from its published algorithm: Higham's (2002) alternating projections with
    Dykstra's correction.  Each iteration projects R = Y - D_S onto the matrices whose eigenvalues exceed
    eig.tol * lambda_max (smaller ones are DROPPED, not clipped), until the relative infinity-norm change is below
    conv.tol; then (do2eigen) eigenvalues below posd.tol * lambda_max are raised to that value and the result is
    rescaled to keep the diagonal.  Returns (mat, converged, iterations)."""
    x = np.asarray(x, dtype=np.float64)
    x = 0.5 * (x + x.T)                                    # ensureSymmetry (symmpart)
    n = x.shape[0]
    D_S = np.zeros_like(x)
    X = x.copy()
    it, converged = 0, False
    while it < maxit and not converged:
        Y = X
        R = Y - D_S
        d, Q = np.linalg.eigh(R)
        p = d > eig_tol * d[-1]
        if not np.any(p):
            raise np.linalg.LinAlgError("Matrix seems negative semi-definite")
        Qp = Q[:, p]
        X = (Qp * d[p]) @ Qp.T
        D_S = X - R
        conv = np.linalg.norm(Y - X, np.inf) / np.linalg.norm(Y, np.inf)
        it += 1
        converged = conv <= conv_tol
    d, Q = np.linalg.eigh(X)
    eps = posd_tol * abs(d[-1])
    if d[0] < eps:
        d = np.where(d < eps, eps, d)
        o_diag = np.diag(X).copy()
        X = Q @ (d[:, None] * Q.T)
        D = np.sqrt(np.maximum(eps, o_diag) / np.diag(X))
        X = D[:, None] * X * D[None, :]
    return X, converged, it


def _solve_as_written(L, rhs):
    """``solve(t(L), solve(L, rhs))`` — two *general* LU solves (dgesv) on the triangular
    factor, as the reference does (ForOptimisationLog.R:55)."""
    return np.linalg.solve(L.T, np.linalg.solve(L, rhs))


# --------------------------------------------------------------------------------------
# ForOptimisationLog.R
# --------------------------------------------------------------------------------------
def nlml(par, X, y, events, cov_form="SqExp", mean_form="Zero", noise_corr=False,
         log_hyp_noise_corr=None, ard_mode="intended", as_written_ops=True, extra_param=None, near_pd_fallback=False):
    """Negative log marginal likelihood (ForOptimisationLog.R:1-68, imposePriors=FALSE).

    ``as_written_ops=True`` repeats the reference's redundant work (Cholesky probed then
    recomputed, :41-42; general solves, :55) — this is what the CPU baseline times.
    ``near_pd_fallback=True``: when chol(Ky) fails, factorise Matrix::nearPD(Ky)$mat instead (:43-51).
    """
    X = np.asarray(X, dtype=np.float64)
    y = np.asarray(y, dtype=np.float64).ravel()
    n, d = X.shape
    log_sn2, log_sf2, log_ell, mean_hyp, lnc = unpack_log_hyp(
        par, d, cov_form, mean_form, noise_corr, log_hyp_noise_corr)
    m = mean_func(mean_hyp, X, mean_form)                                    # :34
    K = cov_func(X, X, math.exp(log_sf2), np.exp(log_ell), cov_form, ard_mode, extra_param)
    Ky = K + np.diag(_noise_diag(n, log_sn2, events, noise_corr, lnc))       # :35-40
    if near_pd_fallback:
        try:
            L = _chol_lower(Ky)                                              # :41-42
        except np.linalg.LinAlgError:
            near, converged, _ = near_pd(Ky)                                 # :43-45
            if not converged:
                raise
            L = _chol_lower(near)                                            # :48
    else:
        if as_written_ops:
            _chol_lower(Ky)                                                  # :41 (probe)
        L = _chol_lower(Ky)                                                  # :42,54
    if as_written_ops:
        alpha = _solve_as_written(L, y - m)                                  # :55
    else:
        alpha = sla.solve_triangular(L, sla.solve_triangular(L, y - m, lower=True),
                                     lower=True, trans="T")
    return 0.5 * float((y - m) @ alpha) + float(np.sum(np.log(np.diag(L)))) \
        + (n / 2.0) * math.log(2.0 * math.pi)                                # :63


# --------------------------------------------------------------------------------------
# OptimisationGradientLog.R
# --------------------------------------------------------------------------------------
def matern_deriv_factor(distance, sf2, flag):
    """Kd = -sf2 f'(r)/r for K = sf2 f(r): dK_ij/dlog(ell) = Kd_ij * r_ij^2 (one shared length-scale, CovFunc.R:33-37).
    Derived, not restated: the reference has no Matern gradient (OptimisationGradientLog.R:6)."""
    r = np.asarray(distance, dtype=np.float64)
    if flag == "1":
        with np.errstate(divide="ignore", invalid="ignore"):
            return np.where(r > 0, float(sf2) * np.exp(-r) / r, 0.0)
    if flag == "3":
        return 3.0 * float(sf2) * np.exp(-(math.sqrt(3) * r))
    if flag == "5":
        return (5.0 / 3.0) * float(sf2) * (1 + math.sqrt(5) * r) * np.exp(-(math.sqrt(5) * r))
    raise ValueError(f"maternParam {flag!r}")


def nlml_grad(par, X, y, events, cov_form="SqExp", mean_form="Zero", noise_corr=False,
              log_hyp_noise_corr=None, ard_mode="intended", grad_mode="analytic", extra_param=None):
    """Gradient of the NLML w.r.t. ``par`` (OptimisationGradientLog.R:1-94).

    SqExp: the reference branch (:50-57), restated operation by operation — it is the
    analytic derivative for a zero mean.  Reference quirk kept: the quadratic forms use
    ``trainingTargets``, not ``trainingTargets - meanTraining`` (:51-53,57).

    ARD, ``grad_mode='analytic'``: the true derivative (north_star item 4):
    W = aa' - Ky^-1, M = W o K, d/dlog ell_k = -(1/ell_k^2) (sum_i rowsum(M)_i x_ik^2 -
    sum_i x_ik (M X)_ik); noise/func terms as in the SqExp branch with their exp() chain
    factors.  ``grad_mode='as_written'`` restates :58-69 literally (SURVEY.md F4 — it is
    not the derivative of the reference's own likelihood; kept for documentation).
    """
    X = np.asarray(X, dtype=np.float64)
    y = np.asarray(y, dtype=np.float64).ravel()
    n, d = X.shape
    log_sn2, log_sf2, log_ell, mean_hyp, lnc = unpack_log_hyp(
        par, d, cov_form, mean_form, noise_corr, log_hyp_noise_corr)
    sn2, sf2, ell = math.exp(log_sn2), math.exp(log_sf2), np.exp(log_ell)
    m = mean_func(mean_hyp, X, mean_form)                                    # :35
    r = rdist(X, X)                                                          # :36 (dist)
    K = cov_func(X, X, sf2, ell, cov_form, ard_mode, extra_param)            # :37
    Ky = K + np.diag(_noise_diag(n, log_sn2, events, noise_corr, lnc))       # :39-46
    Kinv = np.linalg.solve(Ky, np.eye(n))                                    # solve(Ky)
    cens = (np.asarray(events).ravel() == 0).astype(np.float64)

    def half_trace_minus_quad(D):
        # (1/2 tr(Kinv D) - 1/2 y' Kinv D Kinv y) — the pattern of :51-53,57
        a = Kinv @ y
        return 0.5 * float(np.trace(Kinv @ D)) - 0.5 * float(a @ (D @ a))

    dcorr = None
    if cov_form == "SqExp":
        l = float(ell[0])
        E = np.exp(-(r ** 2) / (2.0 * l ** 2))
        dnoise = half_trace_minus_quad(np.eye(n)) * sn2                      # :51
        dfunc = half_trace_minus_quad(E) * sf2                               # :52
        dl = np.array([half_trace_minus_quad(sf2 * (r ** 2 / l ** 3) * E) * l])   # :53
        if noise_corr == "noiseCorrLearned":                                 # :57
            dcorr = half_trace_minus_quad(np.diag(cens)) * math.exp(float(np.ravel(lnc)[0]))
    elif cov_form == "ARD" and grad_mode == "analytic":
        a = Kinv @ y
        W = np.outer(a, a) - Kinv
        M = W * K
        dnoise = -0.5 * float(np.trace(W)) * sn2
        dfunc = -0.5 * float(np.sum(M))
        rs = M.sum(axis=1)
        if ard_mode != "intended":
            raise NotImplementedError("analytic ARD gradient is defined for ard_mode='intended'")
        dl = -(1.0 / ell ** 2) * ((rs[:, None] * X ** 2).sum(axis=0) - (X * (M @ X)).sum(axis=0))
        if noise_corr == "noiseCorrLearned":
            dcorr = -0.5 * float(np.sum(np.diag(W) * cens)) * math.exp(float(np.ravel(lnc)[0]))
    elif cov_form == "Matern":
        # derived (no reference branch): W-form with Kd = -sf2 f'(r)/r in the length-scale term
        a = Kinv @ y
        W = np.outer(a, a) - Kinv
        l = float(ell[0])
        rs_ = rdist(X / l, X / l)
        Kd = matern_deriv_factor(rs_, sf2, str(int(extra_param)))
        dnoise = -0.5 * float(np.trace(W)) * sn2
        dfunc = -0.5 * float(np.sum(W * K))
        dl = np.array([-0.5 * float(np.sum(W * Kd * rs_ ** 2))])
        if noise_corr == "noiseCorrLearned":
            dcorr = -0.5 * float(np.sum(np.diag(W) * cens)) * math.exp(float(np.ravel(lnc)[0]))
    elif cov_form == "ARD" and grad_mode == "as_written":
        if noise_corr == "noiseCorrLearned":
            raise RuntimeError("dvarNoiseSqCorrection undefined in the ARD branch (:58-69,89)")
        dnoise = half_trace_minus_quad(np.eye(n))                            # :58 (no exp factor)
        dfunc = half_trace_minus_quad(np.exp(-(r ** 2) / 2.0))               # :59
        dl = np.zeros(d)
        for k in range(d):                                                   # :63-68
            dsig = sf2 * np.exp((0.25 * r) / (ell[k] ** 3)) + sn2 * np.eye(n)
            dl[k] = half_trace_minus_quad(dsig)
    else:
        raise NotImplementedError((cov_form, grad_mode))

    parts = [np.array([dnoise, dfunc]), np.asarray(dl).ravel()]
    if mean_form != "Zero":                                                  # :54-56, :88
        ac = Kinv @ (y - m)
        parts.append(-(X.T @ ac))
        parts.append(np.array([-float(ac.sum())]))
    if noise_corr == "noiseCorrLearned":                                     # :89
        parts.append(np.array([dcorr]))
    return np.concatenate(parts)


# --------------------------------------------------------------------------------------
# GPRegression.R:60-71 (everything after optim)
# --------------------------------------------------------------------------------------
def regress_finalize(par, X, y, events, cov_form="SqExp", mean_form="Zero", noise_corr=False,
                     log_hyp_noise_corr=None, ard_mode="intended", as_written_ops=True, extra_param=None):
    """Recompute K, L, alpha and the (positive-sign) log marginal likelihood at ``par``
    (GPRegression.R:32-71).  Returns the fields of the reference's list (:73-82)."""
    X = np.asarray(X, dtype=np.float64)
    y = np.asarray(y, dtype=np.float64).ravel()
    n, d = X.shape
    log_sn2, log_sf2, log_ell, mean_hyp, lnc = unpack_log_hyp(
        par, d, cov_form, mean_form, noise_corr, log_hyp_noise_corr)
    sn2, sf2, ell = np.exp(log_sn2), np.exp(log_sf2), np.exp(log_ell)        # :39-52
    m = mean_func(mean_hyp, X, mean_form)                                    # :60
    K = cov_func(X, X, sf2, ell, cov_form, ard_mode, extra_param)            # :61
    # the reference re-logs the exponentiated values (:73): log(exp(x)) is not x bit-for-bit
    log_hyp_chosen = {"noise": float(np.log(sn2)), "func": float(np.log(sf2)),
                      "length": np.log(ell), "mean": np.array(mean_hyp, dtype=np.float64)}
    Ky = K + np.diag(_noise_diag(n, float(log_sn2), events, noise_corr, lnc))   # :62-68
    L = _chol_lower(Ky)                                                      # :65-69
    if as_written_ops:
        alpha = _solve_as_written(L, y - m)                                  # :70
    else:
        alpha = sla.solve_triangular(L, sla.solve_triangular(L, y - m, lower=True),
                                     lower=True, trans="T")
    log_mar_lik = -0.5 * float((y - m) @ alpha) - float(np.sum(np.log(np.diag(L)))) \
        - (n / 2.0) * math.log(2.0 * math.pi)                                # :71
    out = {"meanTraining": m, "K": K, "L": L, "alpha": alpha,
           "logHypChosen": log_hyp_chosen, "logMarLik": log_mar_lik}
    if _has_corr(noise_corr):                                                # :76-81
        lnc = np.asarray(lnc, dtype=np.float64).ravel()
        if lnc.size and lnc[0] == -np.inf:      # length-c condition: R < 4.2 tests the first element only
            lnc = np.array([-1e10])
        out["logHypNoiseCorrection"] = lnc
    return out


# --------------------------------------------------------------------------------------
# GPPredict.R
# --------------------------------------------------------------------------------------
def predict(reg, X, y, events, Xstar, cov_form="SqExp", mean_form="Zero", noise_corr=False,
            log_hyp_noise_corr=None, ard_mode="intended", as_written_ops=True, extra_param=None):
    """GPPredict.R:24-56. ``reg`` is the dict from :func:`regress_finalize` (uses ``L``,
    ``alpha``, ``logHypChosen``).  For GPS2/GPS3 L and alpha are rebuilt from the passed
    noise correction and the *current* targets ``y`` (:36-46)."""
    X = np.asarray(X, dtype=np.float64)
    Xstar = np.asarray(Xstar, dtype=np.float64)
    y = np.asarray(y, dtype=np.float64).ravel()
    n, d = X.shape
    lh = reg["logHypChosen"]
    sn2 = math.exp(lh["noise"])                                              # :24
    sf2 = math.exp(lh["func"])                                               # :27
    ell = np.exp(np.asarray(lh["length"], dtype=np.float64))                 # :28
    mean_hyp = lh["mean"]
    L, alpha = reg["L"], reg["alpha"]
    if _has_corr(noise_corr):                                                # :36-46
        m = mean_func(mean_hyp, X, mean_form)
        K = cov_func(X, X, sf2, ell, cov_form, ard_mode, extra_param) + np.diag(np.full(n, sn2))
        corr = np.zeros(n)
        corr[np.asarray(events).ravel() == 0] = np.exp(
            np.asarray(log_hyp_noise_corr, dtype=np.float64).ravel())
        L = _chol_lower(K + np.diag(corr))
        alpha = _solve_as_written(L, y - m) if as_written_ops else \
            sla.solve_triangular(L, sla.solve_triangular(L, y - m, lower=True), lower=True, trans="T")
    mstar = mean_func(mean_hyp, Xstar, mean_form)                            # :49
    kstar = cov_func(X, Xstar, sf2, ell, cov_form, ard_mode, extra_param)    # :50
    if as_written_ops:
        v = np.linalg.solve(L, kstar)                                        # :51 (general solve)
        kss = cov_func(Xstar, Xstar, sf2, ell, cov_form, ard_mode, extra_param)   # :55 full m x m
        var_function = np.diag(kss - v.T @ v)
    else:
        v = sla.solve_triangular(L, kstar, lower=True)
        var_function = sf2 - np.einsum("ij,ij->j", v, v)
    func_mean_pred = mstar + kstar.T @ alpha                                 # :54
    var_data = var_function + sn2                                            # :56
    return {"funcMeanPred": func_mean_pred, "varFunction": var_function, "varData": var_data}


# --------------------------------------------------------------------------------------
# stats::pnorm / stats::dnorm (third-party: R's nmath, version unpinned — README.md:24-43).
# Restated from the published algorithm: W. J. Cody (1969), "Rational Chebyshev
# approximations for the error function", Math. Comp. 23, as used by R's pnorm_both();
# call sites AdjustTrainingSurvivalMeanVariance.R:21-22.
# --------------------------------------------------------------------------------------
_CODY_A = (2.2352520354606839287, 161.02823106855587881, 1067.6894854603709582,
           18154.981253343561249, 0.065682337918207449113)
_CODY_B = (47.20258190468824187, 976.09855173777669322, 10260.932208618978205,
           45507.789335026729956)
_CODY_C = (0.39894151208813466764, 8.8831497943883759412, 93.506656132177855979,
           597.27027639480026226, 2494.5375852903726711, 6848.1904505362823326,
           11602.651437647350124, 9842.7148383839780218, 1.0765576773720192317e-8)
_CODY_D = (22.266688044328115691, 235.38790178262499861, 1519.377599407554805,
           6485.558298266760755, 18615.571640885098091, 34900.952721145977266,
           38912.003286093271411, 19685.429676859990727)
_CODY_P = (0.21589853405795699, 0.1274011611602473639, 0.022235277870649807,
           0.001421619193227893466, 2.9112874951168792e-5, 0.02307344176494017303)
_CODY_Q = (1.28426009614491121, 0.468238212480865118, 0.0659881378689285515,
           0.00378239633202758244, 7.29751555083966205e-5)
M_1_SQRT_2PI = 0.398942280401432677939946059934
M_SQRT_32 = 5.656854249492380195206754896838


def pnorm_std(x: float) -> float:
    """Lower-tail standard normal CDF, Cody's algorithm, scalar FP64."""
    if math.isnan(x):
        return x
    y = abs(x)
    if y <= 0.67448975:
        if y > 2.220446049250313e-16 * 0.5:     # DBL_EPSILON * 0.5
            xsq = x * x
            xnum = _CODY_A[4] * xsq
            xden = xsq
            for i in range(3):
                xnum = (xnum + _CODY_A[i]) * xsq
                xden = (xden + _CODY_B[i]) * xsq
        else:
            xnum = xden = 0.0
        temp = x * (xnum + _CODY_A[3]) / (xden + _CODY_B[3])
        return 0.5 + temp
    if y <= M_SQRT_32:
        xnum = _CODY_C[8] * y
        xden = y
        for i in range(7):
            xnum = (xnum + _CODY_C[i]) * y
            xden = (xden + _CODY_D[i]) * y
        temp = (xnum + _CODY_C[7]) / (xden + _CODY_D[7])
        xsq = math.trunc(y * 16.0) / 16.0
        dele = (y - xsq) * (y + xsq)
        cum = math.exp(-xsq * xsq * 0.5) * math.exp(-dele * 0.5) * temp
        ccum = 1.0 - cum
        return ccum if x > 0.0 else cum
    if -37.5193 < x < 8.2924:
        xsq = 1.0 / (x * x)
        xnum = _CODY_P[5] * xsq
        xden = xsq
        for i in range(4):
            xnum = (xnum + _CODY_P[i]) * xsq
            xden = (xden + _CODY_Q[i]) * xsq
        temp = xsq * (xnum + _CODY_P[4]) / (xden + _CODY_Q[4])
        temp = (M_1_SQRT_2PI - temp) / y
        xsq = math.trunc(x * 16.0) / 16.0
        dele = (x - xsq) * (x + xsq)
        cum = math.exp(-xsq * xsq * 0.5) * math.exp(-dele * 0.5) * temp
        ccum = 1.0 - cum
        return ccum if x > 0.0 else cum
    return 1.0 if x > 0.0 else 0.0


def dnorm_std(x: float) -> float:
    """Standard normal density as R's dnorm4 computes it (mean 0, sd 1)."""
    if math.isnan(x):
        return x
    x = abs(x)
    if x >= 2.0 * math.sqrt(1.7976931348623157e308):
        return 0.0
    if x < 5.0:
        return M_1_SQRT_2PI * math.exp(-0.5 * x * x)
    if x > math.sqrt(-2.0 * math.log(2.0) * (-1021 + 1 - 53)):
        return 0.0
    x1 = math.ldexp(float(round(math.ldexp(x, 16))), -16)   # R_forceint = nearbyint (ties-even)
    x2 = x - x1
    return M_1_SQRT_2PI * (math.exp(-0.5 * x1 * x1) * math.exp((-0.5 * x2 - x1) * x2))


# --------------------------------------------------------------------------------------
# AdjustTrainingSurvivalMeanVariance.R
# --------------------------------------------------------------------------------------
def adjust_one(a: float, mu: float, sigma: float):
    """AdjustTrainingSurvivalMeanVariance.R:10,21-31 for one censored sample.

    ``sigma`` is whatever the caller passes — ApplyGP.R:173,190 passes the predictive
    *variance* — and is used as a standard deviation (reference quirk, kept).
    ``1 - pnorm(alpha)`` is computed the lossy way on purpose.
    Lines 13-18 of the reference are dead code and are not restated.
    """
    alpha = (a - mu) / sigma                                                 # :10
    phi = dnorm_std(alpha)                                                   # :21
    Phi = pnorm_std(alpha)                                                   # :22
    one_minus = 1.0 - Phi
    if one_minus < 1e-12:                                                    # :28-31
        return a, 0.0
    lam = phi / one_minus
    return mu + sigma * lam, sigma ** 2 * (1.0 - lam * (lam - alpha))        # :25-26


def adjust(a, mu, sigma):
    """Vector form of :func:`adjust_one` (the ``sapply`` of ApplyGP.R:190)."""
    a, mu, sigma = (np.asarray(v, dtype=np.float64).ravel() for v in (a, mu, sigma))
    out = [adjust_one(float(ai), float(mi), float(si)) for ai, mi, si in zip(a, mu, sigma)]
    return np.array([o[0] for o in out]), np.array([o[1] for o in out])


# --------------------------------------------------------------------------------------
# stats::optim (third-party C inside R, version unpinned; call site GPRegression.R:28-31).
# Restated from the published algorithms R implements: Nash (1990), "Compact Numerical
# Methods for Computers", Algorithm 19 (Nelder-Mead, R's nmmin) and Algorithm 21
# (variable metric / BFGS, R's vmmin), with R's defaults alpha=1, beta=0.5, gamma=2,
# abstol=-Inf, reltol=sqrt(.Machine$double.eps).
# --------------------------------------------------------------------------------------
_BIG = 1.0e35
RELTOL = math.sqrt(np.finfo(np.float64).eps)


def nelder_mead(fn, par, maxit=500, abstol=-np.inf, intol=RELTOL, alpha=1.0, bet=0.5, gamm=2.0):
    """R's ``optim(method='Nelder-Mead')``: returns dict(par, value, counts, convergence).
    ``maxit`` bounds the number of *function evaluations* exactly as nmmin does."""
    bvec = np.array(par, dtype=np.float64).ravel()
    n = bvec.size
    if maxit <= 0:
        return {"par": bvec, "value": float(fn(bvec)), "counts": 0, "convergence": 0}
    f = float(fn(bvec))
    if not math.isfinite(f):
        raise RuntimeError("function cannot be evaluated at initial parameters")
    funcount = 1
    convtol = intol * (abs(f) + intol)
    n1 = n + 1
    C = n + 2
    P = np.zeros((n1, n + 2))
    P[n1 - 1, 0] = f
    P[:n, 0] = bvec
    L = 1
    size = 0.0
    step = 0.0
    for i in range(n):
        if 0.1 * abs(bvec[i]) > step:
            step = 0.1 * abs(bvec[i])
    if step == 0.0:
        step = 0.1
    for j in range(2, n1 + 1):
        P[:n, j - 1] = bvec
        trystep = step
        while P[j - 2, j - 1] == bvec[j - 2]:
            P[j - 2, j - 1] = bvec[j - 2] + trystep
            trystep *= 10.0
        size += trystep
    oldsize = size
    calcvert = True
    fail = 0
    while True:
        if calcvert:
            for j in range(n1):
                if j + 1 != L:
                    bvec = P[:n, j].copy()
                    f = float(fn(bvec))
                    if not math.isfinite(f):
                        f = _BIG
                    funcount += 1
                    P[n1 - 1, j] = f
            calcvert = False
        VL = P[n1 - 1, L - 1]
        VH = VL
        H = L
        for j in range(1, n1 + 1):
            if j != L:
                f = P[n1 - 1, j - 1]
                if f < VL:
                    L = j
                    VL = f
                if f > VH:
                    H = j
                    VH = f
        if VH <= VL + convtol or VL <= abstol:
            break
        for i in range(n):
            temp = -P[i, H - 1]
            for j in range(n1):
                temp += P[i, j]
            P[i, C - 1] = temp / n
        bvec = (1.0 + alpha) * P[:n, C - 1] - alpha * P[:n, H - 1]
        f = float(fn(bvec))
        if not math.isfinite(f):
            f = _BIG
        funcount += 1
        VR = f
        if VR < VL:
            P[n1 - 1, C - 1] = f
            for i in range(n):
                f = gamm * bvec[i] + (1.0 - gamm) * P[i, C - 1]
                P[i, C - 1] = bvec[i]
                bvec[i] = f
            f = float(fn(bvec))
            if not math.isfinite(f):
                f = _BIG
            funcount += 1
            if f < VR:
                P[:n, H - 1] = bvec
                P[n1 - 1, H - 1] = f
            else:
                P[:n, H - 1] = P[:n, C - 1]
                P[n1 - 1, H - 1] = VR
        else:
            if VR < VH:
                P[:n, H - 1] = bvec
                P[n1 - 1, H - 1] = VR
            bvec = (1.0 - bet) * P[:n, H - 1] + bet * P[:n, C - 1]
            f = float(fn(bvec))
            if not math.isfinite(f):
                f = _BIG
            funcount += 1
            if f < P[n1 - 1, H - 1]:
                P[:n, H - 1] = bvec
                P[n1 - 1, H - 1] = f
            elif VR >= VH:
                calcvert = True
                size = 0.0
                for j in range(n1):
                    if j + 1 != L:
                        for i in range(n):
                            P[i, j] = bet * (P[i, j] - P[i, L - 1]) + P[i, L - 1]
                            size += abs(P[i, j] - P[i, L - 1])
                if size < oldsize:
                    oldsize = size
                else:
                    fail = 10
                    break
        if not funcount <= maxit:
            break
    value = float(P[n1 - 1, L - 1])
    xout = P[:n, L - 1].copy()
    if funcount > maxit:
        fail = 1
    return {"par": xout, "value": value, "counts": funcount, "convergence": fail}


def bfgs(fn, gr, par, maxit=100, abstol=-np.inf, reltol=RELTOL):
    """R's ``optim(method='BFGS')`` (vmmin).  ``maxit`` bounds iterations (gradient calls)."""
    stepredn, acctol, reltest = 0.2, 1.0e-4, 10.0
    b = np.array(par, dtype=np.float64).ravel()
    n = b.size
    if maxit <= 0:
        return {"par": b, "value": float(fn(b)), "counts": (0, 0), "convergence": 0}
    f = float(fn(b))
    if not math.isfinite(f):
        raise RuntimeError("initial value in 'vmmin' is not finite")
    fmin = f
    funcount = gradcount = 1
    g = np.array(gr(b), dtype=np.float64).ravel()
    it = 1
    ilast = gradcount
    B = np.zeros((n, n))
    t = np.zeros(n)
    while True:
        if ilast == gradcount:
            B = np.eye(n)
        X = b.copy()
        c = g.copy()
        gradproj = 0.0
        for i in range(n):
            s = 0.0
            for j in range(i + 1):
                s -= B[i, j] * g[j]
            for j in range(i + 1, n):
                s -= B[j, i] * g[j]
            t[i] = s
            gradproj += s * g[i]
        if gradproj < 0.0:
            steplength = 1.0
            accpoint = False
            while True:
                count = 0
                for i in range(n):
                    b[i] = X[i] + steplength * t[i]
                    if reltest + X[i] == reltest + b[i]:
                        count += 1
                if count < n:
                    f = float(fn(b))
                    funcount += 1
                    accpoint = math.isfinite(f) and (f <= fmin + gradproj * steplength * acctol)
                    if not accpoint:
                        steplength *= stepredn
                if count == n or accpoint:
                    break
            enough = (f > abstol) and abs(f - fmin) > reltol * (abs(fmin) + reltol)
            if not enough:
                count = n
                fmin = f
            if count < n:
                fmin = f
                g = np.array(gr(b), dtype=np.float64).ravel()
                gradcount += 1
                it += 1
                D1 = 0.0
                for i in range(n):
                    t[i] = steplength * t[i]
                    c[i] = g[i] - c[i]
                    D1 += t[i] * c[i]
                if D1 > 0:
                    D2 = 0.0
                    for i in range(n):
                        s = 0.0
                        for j in range(i + 1):
                            s += B[i, j] * c[j]
                        for j in range(i + 1, n):
                            s += B[j, i] * c[j]
                        X[i] = s
                        D2 += s * c[i]
                    D2 = 1.0 + D2 / D1
                    for i in range(n):
                        for j in range(i + 1):
                            B[i, j] += (D2 * t[i] * t[j] - X[i] * t[j] - t[i] * X[j]) / D1
                else:
                    ilast = gradcount
            else:
                if ilast < gradcount:
                    count = 0
                    ilast = gradcount
        else:
            count = 0
            if ilast == gradcount:
                count = n
            else:
                ilast = gradcount
        if it >= maxit:
            break
        if gradcount - ilast > 2 * n:
            ilast = gradcount
        if count == n and ilast == gradcount:
            break
    return {"par": b, "value": float(fmin), "counts": (funcount, gradcount),
            "convergence": 0 if it < maxit else 1}


# --------------------------------------------------------------------------------------
# GPRegression.R (whole function)
# --------------------------------------------------------------------------------------
def flatten_log_hyp(log_hyp, d, mean_form, noise_corr, log_hyp_noise_corr):
    """GPRegression.R:22-25 — unlist(logHyp), drop the mean block for a Zero mean, append
    the GPS2 correction."""
    vec = np.concatenate([np.atleast_1d(np.asarray(log_hyp["noise"], dtype=np.float64)).ravel(),
                          np.atleast_1d(np.asarray(log_hyp["func"], dtype=np.float64)).ravel(),
                          np.atleast_1d(np.asarray(log_hyp["length"], dtype=np.float64)).ravel(),
                          np.atleast_1d(np.asarray(log_hyp["mean"], dtype=np.float64)).ravel()])
    if mean_form == "Zero":
        vec = vec[:vec.size - (d + 1)]
    if noise_corr == "noiseCorrLearned":
        vec = np.concatenate([vec, np.atleast_1d(np.asarray(log_hyp_noise_corr, dtype=np.float64)).ravel()])
    return vec


def gp_regression(log_hyp, params, data, log_hyp_noise_corr, ard_mode="intended",
                  as_written_ops=True):
    """GPRegression.R:1-83.  ``params`` mirrors parameterStructure (keys noiseCorr, modelType,
    optimType, maxit, maxitSurvival, meanFuncForm, covFuncForm); ``data`` mirrors
    trainingTestStructure (trainingData, trainingTargets, events)."""
    X = np.asarray(data["trainingData"], dtype=np.float64)
    y = np.asarray(data["trainingTargets"], dtype=np.float64).ravel()
    events = np.asarray(data.get("events", np.ones(X.shape[0]))).ravel()
    d = X.shape[1]
    noise_corr = params["noiseCorr"]
    maxit = params["maxitSurvival"] if params["modelType"] == "survival" else params["maxit"]   # :16
    mean_form, cov_form = params["meanFuncForm"], params["covFuncForm"]
    vec = flatten_log_hyp(log_hyp, d, mean_form, noise_corr, log_hyp_noise_corr)                 # :22-25
    extra = params.get("extraParam", params.get("maternParam"))
    kw = dict(cov_form=cov_form, mean_form=mean_form, noise_corr=noise_corr,
              log_hyp_noise_corr=log_hyp_noise_corr, ard_mode=ard_mode)
    fn = lambda p: nlml(p, X, y, events, as_written_ops=as_written_ops, extra_param=extra, **kw)
    gr = lambda p: nlml_grad(p, X, y, events, extra_param=extra, **kw)
    method = params["optimType"]
    if method == "Nelder-Mead":                                                                  # :28-31
        opt = nelder_mead(fn, vec, maxit=maxit)
    elif method == "BFGS":
        opt = bfgs(fn, gr, vec, maxit=maxit)
    else:
        raise NotImplementedError(method)
    vec = opt["par"]
    lnc = log_hyp_noise_corr
    out = regress_finalize(vec, X, y, events, cov_form, mean_form, noise_corr, lnc, ard_mode,
                           as_written_ops, extra)                                                # :32-71
    if noise_corr == "noiseCorrLearned":                                                         # :35-38
        out["logHypNoiseCorrection"] = vec[-1:]
    out["logHyp"] = log_hyp
    out["objective"] = opt["value"]                                                              # :33
    out["optim"] = opt
    return out


# --------------------------------------------------------------------------------------
# ApplyGP.R:77-282 (control flow only; plotting / saving / metrics are out of scope)
# --------------------------------------------------------------------------------------
def apply_gp(tts, params, data_source="Generate", ard_mode="intended", as_written_ops=True):
    """ApplyGP.R:77-282.  ``tts`` mirrors trainingTestStructure (trainingData,
    trainingTargets, events, testData, testTargets); ``params`` mirrors parameterStructure
    (adds tolerance, maxCount, logHypStart; burnIn must be FALSE — SURVEY.md F5/F6).
    Returns the numeric fields of the reference's return list (:440-467) that do not
    depend on the out-of-scope metrics packages."""
    X = np.asarray(tts["trainingData"], dtype=np.float64)
    Xtest = np.asarray(tts["testData"], dtype=np.float64)
    events = np.asarray(tts["events"]).ravel()
    model_type = params["modelType"]
    noise_corr = params["noiseCorr"]
    kw = dict(cov_form=params["covFuncForm"], mean_form=params["meanFuncForm"], ard_mode=ard_mode,
              as_written_ops=as_written_ops, extra_param=params.get("extraParam", params.get("maternParam")))
    y = np.log(np.asarray(tts["trainingTargets"], dtype=np.float64).ravel())        # :79 / :94
    if model_type == "survival" and data_source in ("CuratedOvarian", "TCGA2STAT", "TCGASynapse",
                                                    "TCGAYuanEtAl"):
        y = y - np.median(y)                                                        # :86-92
    log_hyp = params["logHypStart"]                                                 # :130,136 (burnIn FALSE)
    if params.get("burnIn", False):
        raise NotImplementedError("burnIn is broken as shipped (SURVEY.md F6)")
    out = {}
    if model_type == "survival":
        cens = events == 0
        a_start = y[cens].copy()                                                    # :108
        n_cens = int(cens.sum())
        y_learned = y.copy()                                                        # :145
        if noise_corr == "noiseCorrLearned":                                        # :146-152
            lnc = np.atleast_1d(np.asarray(log_hyp["noise"], dtype=np.float64))
        elif noise_corr == "noiseCorrVec":
            lnc = np.full(n_cens, float(np.ravel(log_hyp["noise"])[0]))
        else:
            lnc = None
        count, change, objective = 0, 1.0, 1.0                                      # :140-143
        objective_table = [0.0]
        hyp_table = []
        while (change > params["tolerance"] and count < params["maxCount"]) or count <= 5:   # :162
            data = {"trainingData": X, "trainingTargets": y_learned, "events": events}
            reg = gp_regression(log_hyp, params, data, lnc, ard_mode, as_written_ops)         # :170
            pred = predict(reg, X, y_learned, events, X[cens], noise_corr=noise_corr,
                           log_hyp_noise_corr=reg.get("logHypNoiseCorrection"), **kw)         # :171
            log_hyp = reg["logHypChosen"]                                                     # :185
            hyp_table.append(log_hyp)
            objective_new = reg["objective"]                                                  # :187
            objective_table.append(objective_new)
            upd_mean, upd_var = adjust(a_start, pred["funcMeanPred"], pred["varData"])        # :190
            change = abs(objective - objective_new)                                           # :195
            objective = objective_new
            count += 1
            y_learned = y_learned.copy()
            y_learned[cens] = upd_mean                                                        # :201
            if noise_corr == "noiseCorrVec":                                                  # :208-210
                with np.errstate(divide="ignore"):
                    lnc = np.log(upd_var)
                lnc[np.isinf(lnc)] = -1e10
            elif noise_corr == "noiseCorrLearned":                                            # :211-212
                lnc = reg["logHypNoiseCorrection"]
        out["convergence"] = 0 if count < params["maxCount"] else 1                           # :216-221
        out["count"] = count
        out["objectiveTable"] = np.array(objective_table)
        out["trainingTargetsLearned"] = y_learned
        y_final = y_learned
    else:                                                                           # :224-232
        data = {"trainingData": X, "trainingTargets": y, "events": events}
        params = dict(params, noiseCorr=False)
        noise_corr = False
        reg = gp_regression(log_hyp, params, data, None, ard_mode, as_written_ops)
        lnc = None
        y_final = y
    pred = predict(reg, X, y_final, events, Xtest, noise_corr=noise_corr,
                   log_hyp_noise_corr=lnc, **kw)                                    # :243 / :245
    pred_train = predict(reg, X, y_final, events, X, noise_corr=noise_corr,
                         log_hyp_noise_corr=lnc, **kw)                              # :272 / :280
    out.update({"logHypChosen": reg["logHypChosen"], "funcMeanPred": pred["funcMeanPred"],
                "varFunction": pred["varFunction"], "varData": pred["varData"],
                "trainingSetPredictions": pred_train["funcMeanPred"],
                "logMarLik": reg["logMarLik"], "objective": reg["objective"]})
    return out


# --------------------------------------------------------------------------------------
# CalculateMetrics.R:18,21 (the c-index comes from survcomp, which is not under /root/reference)
# --------------------------------------------------------------------------------------
def concordance_index(x, surv_time, surv_event):
    """survcomp::concordance.index(x, surv.time, surv.event) with its defaults (method='noether', outx=TRUE), restated
    from its published definition (third-party, version unpinned; call site CalculateMetrics.R:18): a pair is usable
    when the shorter observed time is an event; it is concordant when the higher score x (a RISK score in survcomp's
    convention) has the shorter time; pairs tied in x are excluded.  Returns (c.index, concordant, discordant, tied)."""
    x, t, e = (np.asarray(v, dtype=np.float64).ravel() for v in (x, surv_time, surv_event))
    conc = disc = tied = 0
    for i in range(x.size):
        if e[i] != 1.0:
            continue
        later = t[i] < t
        conc += int(np.sum(later & (x[i] > x)))
        disc += int(np.sum(later & (x[i] < x)))
        tied += int(np.sum(later & (x[i] == x)))
    return (conc / (conc + disc) if conc + disc > 0 else float("nan")), conc, disc, tied


def rmse(pred, targets):
    """CalculateMetrics.R:21."""
    pred, targets = np.asarray(pred, dtype=np.float64).ravel(), np.asarray(targets, dtype=np.float64).ravel()
    return float(np.sqrt(np.sum((targets - pred) ** 2) / pred.size))
