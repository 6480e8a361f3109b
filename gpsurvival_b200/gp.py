"""Host-side mirror of the reference's R interface for the GP hot path, over libgpsurv.so.

R is not available in the build image, so this module plays the role of the
drop-in R files in ``r/``: same function names, argument order, option strings, return
fields and error behaviour as

    CovFunc.R, MeanFunc.R, ForOptimisationLog.R, OptimisationGradientLog.R, GPRegression.R,
    GPPredict.R, AdjustTrainingSurvivalMeanVariance.R and ApplyGP.R:77-282

with every formula executed by the CUDA library through the C ABI (include/gpsurv.h).
Nothing here computes covariances, factorisations, solves, gradients or moments on the
CPU, and nothing imports ``oracle``.  If libgpsurv.so or a B200 is missing, calls raise.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _lib
from ._lib import COV, MEAN, NOISE_CORR, ARD_MODE, GpsError, NotPositiveDefinite, check, cov_code, dptr, f64
from . import optim as _optim


# =========================================================================================
# Thin object wrapper over the opaque handle
# =========================================================================================
class Fit:
    """One ``gps_handle``: a fit (or a batch of identically shaped fits) resident on one GPU."""

    def __init__(self, device: int = 0, batch: int = 1):
        self.lib = _lib.load()
        self.h = C.c_void_p()
        self.batch = batch
        self.device = device
        check(self.lib.gps_create_batch(device, batch, C.byref(self.h)))
        self.n = self.d = 0
        self.opts = ("SqExp", "Zero", False, "intended")

    def close(self):
        if getattr(self, "h", None) is not None and self.h:
            self.lib.gps_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _ck(self, rc):
        check(rc, self.h)

    # -- configuration -------------------------------------------------------------------
    def set_options(self, cov_form="SqExp", mean_form="Zero", noise_corr=False, ard_mode="intended", extra_param=None):
        code = cov_code(cov_form, extra_param)
        self._ck(self.lib.gps_set_options(self.h, code, MEAN[mean_form], NOISE_CORR[noise_corr], ARD_MODE[ard_mode]))
        self.opts = (cov_form, mean_form, noise_corr, ard_mode) if cov_form != "Matern" else \
            (cov_form, mean_form, noise_corr, ard_mode, int(extra_param))

    def set_data(self, X, y, events=None, colmajor=False):
        """gps_set_data.  X: n x d (one fit) or [batch, n, d]; with colmajor=True a batch is given the way the ABI
        (and R) stores it, [batch, d, n] C-contiguous = each fit column-major, and is passed through untouched."""
        if np.ndim(X) == 2:
            X = f64(X)
            n, d = X.shape
        elif colmajor:
            X = np.ascontiguousarray(np.asarray(X, dtype=np.float64))
            _, d, n = X.shape
        else:   # [batch, n, d] -> each problem column-major
            X = np.asarray(X, dtype=np.float64)
            _, n, d = X.shape
            X = np.ascontiguousarray(np.transpose(X, (0, 2, 1)))
        y = np.ascontiguousarray(np.asarray(y, dtype=np.float64)).reshape(self.batch, n)
        ev = None if events is None else np.ascontiguousarray(np.asarray(events, dtype=np.float64)).reshape(self.batch, n)
        self._ck(self.lib.gps_set_data(self.h, dptr(X), n, d, dptr(y), dptr(ev)))
        self.n, self.d = n, d

    def set_data_indexed(self, Xpool, ypool, eventspool, rows):
        """gps_set_data_indexed: the fits of this batch as row-index views (rows: [batch, n], 0-based) of one pooled data
        set, uploaded once."""
        Xp = f64(Xpool)
        npool, d = Xp.shape
        rows = np.ascontiguousarray(np.asarray(rows, dtype=np.int32)).reshape(self.batch, -1)
        n = rows.shape[1]
        yp = np.ascontiguousarray(np.asarray(ypool, dtype=np.float64))
        evp = None if eventspool is None else np.ascontiguousarray(np.asarray(eventspool, dtype=np.float64))
        self._ck(self.lib.gps_set_data_indexed(self.h, dptr(Xp), npool, d, dptr(yp), dptr(evp),
                                               rows.ctypes.data_as(C.POINTER(C.c_int)), n))
        self.n, self.d = n, d

    def predict_rows(self, par, rows, reuse_model=False):
        """gps_predict_rows: GPPredict at training rows of each fit (rows: [batch, m], 0-based)."""
        p, ln = self._par(par)
        rows = np.ascontiguousarray(np.asarray(rows, dtype=np.int32)).reshape(self.batch, -1)
        m = rows.shape[1]
        mu, vf, vd = (np.empty((self.batch, m)) for _ in range(3))
        self._ck(self.lib.gps_predict_rows(self.h, dptr(p), ln, 1 if reuse_model else 0, rows.ctypes.data_as(C.POINTER(C.c_int)),
                                           m, dptr(mu), dptr(vf), dptr(vd)))
        return (mu, vf, vd) if self.batch > 1 else (mu[0], vf[0], vd[0])

    def metrics(self, pred, time, event):
        """gps_metrics: CalculateMetrics.R's c.index (survcomp semantics) and rmse for `batch` sets of m predictions."""
        P = np.ascontiguousarray(np.asarray(pred, dtype=np.float64)).reshape(self.batch, -1)
        m = P.shape[1]
        T = np.ascontiguousarray(np.asarray(time, dtype=np.float64)).reshape(self.batch, m)
        E = np.ascontiguousarray(np.asarray(event, dtype=np.float64)).reshape(self.batch, m)
        ci, rm = np.empty(self.batch), np.empty(self.batch)
        pairs = (C.c_longlong * (3 * self.batch))()
        self._ck(self.lib.gps_metrics(self.h, dptr(P), dptr(T), dptr(E), m, dptr(ci), dptr(rm), pairs))
        return {"c.index": ci, "rmse": rm, "pairs": np.array(list(pairs)).reshape(self.batch, 3)}

    def synth_generate(self, seed, n, d, sn2=0.01, sf2=0.5, ell=1.1, jitter=1e-8, n_censor=0, cens_mean=0.0, cens_sd=50.0):
        """gps_synth_generate: MakeSyntheticData ('likeRealData') + CensorData ('NormalLoopSample') on the device, `batch` data sets."""
        B = self.batch
        X = np.empty((B, d, n)); y = np.empty((B, n)); ev = np.empty((B, n)); y0 = np.empty((B, n)); nc = (C.c_int * B)()
        self._ck(self.lib.gps_synth_generate(self.h, int(seed), n, d, float(sn2), float(sf2), float(ell), float(jitter), int(n_censor),
                                             float(cens_mean), float(cens_sd), dptr(X), dptr(y), dptr(ev), dptr(y0), nc))
        self.n, self.d = n, d
        self.opts = ("SqExp", "Zero", False, "intended")
        return {"data": np.transpose(X, (0, 2, 1)), "targets": y, "events": ev, "targetsPreCensoring": y0,
                "numCensored": np.array(list(nc))}

    def set_targets(self, y):
        y = np.ascontiguousarray(np.asarray(y, dtype=np.float64)).reshape(self.batch, self.n)
        self._ck(self.lib.gps_set_targets(self.h, dptr(y)))

    def set_noise_corr(self, log_sc2):
        v = np.ascontiguousarray(np.asarray(log_sc2, dtype=np.float64)).reshape(self.batch, -1)
        self._ck(self.lib.gps_set_noise_corr(self.h, dptr(v), v.shape[1]))

    # -- evaluations ---------------------------------------------------------------------
    def _par(self, par):
        p = np.ascontiguousarray(np.asarray(par, dtype=np.float64)).reshape(self.batch, -1)
        return p, p.shape[1]

    def nlml(self, par, not_pd="raise"):
        """``not_pd='inf'``: a fit whose Ky is not positive definite returns +Inf instead of raising
        (what an optimiser wants: nmmin / vmmin treat a non-finite value as a rejected point)."""
        p, ln = self._par(par)
        out = np.empty(self.batch)
        rc = self.lib.gps_nlml(self.h, dptr(p), ln, dptr(out))
        if not (rc == _lib.GPS_ERR_NOT_PD and not_pd == "inf"):
            self._ck(rc)
        return out if self.batch > 1 else float(out[0])

    def nlml_grad(self, par, not_pd="raise"):
        p, ln = self._par(par)
        out = np.empty(self.batch)
        g = np.empty((self.batch, ln))
        rc = self.lib.gps_nlml_grad(self.h, dptr(p), ln, dptr(out), dptr(g))
        if not (rc == _lib.GPS_ERR_NOT_PD and not_pd == "inf"):
            self._ck(rc)
        return (out, g) if self.batch > 1 else (float(out[0]), g[0])

    def set_near_pd(self, enable=True):
        """gps_set_near_pd: the device-side Matrix::nearPD repair of a non-positive-definite Ky (ForOptimisationLog.R:43-51)."""
        self._ck(self.lib.gps_set_near_pd(self.h, 1 if enable else 0))

    def near_pd_used(self):
        out = (C.c_int * self.batch)()
        self._ck(self.lib.gps_near_pd_used(self.h, out))
        return list(out)

    def pivots(self):
        out = (C.c_int * self.batch)()
        self._ck(self.lib.gps_batch_pivots(self.h, out))
        return list(out)

    def finalize(self, par, want_K=False, want_L=False):
        p, ln = self._par(par)
        B, n = self.batch, self.n
        K = np.empty((B, n, n)) if want_K else None
        L = np.empty((B, n, n)) if want_L else None
        alpha = np.empty((B, n))
        mean = np.empty((B, n))
        lml = np.empty(B)
        self._ck(self.lib.gps_regress_finalize(self.h, dptr(p), ln, dptr(K), dptr(L), dptr(alpha), dptr(mean),
                                               dptr(lml)))
        # device matrices are column-major: transpose the [B, col, row] view into [B, row, col]
        out = {"alpha": alpha, "meanTraining": mean, "logMarLik": lml}
        if want_K:
            out["K"] = np.transpose(K, (0, 2, 1))
        if want_L:
            out["L"] = np.transpose(L, (0, 2, 1))
        if B == 1:
            out = {k: (v[0] if isinstance(v, np.ndarray) and v.ndim >= 1 else v) for k, v in out.items()}
            out["logMarLik"] = float(lml[0])
        return out

    def predict(self, par, Xstar, reuse_model=False):
        p, ln = self._par(par)
        if np.ndim(Xstar) == 2:
            Xs = f64(Xstar)
            m = Xs.shape[0]
        else:
            Xs = np.asarray(Xstar, dtype=np.float64)
            m = Xs.shape[1]
            Xs = np.ascontiguousarray(np.transpose(Xs, (0, 2, 1)))
        mu, vf, vd = (np.empty((self.batch, m)) for _ in range(3))
        self._ck(self.lib.gps_predict(self.h, dptr(p), ln, 1 if reuse_model else 0, dptr(Xs), m, dptr(mu), dptr(vf),
                                      dptr(vd)))
        if self.batch == 1:
            return mu[0], vf[0], vd[0]
        return mu, vf, vd

    def adjust(self, a, mu, s):
        a, mu, s = (np.ascontiguousarray(np.asarray(v, dtype=np.float64).ravel()) for v in (a, mu, s))
        em, ev = np.empty_like(a), np.empty_like(a)
        self._ck(self.lib.gps_adjust(self.h, dptr(a), dptr(mu), dptr(s), a.size, dptr(em), dptr(ev)))
        return em, ev

    def cov(self, x1, x2, sf2, ell, cov_form="SqExp", ard_mode="intended", extra_param=None):
        x1 = f64(x1)
        n1, d = x1.shape
        ell = np.ascontiguousarray(np.atleast_1d(np.asarray(ell, dtype=np.float64)).ravel())
        if x2 is None:
            x2p, n2 = None, n1
        else:
            x2a = f64(x2)
            x2p, n2 = x2a, x2a.shape[0]
        K = np.empty((n1, n2), order="F")
        self._ck(self.lib.gps_cov(self.h, dptr(x1), n1, dptr(x2p), n2, d, float(sf2), dptr(ell), ell.size,
                                  cov_code(cov_form, extra_param), ARD_MODE[ard_mode], dptr(K)))
        return K

    def chol_matvec(self, par, v):
        """t(chol(K(par) + exp(par[0]) I)) %*% v for the handle's data (MakeSyntheticData.R:41,148)."""
        p, ln = self._par(par)
        v = np.ascontiguousarray(np.asarray(v, dtype=np.float64)).reshape(self.batch, self.n)
        out = np.empty_like(v)
        self._ck(self.lib.gps_chol_matvec(self.h, dptr(p), ln, dptr(v), dptr(out)))
        return out if self.batch > 1 else out[0]

    def fit_batch(self, X, y_log, events, par_start, Xtest, method="Nelder-Mead", maxit=10, tolerance=1e-3,
                  max_count=500, survival=True):
        """gps_fit_batch: ApplyGP.R:136-282 for every fit of this batched handle, entirely behind the C ABI
        (optimisers restated in C++).  Arrays carry the batch first: X [B, n, d], Xtest [B, m, d]."""
        X = np.asarray(X, dtype=np.float64)
        Xt = np.asarray(Xtest, dtype=np.float64)
        if X.ndim == 2:
            X, Xt = X[None], Xt[None]
        B, n, d = X.shape
        m = Xt.shape[1]
        Xc = np.ascontiguousarray(np.transpose(X, (0, 2, 1)))
        Xtc = np.ascontiguousarray(np.transpose(Xt, (0, 2, 1)))
        y = np.ascontiguousarray(np.asarray(y_log, dtype=np.float64)).reshape(B, n)
        ev = np.ascontiguousarray(np.asarray(events, dtype=np.float64)).reshape(B, n)
        ps = np.ascontiguousarray(np.asarray(par_start, dtype=np.float64)).reshape(B, -1)
        ln = ps.shape[1]
        par = np.empty((B, ln)); obj = np.empty(B); lml = np.zeros(B); rounds = (C.c_int * B)()
        learned = np.empty((B, n)); tm = np.empty((B, m)); tvf = np.empty((B, m)); tvd = np.empty((B, m)); trm = np.empty((B, n))
        ticks = C.c_longlong()
        self._ck(self.lib.gps_fit_batch(self.h, dptr(Xc), n, d, dptr(y), dptr(ev), dptr(ps), ln,
                                        {"Nelder-Mead": 0, "BFGS": 1, "L-BFGS": 2}[method], int(maxit), float(tolerance), int(max_count),
                                        1 if survival else 0, dptr(Xtc), m, dptr(par), dptr(obj), dptr(lml), rounds,
                                        dptr(learned), dptr(tm), dptr(tvf), dptr(tvd), dptr(trm), C.byref(ticks)))
        self.n, self.d = n, d
        st = (C.c_int * B)()
        self._ck(self.lib.gps_fit_batch_status(self.h, st))
        return {"par": par, "objective": obj, "logMarLik": lml, "count": np.array(list(rounds)), "trainingTargetsLearned": learned,
                "funcMeanPred": tm, "varFunction": tvf, "varData": tvd, "trainingSetPredictions": trm, "ticks": int(ticks.value),
                "status": np.array(list(st))}   # per fit: 0 = ok, 3 = lost (Ky not positive definite at the point optim returned)

    # -- measurement helpers -------------------------------------------------------------
    def launch_count(self):
        return int(self.lib.gps_launch_count(self.h))

    def last_device_ms(self):
        return float(self.lib.gps_last_device_ms(self.h))

    def set_graphs(self, enable=True):
        self._ck(self.lib.gps_set_graphs(self.h, 1 if enable else 0))

    def time_eval(self, kind=1, reps=10):
        ms = C.c_double()
        self._ck(self.lib.gps_time_eval(self.h, kind, reps, C.byref(ms)))
        return ms.value

    def time_stages(self):
        ms = np.zeros(8)
        self._ck(self.lib.gps_time_stages(self.h, dptr(ms)))
        return ms

    def bench_gemm(self, M, N, K, reps=5):
        ms = C.c_double()
        self._ck(self.lib.gps_bench_gemm(self.h, M, N, K, reps, C.byref(ms)))
        return ms.value


# =========================================================================================
# Session: the device-resident state behind the stateless reference signatures
# =========================================================================================
class _Session:
    """Keeps one handle alive across the stateless calls the reference makes (optim calls
    ForOptimisationLog / OptimisationGradientLog with the same trainingData thousands of
    times).  X is re-uploaded only when it changes; targets / noise correction when they do."""

    def __init__(self):
        self.fit = None
        self.device = 0
        self.X = None
        self.y = None
        self.events = None
        self.lnc = None
        self.ard_mode = "intended"

    def reset(self):
        if self.fit is not None:
            self.fit.close()
        self.__init__()

    def bind(self, X, y, events, cov_form, mean_form, noise_corr, lnc, extra_param=None):
        if self.fit is None:
            self.fit = Fit(self.device)
        fit = self.fit
        nc = noise_corr if noise_corr in ("noiseCorrLearned", "noiseCorrVec") else False
        opts = (cov_form, mean_form, nc, self.ard_mode) if cov_form != "Matern" else \
            (cov_form, mean_form, nc, self.ard_mode, int(extra_param) if extra_param is not None else None)
        if fit.opts != opts:
            fit.set_options(cov_form, mean_form, nc, self.ard_mode, extra_param)
        X = np.asarray(X, dtype=np.float64)
        y = np.asarray(y, dtype=np.float64).ravel()
        events = np.ones(X.shape[0]) if events is None else np.asarray(events, dtype=np.float64).ravel()
        same_X = self.X is not None and (X is self.X or (X.shape == self.X.shape and np.array_equal(X, self.X)))
        same_ev = self.events is not None and np.array_equal(events, self.events)
        if not (same_X and same_ev):
            fit.set_data(X, y, events)
            self.X, self.y, self.events = X, y.copy(), events.copy()
            self.lnc = None
        elif not np.array_equal(y, self.y):
            fit.set_targets(y)
            self.y = y.copy()
        if nc == "noiseCorrVec":
            v = np.asarray(lnc, dtype=np.float64).ravel()
            ncens = int(np.sum(events == 0))
            if v.size == 1 and ncens != 1:     # R recycles a scalar over the censored rows
                v = np.full(ncens, v[0])
            if self.lnc is None or not np.array_equal(v, self.lnc):
                fit.set_noise_corr(v)
                self.lnc = v.copy()
        return fit


SESSION = _Session()


def set_device(device: int):
    SESSION.reset()
    SESSION.device = device


def set_ard_mode(mode: str):
    """'intended' (default; north_star) or 'as_written' (reproduces the reference's recycled
    ARD scaling bit for bit in structure — SURVEY.md F3)."""
    if mode not in ARD_MODE:
        raise ValueError(mode)
    SESSION.ard_mode = mode


# =========================================================================================
# Reference-named functions
# =========================================================================================
def CovFunc(x1, x2, x, y, extraParam, varFuncSqHyp, lengthHyp, covFuncForm):
    """CovFunc.R:1 — covariance between the rows of x1 and x2 (x, y unused).  'SqExp', 'ARD' and
    'Matern' (extraParam = maternParam in 1, 3, 5); 'RF', 'InformedARD' raise GPS_ERR_UNSUPPORTED."""
    cov_code(covFuncForm, extraParam)      # raises for RF / InformedARD / bad maternParam
    if SESSION.fit is None:
        SESSION.fit = Fit(SESSION.device)
    x1 = np.asarray(x1, dtype=np.float64)
    x2 = np.asarray(x2, dtype=np.float64)
    same = x1 is x2 or (x1.shape == x2.shape and np.array_equal(x1, x2))   # identical(x1, x2), CovFunc.R:24
    return np.asarray(SESSION.fit.cov(x1, None if same else x2, float(np.ravel(varFuncSqHyp)[0]),
                                      np.ravel(lengthHyp), covFuncForm, SESSION.ard_mode, extraParam))


def MeanFunc(meanHyp, x, meanFuncForm, n):
    """MeanFunc.R:1 — host GEMV exactly as in the reference (the device applies the same mean
    inside the fused evaluation; this standalone copy exists for API completeness)."""
    x = np.asarray(x, dtype=np.float64)
    d = x.shape[1]
    if meanFuncForm == "Zero":
        return np.zeros((n, 1))
    if meanFuncForm == "Linear":
        mh = np.asarray(meanHyp, dtype=np.float64).ravel()
        return (x @ mh[:d] + mh[d]).reshape(n, 1)
    raise ValueError(meanFuncForm)


def _reject_priors(imposePriors):
    if imposePriors:
        raise GpsError(_lib.GPS_ERR_UNSUPPORTED,
                       "imposePriors=TRUE cannot run in the reference as shipped (LogPriorX arity mismatch, "
                       "SURVEY.md F6); evaluate with imposePriors=FALSE")


def ForOptimisationLog(logHyp, trainingData, trainingTargets, trainingEvents, meanFuncForm, dimension, extraParam,
                       covFuncForm, nSamples, noiseCorr, logHypNoiseCorrection, imposePriors):
    """ForOptimisationLog.R:1 — negative log marginal likelihood, returned as a 1x1 matrix (:67).
    A non-positive-definite Ky raises NotPositiveDefinite (the nearPD fallback, :43-51, stays
    on the host side of the boundary)."""
    _reject_priors(imposePriors)
    fit = SESSION.bind(trainingData, trainingTargets, trainingEvents, covFuncForm, meanFuncForm, noiseCorr,
                       logHypNoiseCorrection, extraParam)
    return np.array([[fit.nlml(np.ravel(logHyp))]])


def OptimisationGradientLog(logHyp, trainingData, trainingTargets, trainingEvents, meanFuncForm, dimension,
                            extraParam, covFuncForm, nSamples, noiseCorr, logHypNoiseCorrection, imposePriors):
    """OptimisationGradientLog.R:1 — gradient as a k x 1 matrix (:93), same order as logHyp."""
    _reject_priors(imposePriors)
    fit = SESSION.bind(trainingData, trainingTargets, trainingEvents, covFuncForm, meanFuncForm, noiseCorr,
                       logHypNoiseCorrection, extraParam)
    _, g = fit.nlml_grad(np.ravel(logHyp))
    return g.reshape(-1, 1)


def _flatten(logHyp, d, meanFuncForm, noiseCorr, lnc):
    vec = np.concatenate([np.ravel(np.asarray(logHyp["noise"], dtype=np.float64)),
                          np.ravel(np.asarray(logHyp["func"], dtype=np.float64)),
                          np.ravel(np.asarray(logHyp["length"], dtype=np.float64)),
                          np.ravel(np.asarray(logHyp["mean"], dtype=np.float64))])      # GPRegression.R:22
    if meanFuncForm == "Zero":
        vec = vec[:vec.size - (d + 1)]                                                  # :23
    if noiseCorr == "noiseCorrLearned":
        vec = np.concatenate([vec, np.ravel(np.asarray(lnc, dtype=np.float64))])        # :24
    return vec


def _par_from_chosen(logHypChosen, d, meanFuncForm, noiseCorr, lnc):
    return _flatten(logHypChosen, d, meanFuncForm, noiseCorr, lnc)


def GPRegression(logHyp, parameterStructure, trainingTestStructure, logHypNoiseCorrection, want_matrices=False):
    """GPRegression.R:1 — optimise the hyper-parameters with stats::optim's algorithm, then
    rebuild the model at the optimum.  Returns the reference's list (:75-81).  ``K`` and ``L``
    (8n^2 bytes each) are only copied back when ``want_matrices`` is true; the factor stays
    resident on the GPU for GPPredict either way."""
    X = trainingTestStructure["trainingData"]
    y = np.asarray(trainingTestStructure["trainingTargets"], dtype=np.float64).ravel()
    events = trainingTestStructure.get("events")
    d = np.asarray(X).shape[1]
    ps = parameterStructure
    noiseCorr = ps["noiseCorr"]
    maxit = ps["maxitSurvival"] if ps["modelType"] == "survival" else ps["maxit"]        # :16
    meanFuncForm, covFuncForm = ps["meanFuncForm"], ps["covFuncForm"]
    _reject_priors(ps.get("imposePriors", False))
    vec = _flatten(logHyp, d, meanFuncForm, noiseCorr, logHypNoiseCorrection)
    extra = ps.get("extraParam", ps.get("maternParam"))                                  # :19
    fit = SESSION.bind(X, y, events, covFuncForm, meanFuncForm, noiseCorr, logHypNoiseCorrection, extra)
    # a non-positive-definite Ky inside optim is reported to the optimiser as +Inf (the reference
    # would try Matrix::nearPD there, ForOptimisationLog.R:43-51 — out of scope)
    opt = _optim.optim(vec, fn=lambda p: fit.nlml(p, not_pd="inf"), gr=lambda p: fit.nlml_grad(p, not_pd="inf")[1],
                       method=ps["optimType"], maxit=maxit)                              # :28-31
    vec = opt["par"]
    fin = fit.finalize(vec, want_K=want_matrices, want_L=want_matrices)                  # :60-71
    lnc = logHypNoiseCorrection
    body = vec
    if noiseCorr == "noiseCorrLearned":                                                  # :35-38
        lnc = vec[-1:]
        body = vec[:-1]
    p = d if covFuncForm == "ARD" else 1
    mean = body[body.size - d - 1:] if meanFuncForm == "Linear" else np.zeros(d + 1)     # :54-57
    # the reference stores log(exp(.)) of the optimum (:39-52,73)
    chosen = {"noise": float(np.log(np.exp(body[0]))), "func": float(np.log(np.exp(body[1]))),
              "length": np.log(np.exp(body[2:2 + p])), "mean": np.array(mean, dtype=np.float64)}
    out = {"logHyp": logHyp, "meanTraining": fin["meanTraining"].reshape(-1, 1), "alpha": fin["alpha"].reshape(-1, 1),
           "logHypChosen": chosen, "logMarLik": fin["logMarLik"], "objective": opt["value"],
           "K": fin.get("K"), "L": fin.get("L"), "optim": opt, "_par": vec.copy()}
    if noiseCorr in ("noiseCorrVec", "noiseCorrLearned"):                                # :76-81
        lnc = np.asarray(lnc, dtype=np.float64).ravel()
        if lnc.size and lnc[0] == -np.inf:
            lnc = np.array([-1e10])
        out["logHypNoiseCorrection"] = lnc
    return out


def GPPredict(regressionStructure, parameterStructure, trainingTestStructure, logHypNoiseCorrection):
    """GPPredict.R:1 — predictive mean / variances at testData.  For GPS2 / GPS3 the factor is
    rebuilt from the passed noise correction and the current trainingTargets (:36-46); for
    GPR / GPS1 the factor and alpha of ``regressionStructure`` (resident on the GPU) are used."""
    ps = parameterStructure
    meanFuncForm, covFuncForm, noiseCorr = ps["meanFuncForm"], ps["covFuncForm"], ps["noiseCorr"]
    X = trainingTestStructure["trainingData"]
    y = trainingTestStructure["trainingTargets"]
    events = trainingTestStructure.get("events")
    Xs = np.asarray(trainingTestStructure["testData"], dtype=np.float64)
    d = np.asarray(X).shape[1]
    corr = noiseCorr in ("noiseCorrVec", "noiseCorrLearned")
    par = _par_from_chosen(regressionStructure["logHypChosen"], d, meanFuncForm, noiseCorr, logHypNoiseCorrection)
    if corr:
        fit = SESSION.bind(X, y, events, covFuncForm, meanFuncForm, noiseCorr, logHypNoiseCorrection,
                           ps.get("extraParam", ps.get("maternParam")))
        mu, vf, vd = fit.predict(par, Xs, reuse_model=False)
    else:
        fit = SESSION.fit
        if fit is None:
            raise GpsError(_lib.GPS_ERR_STATE, "GPPredict: no resident model; call GPRegression first")
        mu, vf, vd = fit.predict(par, Xs, reuse_model=True)
    return {"funcMeanPred": mu.reshape(-1, 1), "varFunction": vf, "varData": vd}


def AdjustTrainingSurvivalMeanVariance(a, mu, sigma):
    """AdjustTrainingSurvivalMeanVariance.R:1 — accepts scalars (as the reference's sapply
    passes them) or vectors; returns list('mu'=, 'sigma'=)."""
    if SESSION.fit is None:
        SESSION.fit = Fit(SESSION.device)
    scalar = np.ndim(a) == 0
    em, ev = SESSION.fit.adjust(np.atleast_1d(a), np.atleast_1d(mu), np.atleast_1d(sigma))
    if scalar:
        return {"mu": float(em[0]), "sigma": float(ev[0])}
    return {"mu": em, "sigma": ev}


_REAL_SOURCES = ("CuratedOvarian", "TCGA2STAT", "TCGASynapse", "TCGAYuanEtAl")


def ApplyGP(trainingTestStructure, dataOptionsStructure, parameterStructure, plotSaveOptions=None):
    """ApplyGP.R:1, lines 77-282 — target transforms, the GPS outer loop, final predictions.
    Plotting, saving and the survcomp / survAUC metrics (:287-435) are out of scope."""
    tts = dict(trainingTestStructure)
    ps = dict(parameterStructure)
    modelType, noiseCorr = ps["modelType"], ps["noiseCorr"]
    if ps.get("burnIn", False):
        raise GpsError(_lib.GPS_ERR_UNSUPPORTED, "burnIn=TRUE is broken in the reference as shipped (SURVEY.md F6)")
    source = (dataOptionsStructure or {}).get("dataSource", "Generate")
    X = np.asarray(tts["trainingData"], dtype=np.float64)
    events = np.asarray(tts["events"], dtype=np.float64).ravel()
    y = np.log(np.asarray(tts["trainingTargets"], dtype=np.float64).ravel())            # :79 / :94
    if modelType == "survival" and source in _REAL_SOURCES:
        y = y - np.median(y)                                                            # :86-92
    logHyp = ps["logHypStart"]                                                          # :130,136
    out = {}
    if modelType == "survival":
        cens = events == 0
        a_start = y[cens].copy()                                                        # :108
        n_cens = int(cens.sum())
        learned = y.copy()                                                              # :145
        if noiseCorr == "noiseCorrLearned":                                             # :146-152
            lnc = np.atleast_1d(np.asarray(logHyp["noise"], dtype=np.float64))
        elif noiseCorr == "noiseCorrVec":
            lnc = np.full(n_cens, float(np.ravel(logHyp["noise"])[0]))
        else:
            lnc = None
        count, change, objective = 0, 1.0, 1.0                                          # :140-143
        objective_table = [0.0]
        while (change > ps["tolerance"] and count < ps["maxCount"]) or count <= 5:      # :162
            learning = {"trainingData": X, "trainingTargets": learned, "testData": X[cens], "events": events}
            reg = GPRegression(logHyp, ps, learning, lnc)                               # :170
            pred = GPPredict(reg, ps, learning, reg.get("logHypNoiseCorrection"))       # :171
            logHyp = reg["logHypChosen"]                                                # :185
            objective_new = reg["objective"]                                            # :187
            objective_table.append(objective_new)
            adj = AdjustTrainingSurvivalMeanVariance(a_start, pred["funcMeanPred"].ravel(), pred["varData"])   # :190
            change = abs(objective - objective_new)                                     # :195
            objective = objective_new
            count += 1
            learned = learned.copy()
            learned[cens] = adj["mu"]                                                   # :201
            if noiseCorr == "noiseCorrVec":                                             # :208-210
                with np.errstate(divide="ignore"):
                    lnc = np.log(adj["sigma"])
                lnc[np.isinf(lnc)] = -1e10
            elif noiseCorr == "noiseCorrLearned":                                       # :211-212
                lnc = reg["logHypNoiseCorrection"]
        out["convergence"] = 0 if count < ps["maxCount"] else 1                         # :216-221
        out["count"] = count
        out["objectiveTable"] = np.array(objective_table)
        out["trainingTargetsLearned"] = learned
        final_targets = learned
    else:                                                                               # :224-232
        ps["noiseCorr"] = False
        learning = {"trainingData": X, "trainingTargets": y, "events": events}
        reg = GPRegression(logHyp, ps, learning, None)
        lnc = None
        final_targets = y
    final = {"trainingData": X, "trainingTargets": final_targets, "testData": tts["testData"], "events": events}
    pred = GPPredict(reg, ps, final, lnc)                                               # :243 / :245
    train = dict(final, testData=X)
    pred3 = GPPredict(reg, ps, train, lnc)                                              # :272 / :280
    out.update({"logHypChosen": reg["logHypChosen"], "funcMeanPred": pred["funcMeanPred"],
                "varFunction": pred["varFunction"], "varData": pred["varData"],
                "trainingSetPredictions": pred3["funcMeanPred"], "logMarLik": reg["logMarLik"],
                "objective": reg["objective"]})
    return out


# =========================================================================================
# Several GPUs from one host process (gps_fit_batch_multi / gps_fit_folds_multi)
# =========================================================================================
_METHODS = {"Nelder-Mead": 0, "BFGS": 1, "L-BFGS": 2}


def partition_lpt(cost, nparts):
    """gps_partition_lpt: longest-processing-time-first assignment of jobs to `nparts` bins (pure host code in the library)."""
    lib = _lib.load()
    cost = np.ascontiguousarray(np.asarray(cost, dtype=np.float64))
    out = (C.c_int * cost.size)()
    check(lib.gps_partition_lpt(dptr(cost), cost.size, int(nparts), out))
    return np.array(list(out), dtype=int)


def _multi_outputs(T, n, m, ln):
    return dict(par=np.empty((T, ln)), objective=np.empty(T), logMarLik=np.zeros(T), rounds=(C.c_int * T)(),
                learned=np.empty((T, n)), tm=np.empty((T, m)), tvf=np.empty((T, m)), tvd=np.empty((T, m)), trm=np.empty((T, n)),
                status=(C.c_int * T)(), ticks=C.c_longlong())


def _multi_result(o):
    return {"par": o["par"], "objective": o["objective"], "logMarLik": o["logMarLik"], "count": np.array(list(o["rounds"])),
            "trainingTargetsLearned": o["learned"], "funcMeanPred": o["tm"], "varFunction": o["tvf"], "varData": o["tvd"],
            "trainingSetPredictions": o["trm"], "status": np.array(list(o["status"])), "ticks": int(o["ticks"].value)}


def fit_batch_multi(devices, X, y_log, events, par_start, Xtest, cov_form="SqExp", mean_form="Zero", noise_corr=False,
                    ard_mode="intended", extra_param=None, method="Nelder-Mead", maxit=10, tolerance=1e-3, max_count=500,
                    survival=True):
    """gps_fit_batch_multi: `total` independent fits (X [T, n, d], Xtest [T, m, d]) over the listed devices, one host thread and
    one handle per device behind the C ABI, no collective."""
    lib = _lib.load()
    X = np.asarray(X, dtype=np.float64)
    Xt = np.asarray(Xtest, dtype=np.float64)
    T, n, d = X.shape
    m = Xt.shape[1]
    Xc = np.ascontiguousarray(np.transpose(X, (0, 2, 1)))
    Xtc = np.ascontiguousarray(np.transpose(Xt, (0, 2, 1)))
    y = np.ascontiguousarray(np.asarray(y_log, dtype=np.float64)).reshape(T, n)
    ev = np.ascontiguousarray(np.asarray(events, dtype=np.float64)).reshape(T, n)
    ps = np.ascontiguousarray(np.asarray(par_start, dtype=np.float64)).reshape(T, -1)
    ln = ps.shape[1]
    o = _multi_outputs(T, n, m, ln)
    dev = (C.c_int * len(devices))(*[int(v) for v in devices])
    check(lib.gps_fit_batch_multi(dev, len(devices), T, cov_code(cov_form, extra_param), MEAN[mean_form], NOISE_CORR[noise_corr],
                                  ARD_MODE[ard_mode], dptr(Xc), n, d, dptr(y), dptr(ev), dptr(ps), ln, _METHODS[method], int(maxit),
                                  float(tolerance), int(max_count), 1 if survival else 0, dptr(Xtc), m, dptr(o["par"]),
                                  dptr(o["objective"]), dptr(o["logMarLik"]), o["rounds"], dptr(o["learned"]), dptr(o["tm"]),
                                  dptr(o["tvf"]), dptr(o["tvd"]), dptr(o["trm"]), o["status"], C.byref(o["ticks"])))
    return _multi_result(o)


def fit_folds_multi(devices, Xpool, y_log_pool, events_pool, train_rows, test_rows, par_start, cov_form="SqExp", mean_form="Zero",
                    noise_corr=False, ard_mode="intended", extra_param=None, method="Nelder-Mead", maxit=10, tolerance=1e-3,
                    max_count=500, survival=True):
    """gps_fit_folds_multi: CV folds / restarts of ONE data set as row-index views (train_rows [T, n], test_rows [T, m]); the
    pool crosses the bus once per device."""
    lib = _lib.load()
    Xp = f64(Xpool)
    npool, d = Xp.shape
    tr = np.ascontiguousarray(np.asarray(train_rows, dtype=np.int32))
    te = np.ascontiguousarray(np.asarray(test_rows, dtype=np.int32))
    T, n = tr.shape
    m = te.shape[1]
    yp = np.ascontiguousarray(np.asarray(y_log_pool, dtype=np.float64))
    evp = np.ascontiguousarray(np.asarray(events_pool, dtype=np.float64))
    ps = np.ascontiguousarray(np.asarray(par_start, dtype=np.float64)).reshape(T, -1)
    ln = ps.shape[1]
    o = _multi_outputs(T, n, m, ln)
    dev = (C.c_int * len(devices))(*[int(v) for v in devices])
    ip = C.POINTER(C.c_int)
    check(lib.gps_fit_folds_multi(dev, len(devices), T, cov_code(cov_form, extra_param), MEAN[mean_form], NOISE_CORR[noise_corr],
                                  ARD_MODE[ard_mode], dptr(Xp), npool, d, dptr(yp), dptr(evp), tr.ctypes.data_as(ip), n,
                                  te.ctypes.data_as(ip), m, dptr(ps), ln, _METHODS[method], int(maxit), float(tolerance),
                                  int(max_count), 1 if survival else 0, dptr(o["par"]), dptr(o["objective"]), dptr(o["logMarLik"]),
                                  o["rounds"], dptr(o["learned"]), dptr(o["tm"]), dptr(o["tvf"]), dptr(o["tvd"]), dptr(o["trm"]),
                                  o["status"], C.byref(o["ticks"])))
    return _multi_result(o)
