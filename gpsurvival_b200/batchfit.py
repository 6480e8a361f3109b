"""Batched whole fits: many independent ApplyGP runs (restarts, CV folds, repetitions) advanced
in lock-step on one GPU through a batched ``gps_handle``.

The control flow per fit is exactly ApplyGP.R:77-282 / GPRegression.R:1-83 (the same code path
as ``gp.ApplyGP``): every fit owns its own optimiser state (a generator from ``optim.py``), and
each tick of the driver collects the next evaluation point of every fit, serves all of them with
ONE batched NLML(+gradient) evaluation, and feeds the values back.  Fits that have finished an
optimiser run, or converged in the GPS loop, simply idle (their slot is evaluated at their last
point, which costs nothing extra — the batched launch covers the whole batch anyway).

Requirements: all fits of a batch share (n, d, number of censored rows) and the option set.
Nothing here imports ``oracle``; all arithmetic is in libgpsurv.so.
"""
from __future__ import annotations

import numpy as np

from . import gp
from . import optim as _optim


def _flatten_batch(log_hyps, d, mean_form, noise_corr, lncs):
    return np.stack([gp._flatten(lh, d, mean_form, noise_corr, None if lncs is None else lncs[b])
                     for b, lh in enumerate(log_hyps)])


def _run_optimisers(fit, starts, method, maxit, active):
    """Advance one optimiser per fit to completion.  Returns list of optim() result dicts (None
    for inactive fits).  Evaluations per tick: one batched call."""
    B = starts.shape[0]
    if method == "BFGS":
        # dense inverse-Hessian state of all fits in one GPU tensor, vmmin's control flow as masks
        from .batch_bfgs import batch_bfgs

        def evaluate(P):
            f, g = fit.nlml_grad(P, not_pd="inf")
            return np.atleast_1d(f), np.atleast_2d(g)

        return batch_bfgs(evaluate, starts, maxit=maxit, active=active, device=f"cuda:{fit.device}")
    gens = [(_optim.optim_steps(starts[b], method, maxit) if active[b] else None) for b in range(B)]
    req = [None] * B
    results = [None] * B
    point = starts.copy()
    for b in range(B):
        if gens[b] is not None:
            req[b] = next(gens[b])
            point[b] = req[b][1]
    ticks = 0
    while any(r is not None for r in req):
        want_grad = (method == "BFGS") or any(r is not None and r[0] == "g" for r in req)
        if want_grad:
            f, g = fit.nlml_grad(point, not_pd="inf")
        else:
            f, g = fit.nlml(point, not_pd="inf"), None
        f = np.atleast_1d(f)
        ticks += 1
        for b in range(B):
            if req[b] is None:
                continue
            try:
                req[b] = gens[b].send((f[b], None if g is None else np.atleast_2d(g)[b].copy()))
                point[b] = req[b][1]
            except StopIteration as stop:
                results[b] = stop.value
                req[b] = None
                point[b] = stop.value["par"]
    return results, ticks


def apply_gp_batch(problems, parameterStructure, data_source="Generate", device=0, fit=None):
    """Run ApplyGP (ApplyGP.R:77-282) for every element of ``problems`` (each a
    trainingTestStructure dict) in lock-step on one GPU.  Returns a list of result dicts with
    the same fields as ``gp.ApplyGP`` plus ``ticks`` (batched evaluations used)."""
    ps = dict(parameterStructure)
    B = len(problems)
    model_type, noise_corr = ps["modelType"], ps["noiseCorr"]
    if model_type != "survival":
        noise_corr = False
    cov_form, mean_form = ps["covFuncForm"], ps["meanFuncForm"]
    method = ps["optimType"]
    maxit = ps["maxitSurvival"] if model_type == "survival" else ps["maxit"]
    X = np.stack([np.asarray(p["trainingData"], dtype=np.float64) for p in problems])
    Xtest = np.stack([np.asarray(p["testData"], dtype=np.float64) for p in problems])
    ev = np.stack([np.asarray(p["events"], dtype=np.float64).ravel() for p in problems])
    y = np.log(np.stack([np.asarray(p["trainingTargets"], dtype=np.float64).ravel() for p in problems]))   # :79
    if model_type == "survival" and data_source in gp._REAL_SOURCES:
        y = y - np.median(y, axis=1, keepdims=True)                                                         # :86-92
    n, d = X.shape[1], X.shape[2]
    cens = ev == 0
    ncens = cens.sum(axis=1)
    if model_type == "survival" and not np.all(ncens == ncens[0]):
        raise ValueError("all fits of a batch must have the same number of censored rows")
    c = int(ncens[0])
    own = fit is None
    if own:
        fit = gp.Fit(device, batch=B)
    nc_opt = noise_corr if noise_corr in ("noiseCorrLearned", "noiseCorrVec") else False
    fit.set_options(cov_form, mean_form, nc_opt, gp.SESSION.ard_mode, ps.get("extraParam", ps.get("maternParam")))
    fit.set_data(X, y, ev)
    starts = ps["logHypStart"]
    log_hyps = list(starts) if isinstance(starts, (list, tuple)) else [starts] * B      # per-fit restarts allowed
    total_ticks = 0
    out = [dict() for _ in range(B)]

    def finalize_and_chosen(vecs):
        fin = fit.finalize(vecs)
        chosen, lnc_out = [], []
        for b in range(B):
            body = vecs[b][:-1] if noise_corr == "noiseCorrLearned" else vecs[b]
            p = d if cov_form == "ARD" else 1
            mean = body[body.size - d - 1:] if mean_form == "Linear" else np.zeros(d + 1)
            chosen.append({"noise": float(np.log(np.exp(body[0]))), "func": float(np.log(np.exp(body[1]))),
                           "length": np.log(np.exp(body[2:2 + p])), "mean": np.array(mean, dtype=np.float64)})   # :73
            lnc_out.append(vecs[b][-1:] if noise_corr == "noiseCorrLearned" else None)
        return fin, chosen, lnc_out

    if model_type == "survival":
        a_start = np.stack([y[b][cens[b]] for b in range(B)])                           # :108
        learned = y.copy()                                                              # :145
        if noise_corr == "noiseCorrLearned":                                            # :146-152
            lnc = [np.atleast_1d(np.asarray(lh["noise"], dtype=np.float64)) for lh in log_hyps]
        elif noise_corr == "noiseCorrVec":
            lnc = np.stack([np.full(c, float(np.ravel(lh["noise"])[0])) for lh in log_hyps])
        else:
            lnc = None
        count = np.zeros(B, dtype=int)
        change = np.ones(B)
        objective = np.ones(B)
        tables = [[0.0] for _ in range(B)]
        Xcens = np.stack([X[b][cens[b]] for b in range(B)])
        active = np.ones(B, dtype=bool)
        last_vec = None
        # targets / noise correction as used by each fit's LAST regression: a fit that has left the
        # loop keeps the state its regressionStructure was built from (GPS1's final GPPredict uses
        # that alpha, ApplyGP.R:243 with GPPredict.R:12-13), while the others move on
        used_targets = learned.copy()
        used_lnc = None if lnc is None else [np.array(v, dtype=np.float64) for v in lnc]
        while True:
            active = ((change > ps["tolerance"]) & (count < ps["maxCount"])) | (count <= 5)       # :162
            if not active.any():
                break
            for b in range(B):
                if active[b]:
                    used_targets[b] = learned[b]
                    if used_lnc is not None:
                        used_lnc[b] = np.array(lnc[b], dtype=np.float64)
            fit.set_targets(used_targets)
            if noise_corr == "noiseCorrVec":
                fit.set_noise_corr(np.stack(used_lnc))
            starts_vec = _flatten_batch(log_hyps, d, mean_form, noise_corr,
                                        used_lnc if noise_corr == "noiseCorrLearned" else None)
            res, ticks = _run_optimisers(fit, starts_vec, method, maxit, active)                   # :170 (optim)
            total_ticks += ticks
            vecs = np.stack([res[b]["par"] if active[b] else last_vec[b] for b in range(B)])
            fin, chosen, lnc_l = finalize_and_chosen(vecs)                                          # GPRegression.R:60-82
            pars_pred = np.stack([gp._par_from_chosen(chosen[b], d, mean_form, noise_corr, lnc_l[b]) for b in range(B)])
            mu, vf, vd = (np.atleast_2d(v) for v in fit.predict(pars_pred, Xcens, reuse_model=not bool(nc_opt)))   # :171
            em, evr = fit.adjust(a_start.ravel(), mu.ravel(), vd.ravel())                           # :190
            em, evr = em.reshape(B, c), evr.reshape(B, c)
            new_learned = learned.copy()
            for b in range(B):
                if not active[b]:
                    continue
                log_hyps[b] = chosen[b]                                                             # :185
                obj_new = res[b]["value"]
                tables[b].append(obj_new)
                change[b] = abs(objective[b] - obj_new)                                             # :195
                objective[b] = obj_new
                count[b] += 1
                new_learned[b][cens[b]] = em[b]                                                     # :201
                if noise_corr == "noiseCorrVec":                                                    # :208-210
                    with np.errstate(divide="ignore"):
                        v = np.log(evr[b])
                    v[np.isinf(v)] = -1e10
                    lnc[b] = v
                elif noise_corr == "noiseCorrLearned":                                              # :211-212
                    lnc[b] = lnc_l[b]
                out[b]["logMarLik"] = float(np.atleast_1d(fin["logMarLik"])[b])
            learned = new_learned
            last_vec = vecs
        for b in range(B):
            out[b].update(convergence=0 if count[b] < ps["maxCount"] else 1, count=int(count[b]),
                          objectiveTable=np.array(tables[b]), trainingTargetsLearned=learned[b].copy(),
                          objective=float(objective[b]))
        final_targets = learned
        final_lnc = lnc
        final_chosen = log_hyps
    else:                                                                               # :224-232
        starts_vec = _flatten_batch(log_hyps, d, mean_form, False, None)
        res, ticks = _run_optimisers(fit, starts_vec, method, maxit, np.ones(B, dtype=bool))
        total_ticks += ticks
        vecs = np.stack([r["par"] for r in res])
        fin, final_chosen, _ = finalize_and_chosen(vecs)
        for b in range(B):
            out[b].update(objective=float(res[b]["value"]), logMarLik=float(np.atleast_1d(fin["logMarLik"])[b]))
        final_targets, final_lnc = y, None

    # final predictions (ApplyGP.R:237-282): post-update targets and noise correction
    if model_type == "survival" and nc_opt:
        fit.set_targets(final_targets)
        if noise_corr == "noiseCorrVec":
            fit.set_noise_corr(np.stack([np.asarray(v, dtype=np.float64) for v in final_lnc]))
    pars_pred = np.stack([gp._par_from_chosen(final_chosen[b], d, mean_form, noise_corr,
                                              final_lnc[b] if noise_corr == "noiseCorrLearned" else None)
                          for b in range(B)])
    reuse = not bool(nc_opt)
    mu, vf, vd = (np.atleast_2d(v) for v in fit.predict(pars_pred, Xtest, reuse_model=reuse))   # :243 / :245
    mu_tr = np.atleast_2d(fit.predict(pars_pred, X, reuse_model=reuse)[0])              # :272 / :280
    for b in range(B):
        out[b].update(logHypChosen=final_chosen[b], funcMeanPred=mu[b].reshape(-1, 1), varFunction=vf[b],
                      varData=vd[b], trainingSetPredictions=mu_tr[b].reshape(-1, 1), ticks=total_ticks)
    if own:
        fit.close()
    return out
