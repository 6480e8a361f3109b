"""Small end-to-end exercise for compute-sanitizer: every kernel family at odd sizes."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from gpsurvival_b200 import gp, synth, batchfit
for (n, m, d, c, cov, mean, nc, extra) in [(131, 37, 3, 50, "ARD", "Linear", "noiseCorrLearned", None),
                                           (67, 9, 300, 20, "ARD", "Zero", "noiseCorrVec", None),
                                           (40, 5, 1, 10, "Matern", "Zero", False, 5)]:
    pb = synth.make_problem(n, m, d, c, 3)
    X, y, ev, Xs = pb["trainingData"], np.log(pb["trainingTargets"]), pb["events"], pb["testData"]
    fit = gp.Fit(0)
    fit.set_options(cov, mean, nc, "intended", extra)
    fit.set_data(X, y, ev)
    p = d if cov == "ARD" else 1
    par = [np.log(0.2), np.log(0.8)] + [np.log(1.7 * np.sqrt(d))] * p + ([0.01] * (d + 1) if mean == "Linear" else []) + \
          ([np.log(0.3)] if nc == "noiseCorrLearned" else [])
    if nc == "noiseCorrVec":
        fit.set_noise_corr(np.full(c, np.log(0.2)))
    par = np.array(par)
    for _ in range(3):
        f = fit.nlml(par)
    if cov != "Matern":
        for _ in range(3):
            f, g = fit.nlml_grad(par)
    fit.finalize(par, want_K=True, want_L=True)
    mu, vf, vd = fit.predict(par, Xs)
    fit.adjust(y[:c], mu[:c] if mu.size >= c else np.resize(mu, c), np.resize(vd, c))
    fit.cov(X, Xs, 0.8, np.exp(par[2:2 + p]), cov, "intended", extra)
    fit.cov(X, None, 0.8, np.exp(par[2:2 + p]), cov, "as_written" if cov == "ARD" else "intended", extra)
    fit.close()
    print("ok", n, d, cov, f)
pbs = [synth.make_problem(70, 15, 2, 28, 900 + b) for b in range(3)]
params = dict(modelType="survival", noiseCorr="noiseCorrVec", optimType="BFGS", maxit=20, maxitSurvival=3, meanFuncForm="Zero",
              covFuncForm="ARD", tolerance=1e300, maxCount=2, logHypStart=synth.start_log_hyp(2, "ARD"), burnIn=False, imposePriors=False)
r = batchfit.apply_gp_batch(pbs, params)
print("batch ok", r[0]["objective"])
