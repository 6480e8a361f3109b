"""C1 = runSyntheticExp1.R's default (the only configuration the reference ships: n = 100, d = 1, 40 % censored,
SqExp, Nelder-Mead; runSyntheticExp1.R:92-139).  Far too small for a roofline (n^3/3 = 3.3e5 flop): this reports
absolute evaluations/s and whole GPS1 / GPR fits/s, lone and as a batch of repetitions (the driver's nReps loop,
runSyntheticExp1.R:208-213), next to the numpy restatement of the reference on the host cores.
Usage: python scripts/c1_bench.py [reps]"""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from gpsurvival_b200 import gp, synth
from oracle import gp_oracle as orc

reps = int(sys.argv[1]) if len(sys.argv) > 1 else 256
cfg = synth.CONFIGS["C1"]
pbs = [synth.make_config_problem(cfg, fit_index=i) for i in range(min(reps, 32))]
n, d = pbs[0]["trainingData"].shape
par0 = synth.start_par(cfg, d=d)

def arrays(B):
    X = np.stack([pbs[i % len(pbs)]["trainingData"] for i in range(B)])
    y = np.stack([np.log(pbs[i % len(pbs)]["trainingTargets"]) for i in range(B)])
    ev = np.stack([pbs[i % len(pbs)]["events"] for i in range(B)])
    Xt = np.stack([pbs[i % len(pbs)]["testData"] for i in range(B)])
    return X, y, ev, Xt

out = {"config": f"C1: n={n} d={d} SqExp (runSyntheticExp1.R default)"}
for B in (1, reps):
    X, y, ev, Xt = arrays(B)
    fit = gp.Fit(0, batch=B)
    fit.set_options("SqExp", "Zero", False, "intended")
    fit.set_data(X, y, ev)
    par = np.tile(par0, (B, 1)) + 0.05 * np.random.default_rng(B).standard_normal((B, par0.size))
    for _ in range(3):
        fit.nlml_grad(par); fit.nlml(par + 1e-3)
    ms_g = fit.time_eval(1, 50)
    ms_f = fit.time_eval(0, 50)
    t0 = time.perf_counter()
    rr = fit.fit_batch(X, y, ev, par, Xt, method="Nelder-Mead", maxit=10, tolerance=1e-3, max_count=500, survival=True)
    t_gps1 = time.perf_counter() - t0
    t0 = time.perf_counter()
    r2 = fit.fit_batch(X, y, ev, par, Xt, method="Nelder-Mead", maxit=2000, survival=False)
    t_gpr = time.perf_counter() - t0
    out[f"batch_{B}"] = {"nlml_evals_per_s": B / ms_f * 1e3, "nlml_grad_evals_per_s": B / ms_g * 1e3,
                         "gps1_fits_per_s": B / t_gps1, "gps1_rounds_mean": float(np.mean(rr["count"])), "gps1_ticks": rr["ticks"],
                         "gpr_fits_per_s": B / t_gpr, "gpr_ticks": r2["ticks"]}
    fit.close()
# the reference algorithm restated in numpy, one fit on the host
pb = pbs[0]
params = dict(modelType="survival", noiseCorr=False, optimType="Nelder-Mead", maxit=2000, maxitSurvival=10, meanFuncForm="Zero",
              covFuncForm="SqExp", tolerance=1e-3, maxCount=500, logHypStart=synth.start_log_hyp(d, "SqExp"), burnIn=False,
              imposePriors=False)
t0 = time.perf_counter(); ro = orc.apply_gp(pb, params); t_cpu = time.perf_counter() - t0
X1, y1, ev1 = pb["trainingData"], np.log(pb["trainingTargets"]), pb["events"]
t0 = time.perf_counter()
for _ in range(20):
    orc.nlml(par0, X1, y1, ev1, as_written_ops=True)
t_eval = (time.perf_counter() - t0) / 20
out["cpu_port"] = {"gps1_fits_per_s": 1.0 / t_cpu, "gps1_rounds": int(ro["count"]), "nlml_evals_per_s": 1.0 / t_eval,
                   "cores": os.cpu_count(), "note": "numpy restatement of the reference; R unavailable"}
import json
print(json.dumps(out))
