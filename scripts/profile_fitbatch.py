"""Short program for ncu / timing: whole GPS3 fits of a named config behind the C ABI (gps_fit_batch).
Usage: python scripts/profile_fitbatch.py C4 32 BFGS 2"""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from gpsurvival_b200 import gp, synth
name = sys.argv[1] if len(sys.argv) > 1 else "C4"
B = int(sys.argv[2]) if len(sys.argv) > 2 else 32
method = sys.argv[3] if len(sys.argv) > 3 else "BFGS"
rounds = int(sys.argv[4]) if len(sys.argv) > 4 else 6
cfg = synth.CONFIGS[name]
pbs = [synth.make_config_problem(cfg, fit_index=i) for i in range(min(B, 4))]
d = pbs[0]["trainingData"].shape[1]
X = np.stack([pbs[i % len(pbs)]["trainingData"] for i in range(B)])
y = np.stack([np.log(pbs[i % len(pbs)]["trainingTargets"]) for i in range(B)])
ev = np.stack([pbs[i % len(pbs)]["events"] for i in range(B)])
Xt = np.stack([pbs[i % len(pbs)]["testData"] for i in range(B)])
rng = np.random.default_rng(3)
par0 = np.stack([gp._flatten(synth.start_log_hyp(d, cfg.cov_form, rng), d, "Zero", cfg.noise_corr, None) for _ in range(B)])
fit = gp.Fit(0, batch=B)
fit.set_options(cfg.cov_form, "Zero", cfg.noise_corr, "intended")
for rep in range(2):
    l0 = fit.launch_count()
    t0 = time.perf_counter()
    rr = fit.fit_batch(X, y, ev, par0, Xt, method=method, maxit=10, tolerance=1e300, max_count=rounds, survival=True)
    dt = time.perf_counter() - t0
    print(f"{name} B={B} {method} rounds={rounds}: {dt:.3f} s, {B/dt:.1f} fits/s, ticks={rr['ticks']}, launches={fit.launch_count()-l0}", flush=True)
