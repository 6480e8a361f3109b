"""Whole-fit throughput through the C ABI (gps_fit_batch) — same protocol as scripts/fit_bench.py."""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from gpsurvival_b200 import synth, gp
name = sys.argv[1] if len(sys.argv) > 1 else "C2"
B = int(sys.argv[2]) if len(sys.argv) > 2 else 16
method = sys.argv[3] if len(sys.argv) > 3 else "Nelder-Mead"
cfg = synth.CONFIGS[name]
base = [synth.make_config_problem(cfg, fit_index=i) for i in range(min(B, 4))]
pbs = [base[i % len(base)] for i in range(B)]
d = pbs[0]["trainingData"].shape[1]
rng = np.random.default_rng(1)
starts = [synth.start_log_hyp(d, cfg.cov_form, rng) for _ in range(B)]
X = np.stack([p["trainingData"] for p in pbs]); Xt = np.stack([p["testData"] for p in pbs])
y = np.log(np.stack([p["trainingTargets"] for p in pbs])); ev = np.stack([p["events"] for p in pbs])
par0 = np.stack([gp._flatten(s, d, "Zero", cfg.noise_corr, None) for s in starts])
fit = gp.Fit(0, batch=B)
fit.set_options(cfg.cov_form, "Zero", cfg.noise_corr)
for rep in range(2):
    t = time.time()
    out = fit.fit_batch(X, y, ev, par0, Xt, method=method, maxit=10, tolerance=1e300, max_count=6, survival=True)
    dt = time.time() - t
    print(f"{name} B={B} {method} [C ABI]: {dt:.3f}s -> {B/dt:.2f} fits/s ; rounds={out['count'][0]} ticks={out['ticks']} objective[0]={out['objective'][0]:.4f}", flush=True)
