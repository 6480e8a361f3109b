"""Full GPS3 fits / s: the fixed protocol of SURVEY.md §8d — 6 outer GPS rounds (the reference's
minimum, ApplyGP.R:162) x one optim() run of maxitSurvival iterations + test and training-set
predictions — for a batch of independent C2- or C4-shaped fits on one GPU."""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from gpsurvival_b200 import batchfit, synth, gp

name = sys.argv[1] if len(sys.argv) > 1 else "C2"
B = int(sys.argv[2]) if len(sys.argv) > 2 else 16
method = sys.argv[3] if len(sys.argv) > 3 else "Nelder-Mead"
cfg = synth.CONFIGS[name]
t = time.time()
base = [synth.make_config_problem(cfg, fit_index=i) for i in range(min(B, 4))]
pbs = [base[i % len(base)] for i in range(B)]
d = pbs[0]["trainingData"].shape[1]
rng = np.random.default_rng(1)
starts = [synth.start_log_hyp(d, cfg.cov_form, rng) for _ in range(B)]      # restarts: perturbed start hyper-parameters
print(f"data generation {time.time()-t:.1f}s", flush=True)
params = dict(modelType="survival", noiseCorr=cfg.noise_corr, optimType=method, maxit=100, maxitSurvival=10,
              meanFuncForm="Zero", covFuncForm=cfg.cov_form, tolerance=1e300, maxCount=6, logHypStart=starts,
              burnIn=False, imposePriors=False)
fit = gp.Fit(0, batch=B)
for rep in range(2):
    t = time.time()
    res = batchfit.apply_gp_batch(pbs, params, fit=fit)
    dt = time.time() - t
    print(f"{name} B={B} {method}: {dt:.3f}s -> {B/dt:.2f} fits/s ; rounds={res[0]['count']} ticks={res[0]['ticks']} launches={fit.launch_count()} "
          f"objective[0]={res[0]['objective']:.4f}", flush=True)
