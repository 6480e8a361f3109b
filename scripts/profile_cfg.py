"""Short program for ncu: `batch` independent fits of a named config (C2/C3/C4), 3 warm-up batched
NLML+gradient evaluations then 2 more, CUDA graphs off so every kernel is a launch.
Usage: python scripts/profile_cfg.py C3 32"""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from gpsurvival_b200 import gp, synth
name = sys.argv[1] if len(sys.argv) > 1 else "C3"
B = int(sys.argv[2]) if len(sys.argv) > 2 else 32
cfg = synth.CONFIGS[name]
pbs = [synth.make_config_problem(cfg, fit_index=i) for i in range(min(B, 4))]
n, d = pbs[0]["trainingData"].shape
X = np.stack([pbs[i % len(pbs)]["trainingData"] for i in range(B)])
y = np.stack([np.log(pbs[i % len(pbs)]["trainingTargets"]) for i in range(B)])
ev = np.stack([pbs[i % len(pbs)]["events"] for i in range(B)])
fit = gp.Fit(0, batch=B)
fit.set_graphs(False)
fit.set_options(cfg.cov_form, "Zero", cfg.noise_corr, "intended")
fit.set_data(X, y, ev)
if cfg.noise_corr == "noiseCorrVec":
    fit.set_noise_corr(np.full((B, int(np.sum(ev[0] == 0))), np.log(0.2)))
rng = np.random.default_rng(0)
par0 = synth.start_par(cfg, d=d)
l0 = fit.launch_count()
for i in range(5):
    f, g = fit.nlml_grad(par0[None, :] + 0.05 * rng.standard_normal((B, par0.size)))
    if i == 0:
        per = fit.launch_count() - l0
print(f"{name} B={B} nlml[0]={f[0]:.6f} launches_per_eval={per} setup_launches={l0}")
