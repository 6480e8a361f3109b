"""Short program for ncu: 3 warm-up NLML+gradient evaluations of the bench workload (C2 by
default), then 2 more.  Usage: python scripts/profile_eval.py [C2|C3|C1] [graphs:0|1]"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from gpsurvival_b200 import gp, synth   # noqa: E402

name = sys.argv[1] if len(sys.argv) > 1 else "C2"
graphs = int(sys.argv[2]) if len(sys.argv) > 2 else 0
cfg = synth.CONFIGS[name]
pb = synth.make_config_problem(cfg)
X, y, ev = pb["trainingData"], np.log(pb["trainingTargets"]), pb["events"]
fit = gp.Fit(0)
fit.set_graphs(bool(graphs))
fit.set_options(cfg.cov_form, "Zero", cfg.noise_corr, "intended")
fit.set_data(X, y, ev)
if cfg.noise_corr == "noiseCorrVec":
    fit.set_noise_corr(np.full(int(np.sum(ev == 0)), np.log(0.2)))
par = synth.start_par(cfg, d=X.shape[1])
l0 = fit.launch_count()
for i in range(5):
    f, g = fit.nlml_grad(par + 1e-3 * i)
    if i == 0:
        per = fit.launch_count() - l0
print(f"{name}: nlml={f:.6f} launches_per_eval={per} setup_launches={l0}")
fit.close()
