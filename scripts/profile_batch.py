"""Short program for ncu: the bench workload (C2, 32 independent fits per launch), 3 warm-up
batched NLML+gradient evaluations then 2 more, CUDA graphs off so every kernel is a launch."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from gpsurvival_b200 import gp
B = int(sys.argv[1]) if len(sys.argv) > 1 else 32
X, y, ev, lnc, par0 = bench.batch_workload(0, B)
fit = gp.Fit(0, batch=B)
fit.set_graphs(False)
fit.set_options("ARD", "Zero", "noiseCorrVec", "intended")
fit.set_data(X, y, ev)
fit.set_noise_corr(lnc)
rng = np.random.default_rng(0)
l0 = fit.launch_count()
for i in range(5):
    f, g = fit.nlml_grad(par0[None, :] + 0.05 * rng.standard_normal((B, par0.size)))
    if i == 0:
        per = fit.launch_count() - l0
print(f"B={B} nlml[0]={f[0]:.6f} launches_per_eval={per} setup_launches={l0}")
