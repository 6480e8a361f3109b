"""Cholesky-stage time of the bench workload (C2 x 64) for the one-kernel Cholesky under several (groups, skew) settings."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from gpsurvival_b200 import gp
B = int(sys.argv[1]) if len(sys.argv) > 1 else 64
X, y, ev, lnc, par0 = bench.batch_workload(0, B)
rng = np.random.default_rng(0)
par = par0[None, :] + 0.05 * rng.standard_normal((B, par0.size))
ref = None
cfgs = [("0", 1, 0)] + [("1", g, s) for g in (1, 2, 4, 8) for s in (0, 6, 12, 20)]
if len(sys.argv) > 2:
    cfgs = [("0", 1, 0)] + [("1", int(a.split(",")[0]), int(a.split(",")[1])) for a in sys.argv[2:]]
for dag, G, S in cfgs:
    os.environ["GPS_DAG"] = dag
    os.environ["GPS_DAG_GROUPS"] = str(G)
    os.environ["GPS_DAG_SKEW"] = str(S)
    fit = gp.Fit(0, batch=B)
    fit.set_options("ARD", "Zero", "noiseCorrVec", "intended")
    fit.set_data(X, y, ev)
    fit.set_noise_corr(lnc)
    for _ in range(3):
        f, g = fit.nlml_grad(par)
    if ref is None:
        ref = (f.copy(), g.copy())
    ms = fit.time_eval(kind=1, reps=10)
    fit.time_stages()
    st = fit.time_stages()
    err = max(np.max(np.abs(f - ref[0]) / np.abs(ref[0])), np.max(np.abs(g - ref[1])) / np.max(np.abs(ref[1])))
    print(f"dag={dag} groups={G} skew={S}: step {ms:.3f} ms, cholesky stage {st[1]:.3f} ms, rel diff vs multi-launch {err:.1e}", flush=True)
    fit.close()
