"""Cost of one nearPD repair (ForOptimisationLog.R:43-51 on the device) at several n: duplicated rows, sn2 = exp(-800)."""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from gpsurvival_b200 import gp, synth
for n_half in [int(a) for a in sys.argv[1:]] or [100, 300, 1000]:
    pb = synth.make_problem(n_half, 5, 2, 10, 1)
    X, y, ev = pb["trainingData"], np.log(pb["trainingTargets"]), pb["events"]
    Xd, yd, evd = np.vstack([X, X]), np.concatenate([y, y]), np.concatenate([ev, ev])
    fit = gp.Fit(0)
    fit.set_data(Xd, yd, evd)
    par = np.array([-800.0, 0.0, 0.0])
    t0 = time.perf_counter()
    try:
        f = fit.nlml(par)
        used = fit.near_pd_used()
    except Exception as exc:
        f, used = repr(exc), None
    dt = time.perf_counter() - t0
    print(f"n = {2 * n_half}: nlml = {f}, nearPD used {used}, {dt:.2f} s", flush=True)
    fit.close()
