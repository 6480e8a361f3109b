import os, sys, ctypes as C
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from gpsurvival_b200 import gp, synth
cfg = synth.CONFIGS["C2"]
pb = synth.make_config_problem(cfg, scale=0.25)
X, y, ev = pb["trainingData"], np.log(pb["trainingTargets"]), pb["events"]
fit = gp.Fit(0)
fit.set_options("ARD", "Zero", "noiseCorrVec")
fit.set_data(X, y, ev)
fit.set_noise_corr(np.full(int(np.sum(ev == 0)), np.log(0.2)))
fit.nlml(synth.start_par(cfg))
out = (C.c_longlong * 16)()
fit._ck(fit.lib.gps_debug_panel_clocks(fit.h, out, 16))
v = list(out)
names = ["start", "stage+sync", "sub0", "trail0", "sub1", "trail1", "sub2", "trail2", "sub3", "trinv16", "merges", "publish", "trsm+store"]
prev = 0
for nm, c in zip(names, v):
    print(f"{nm:12s} t={c:8d}  d={c-prev:7d} cycles")
    prev = c
