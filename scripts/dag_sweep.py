"""Step and Cholesky-stage time of the bench workload (C2 x B) for the one-kernel Cholesky (GPS_DAG=1) / Cholesky + triangular
inverse (GPS_DAG=2) under several scheduler settings: python scripts/dag_sweep.py B [dag,groups,slice,lat,offset ...]"""
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
cfgs = [("0", 1, 48, 1.0, 0.0)]
for a in sys.argv[2:]:
    f = a.split(",")
    cfgs.append((f[0], int(f[1]), int(f[2]), float(f[3]), float(f[4])))
for dag, G, S, LAT, OFF in cfgs:
    os.environ.update(GPS_DAG=dag, GPS_DAG_GROUPS=str(G), GPS_DAG_SLICE=str(S), GPS_DAG_LAT=str(LAT), GPS_DAG_OFFSET=str(OFF))
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
    print(f"dag={dag} groups={G} slice={S} lat={LAT} offset={OFF}: step {ms:.3f} ms, cholesky stage {st[1]:.3f} ms, rel diff vs multi-launch {err:.1e}", flush=True)
    fit.close()
