"""Timeline of the one-kernel Cholesky (+ inverse) from its per-task time stamps (GPS_DAG_TRACE=1):
python scripts/dag_trace.py B dag groups slice lat offset"""
import os, sys, ctypes as C
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
B = int(sys.argv[1]); dag = sys.argv[2]
os.environ.update(GPS_DAG=dag, GPS_DAG_GROUPS=sys.argv[3], GPS_DAG_SLICE=sys.argv[4], GPS_DAG_LAT=sys.argv[5],
                  GPS_DAG_OFFSET=sys.argv[6], GPS_DAG_TRACE="1")
import bench
from gpsurvival_b200 import gp
X, y, ev, lnc, par0 = bench.batch_workload(0, B)
rng = np.random.default_rng(0)
par = par0[None, :] + 0.05 * rng.standard_normal((B, par0.size))
fit = gp.Fit(0, batch=B)
fit.set_options("ARD", "Zero", "noiseCorrVec", "intended")
fit.set_data(X, y, ev)
fit.set_noise_corr(lnc)
for _ in range(4):
    fit.nlml_grad(par)
plan = 1 if dag == "2" else 0
nt = fit.lib.gps_debug_dag_trace(fit.h, plan, None, 0)
buf = np.zeros(4 * nt, dtype=np.int64)
fit.lib.gps_debug_dag_trace(fit.h, plan, buf.ctypes.data_as(C.POINTER(C.c_longlong)), buf.size)
fit.close()
tr = buf.reshape(nt, 4)
ok = tr[:, 2] > 0
tr = tr[ok]
t0 = tr[:, 0].min()
pick, ready, done = (tr[:, 0] - t0) / 1e3, (tr[:, 1] - t0) / 1e3, (tr[:, 2] - t0) / 1e3
meta = tr[:, 3]
sm = meta & 0xff; leaf = (meta >> 8) & 1; K16 = (meta >> 9) & 0x7fff; fitid = (meta >> 24) & 0xff; seg = meta >> 32
span = done.max()
spin, run = ready - pick, done - ready
print(f"tasks {len(tr)} (of {nt}), span {span:.0f} us; CTA-time: spin {spin.sum()/1e3:.1f} ms, run {run.sum()/1e3:.1f} ms; "
      f"capacity {444 * span / 1e3:.1f} ms -> busy {100 * run.sum() / (444 * span):.1f} %, spinning {100 * spin.sum() / (444 * span):.1f} %")
for name, m in [("leaf", leaf == 1), ("K<=128", (leaf == 0) & (K16 <= 8)), ("K=256", (leaf == 0) & (K16 > 8) & (K16 <= 16)),
                ("K=512", (leaf == 0) & (K16 > 16) & (K16 <= 32)), ("K>512", (leaf == 0) & (K16 > 32))]:
    if m.any():
        print(f"  {name:7s}: {m.sum():6d} tasks, run mean {run[m].mean():7.1f} us (p50 {np.median(run[m]):.1f}, p95 {np.percentile(run[m], 95):.1f}), "
              f"spin mean {spin[m].mean():6.1f} us (p95 {np.percentile(spin[m], 95):.1f}), CTA-time run {run[m].sum()/1e3:.1f} ms spin {spin[m].sum()/1e3:.1f} ms")
# utilisation over time: CTAs running (not spinning) per 100 us bucket
nb = int(span // 100) + 1
busy = np.zeros(nb); sp = np.zeros(nb)
for a, b, arr in [(ready, done, busy), (pick, ready, sp)]:
    for i in range(nb):
        lo, hi = 100.0 * i, 100.0 * (i + 1)
        arr[i] = np.clip(np.minimum(b, hi) - np.maximum(a, lo), 0, None).sum() / 100.0
print("running CTAs per 100 us:", " ".join(f"{v:.0f}" for v in busy))
print("spinning CTAs per 100 us:", " ".join(f"{v:.0f}" for v in sp))
# per-fit chain for fit 0: leaf tasks' times
m = (leaf == 1) & (fitid == 0)
o = np.argsort(pick[m])
print("fit 0 leaves (pick, ready, done us):", " ".join(f"({a:.0f},{b:.0f},{c:.0f})" for a, b, c in zip(pick[m][o][:40], ready[m][o][:40], done[m][o][:40])))
