"""Lone-fit latency at the sizes the reference's own drivers use (n = 100 / 200 / 500, SqExp, d small):
device time of one NLML and one NLML+gradient evaluation (CUDA-graph replay) and the wall time of the
public call (H2D of par + graph + D2H + sync), next to the numpy restatement on the host."""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from gpsurvival_b200 import gp
from oracle import gp_oracle as orc
for n, d in [(60, 1), (100, 1), (200, 4), (500, 6), (1000, 10), (2000, 20)]:
    rng = np.random.default_rng(n)
    X = np.asfortranarray(rng.standard_normal((n, d))); y = rng.standard_normal(n); ev = (rng.random(n) > 0.4).astype(float)
    par = np.log([0.2, 0.8, 1.7 * np.sqrt(d)])
    fit = gp.Fit(0); fit.set_options("SqExp", "Zero", False); fit.set_data(X, y, ev)
    for _ in range(3):
        fit.nlml(par + 1e-3 * _); fit.nlml_grad(par + 1e-3 * _)
    l0 = fit.launch_count(); fit.nlml(par + 0.5); lf = fit.launch_count() - l0
    l0 = fit.launch_count(); fit.nlml_grad(par + 0.6); lg = fit.launch_count() - l0
    ms_f = fit.time_eval(0, 100); ms_g = fit.time_eval(1, 100)
    t0 = time.perf_counter()
    for i in range(200):
        fit.nlml(par + 1e-6 * i)
    wall_f = (time.perf_counter() - t0) / 200
    reps = 20 if n <= 500 else 3
    t0 = time.perf_counter()
    for i in range(reps):
        orc.nlml(par, X, y, ev, as_written_ops=True)
    cpu_f = (time.perf_counter() - t0) / reps
    print(f"n={n:5d} d={d:3d}: NLML device {ms_f*1e3:7.1f} us ({lf} kernels), NLML+grad {ms_g*1e3:7.1f} us ({lg}), public nlml() call {wall_f*1e6:7.1f} us, "
          f"host restatement {cpu_f*1e6:9.1f} us", flush=True)
    fit.close()
