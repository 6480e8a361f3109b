"""One-kernel Cholesky (GPS_DAG=1) against the multi-launch path on the same inputs: NLML, gradient, L."""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from gpsurvival_b200 import gp, synth

def run(n, d, B, dag, seed=3):
    os.environ["GPS_DAG"] = "1" if dag else "0"
    rng = np.random.default_rng(seed)
    pbs = [synth.make_problem(n, 4, d, n // 2, 100 + b) for b in range(min(B, 3))]
    X = np.stack([pbs[b % len(pbs)]["trainingData"] for b in range(B)])
    y = np.log(np.stack([pbs[b % len(pbs)]["trainingTargets"] for b in range(B)]))
    ev = np.stack([pbs[b % len(pbs)]["events"] for b in range(B)])
    c = int(np.sum(ev[0] == 0))
    lnc = np.log(0.2) + 0.3 * rng.standard_normal((B, c))
    par = np.concatenate([np.log([0.2, 0.8]), np.log(1.7 * np.sqrt(d)) * np.ones(d)])[None, :] + 0.05 * rng.standard_normal((B, 2 + d))
    fit = gp.Fit(0, batch=B)
    fit.set_options("ARD", "Zero", "noiseCorrVec", "intended")
    fit.set_data(X, y, ev)
    fit.set_noise_corr(lnc)
    out = []
    for rep in range(3):            # direct run, graph capture, replay
        f, g = fit.nlml_grad(par)
        out.append((f.copy(), g.copy()))
    t = fit.time_eval(kind=1, reps=5)
    st = fit.time_stages()
    fin = fit.finalize(par, want_L=True)
    fit.close()
    return out, t, st, fin["L"]

for (n, d, B) in [(300, 3, 9), (130, 2, 16), (600, 4, 32), (2000, 20, 64)]:
    if len(sys.argv) > 1 and int(sys.argv[1]) < n:
        continue
    a, ta, sa, La = run(n, d, B, False)
    b, tb, sb, Lb = run(n, d, B, True)
    for rep in range(3):
        ef = np.max(np.abs(a[rep][0] - b[rep][0]) / np.abs(a[rep][0]))
        eg = np.max(np.abs(a[rep][1] - b[rep][1])) / np.max(np.abs(a[rep][1]))
        print(f"n={n} d={d} B={B} rep={rep}: nlml rel diff {ef:.2e}, grad rel diff {eg:.2e}", flush=True)
    print(f"   L max abs diff {np.max(np.abs(np.tril(La) - np.tril(Lb))):.2e}; eval ms multi-launch {ta:.3f} vs one-kernel {tb:.3f}; "
          f"cholesky stage {sa[1]:.3f} vs {sb[1]:.3f} ms", flush=True)
