import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from gpsurvival_b200 import gp, synth
name = sys.argv[1] if len(sys.argv) > 1 else "C2"
cfg = synth.CONFIGS[name]
Bs = [int(b) for b in sys.argv[2].split(",")] if len(sys.argv) > 2 else [1, 4, 8, 16, 32]
pbs = [synth.make_config_problem(cfg, fit_index=i) for i in range(max(Bs) if max(Bs) <= 8 else 8)]
n, d = pbs[0]["trainingData"].shape
flops = 3.0 * n * n * d + float(n) ** 3
for B in Bs:
    X = np.stack([pbs[i % len(pbs)]["trainingData"] for i in range(B)])
    y = np.stack([np.log(pbs[i % len(pbs)]["trainingTargets"]) for i in range(B)])
    ev = np.stack([pbs[i % len(pbs)]["events"] for i in range(B)])
    c = int(np.sum(ev[0] == 0))
    fit = gp.Fit(0, batch=B)
    fit.set_options(cfg.cov_form, "Zero", cfg.noise_corr, "intended")
    fit.set_data(X, y, ev)
    if cfg.noise_corr == "noiseCorrVec":
        fit.set_noise_corr(np.full((B, c), np.log(0.2)))
    rng = np.random.default_rng(B)
    par = np.tile(synth.start_par(cfg, d=d), (B, 1)) + 0.05 * rng.standard_normal((B, cfg.n_hyp if cfg.d == d else 2 + d))
    for _ in range(3):
        f, g = fit.nlml_grad(par)
    ms = fit.time_eval(1, 5)
    ms0 = fit.time_eval(0, 5)
    st = fit.time_stages()
    print(f"{name} B={B:3d}: nlml+grad {ms:8.3f} ms/batch -> {B/ms*1e3:8.1f} evals/s, {B*flops/ms/1e9:6.2f} TF/s ; nlml only {ms0:8.3f} ms -> {B/ms0*1e3:8.1f} evals/s ; stages", np.array2string(st[:5], precision=2), flush=True)
    fit.close()
