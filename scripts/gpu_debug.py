import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from gpsurvival_b200 import gp, synth
from oracle import gp_oracle as orc

def covcase(n1, n2, d, ks):
    os.environ["GPS_KSPLIT"] = str(ks)
    fit = gp.Fit(0)
    rng = np.random.default_rng(1)
    x1 = rng.standard_normal((n1, d)); x2 = rng.standard_normal((n2, d))
    ell = np.array([1.7*np.sqrt(d)])
    K = fit.cov(x1, x2, 0.8, ell, "SqExp", "intended")
    Ko = orc.cov_func(x1, x2, 0.8, ell, "SqExp")
    E = np.abs(K-Ko)
    i, j = np.unravel_index(np.argmax(E), E.shape)
    print(f"cov n1={n1} n2={n2} d={d} ksplit={ks}: max err {E.max():.3e} at ({i},{j}); col-max err by 8-col group:",
          np.array2string(E.max(axis=0).reshape(-1, 8).max(axis=1), precision=1), flush=True)
    print("   row-max by 8:", np.array2string(E.max(axis=1)[:n1//8*8].reshape(-1, 8).max(axis=1), precision=1), flush=True)
    fit.close()

# first performance numbers
fit = gp.Fit(0)
for (M, N, K) in [(4096,4096,4096),(8192,8192,8192),(16384,16384,256),(2048,2048,2048)]:
    ms = fit.bench_gemm(M, N, K, 3)
    print(f"dmma gemm {M}x{N}x{K}: {ms:.3f} ms  {2.0*M*N*K/ms/1e9:.1f} TF/s", flush=True)
fit.close()
for name in ("C2", "C3", "C1"):
    cfg = synth.CONFIGS[name]
    pb = synth.make_config_problem(cfg)
    X, y, ev = pb["trainingData"], np.log(pb["trainingTargets"]), pb["events"]
    fit = gp.Fit(0)
    fit.set_options(cfg.cov_form, "Zero", cfg.noise_corr, "intended")
    fit.set_data(X, y, ev)
    if cfg.noise_corr == "noiseCorrVec":
        fit.set_noise_corr(np.full(int(np.sum(ev == 0)), np.log(0.2)))
    par = synth.start_par(cfg, d=X.shape[1])
    for g in (0, 1):
        fit.set_graphs(bool(g))
        for _ in range(3):
            f, gr = fit.nlml_grad(par)
        t = time.perf_counter()
        for _ in range(5):
            f, gr = fit.nlml_grad(par + 1e-9)
            f, gr = fit.nlml_grad(par)
        te = (time.perf_counter() - t) / 10
        print(f"{name} n={X.shape[0]} d={X.shape[1]} graphs={g}: nlml={f:.6f} e2e {te*1e3:.3f} ms/eval; device-only nlml {fit.time_eval(0,10):.3f} ms, nlml+grad {fit.time_eval(1,10):.3f} ms; launches={fit.launch_count()}", flush=True)
    print("   stages [gram, chol, trtri, kinv, wk+grad, -, -, total]:", np.array2string(fit.time_stages(), precision=3), flush=True)
    fit.close()
