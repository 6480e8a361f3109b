"""Short program for ncu: the HBM-bound kernels of the path on sizes that fill the GPU — cov_finish_kernel / gram_diag_kernel (split-K
Gram of C4-shaped fits), alpha_kernel, wk_kernel (Matern gradient route), predict_finish_kernel, adjust_kernel.
    ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none \
        -k regex:'cov_finish|gram_diag|alpha_kernel|wk_kernel|predict_finish|adjust_kernel|nlml_finish|rs_reduce|fill_aug' ...
Graphs are off so that every kernel is a launch."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from gpsurvival_b200 import gp, synth

rng = np.random.default_rng(0)
# (1) C4-shaped batch (n = 400, d = 2000): split-K Gram -> gram_diag + cov_finish; alpha; fused W o K (no wk_kernel)
cfg = synth.CONFIGS["C4"]
B = 128
pbs = [synth.make_config_problem(cfg, fit_index=i) for i in range(2)]
n, d = pbs[0]["trainingData"].shape
X = np.stack([pbs[i % 2]["trainingData"] for i in range(B)])
y = np.stack([np.log(pbs[i % 2]["trainingTargets"]) for i in range(B)])
ev = np.stack([pbs[i % 2]["events"] for i in range(B)])
fit = gp.Fit(0, batch=B)
fit.set_graphs(False)
fit.set_options(cfg.cov_form, "Zero", cfg.noise_corr, "intended")
fit.set_data(X, y, ev)
fit.set_noise_corr(np.full((B, int(np.sum(ev[0] == 0))), np.log(0.2)))
par0 = synth.start_par(cfg, d=d)
for i in range(2):
    f, g = fit.nlml_grad(par0[None, :] + 0.05 * rng.standard_normal((B, par0.size)))
print("C4 x", B, "nlml[0] =", f[0])
fit.close()

# (2) Matern 5/2 gradient on C2-shaped fits (n = 2000): wk_kernel (the un-fused W o Kd pass), alpha_kernel at n = 2000; prediction
import bench
B = 32
X, y, ev, lnc, par0 = bench.batch_workload(0, B)
fit = gp.Fit(0, batch=B)
fit.set_graphs(False)
fit.set_options("Matern", "Zero", "noiseCorrVec", "intended", extra_param=5)
fit.set_data(X, y, ev)
fit.set_noise_corr(lnc)
par = par0[None, :3] + 0.05 * rng.standard_normal((B, 3))          # Matern: one length scale
for i in range(2):
    f, g = fit.nlml_grad(par)
print("Matern C2 x", B, "nlml[0] =", f[0])
Xs = rng.standard_normal((B, 4096, X.shape[2]))
out = fit.predict(par, Xs)
print("predict m = 4096:", np.asarray(out[0]).shape)
fit.close()

# (3) adjust: 16 M censored samples (40 c bytes of traffic)
c = 1 << 24
one = gp.Fit(0)
a = rng.uniform(-3, 3, c); mu = rng.standard_normal(c); s = rng.uniform(0.5, 2.0, c)
for i in range(2):
    m, v = one.adjust(a, mu, s)
print("adjust c =", c, "mean[0] =", m[0])
one.close()
