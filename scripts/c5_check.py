"""C5 (n=16384, d=50, SqExp GPS1): properties + stage timings of one NLML+gradient evaluation, a
finalize and a prediction (m=4096).  Inputs are random (no GP sampling: the CPU Cholesky of a
20480^2 matrix would dominate the script)."""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from gpsurvival_b200 import gp
n = int(sys.argv[1]) if len(sys.argv) > 1 else 16384
d, m = 50, 4096
rng = np.random.default_rng(4)
X = np.asfortranarray(rng.standard_normal((n, d)))
Xs = np.asfortranarray(rng.standard_normal((m, d)))
y = rng.standard_normal(n)
ev = (rng.random(n) > 0.5).astype(np.float64)
par = np.log([0.2, 0.8, 1.7 * np.sqrt(d)])
fit = gp.Fit(0)
fit.set_options("SqExp", "Zero", False)
t = time.time(); fit.set_data(X, y, ev); print(f"set_data {time.time()-t:.2f}s", flush=True)
for i in range(2):
    t = time.time(); f, g = fit.nlml_grad(par + 1e-3 * i); print(f"nlml_grad call {i}: {time.time()-t:.3f}s nlml={f:.6f} grad={g}", flush=True)
st = fit.time_stages()
fl = dict(gram=n * n * d, chol=n ** 3 / 3, trtri=n ** 3 / 3, kinv=n ** 3 / 3, wk=2.0 * n * n * (d + 1))
for nm, ms, f_ in zip(("gram", "chol", "trtri", "kinv", "wk"), st[:5], fl.values()):
    print(f"  stage {nm:6s} {ms:9.3f} ms  {f_/ms/1e9:6.2f} TF/s (algorithmic)")
print(f"  total {st[7]:.3f} ms -> {(n*n*d + n**3)/st[7]/1e9:.2f} TF/s algorithmic (SqExp: n^2 d + n^3)")
# directional derivative check
v = rng.standard_normal(3); v /= np.linalg.norm(v); h = 1e-4
fd = (fit.nlml(par + h * v) - fit.nlml(par - h * v)) / (2 * h)
f, g = fit.nlml_grad(par)
print(f"directional derivative: fd={fd:.8f} analytic={g @ v:.8f} rel={abs(fd - g @ v)/abs(fd):.2e}")
t = time.time(); fin = fit.finalize(par); print(f"finalize {time.time()-t:.3f}s logMarLik={fin['logMarLik']:.6f}")
alpha = fin["alpha"]
# residual check Ky alpha = y on a row sample using the library's own covariance
idx = rng.choice(n, 64, replace=False)
Krows = fit.cov(X[idx], X, 0.8, [1.7 * np.sqrt(d)], "SqExp")
res = Krows @ alpha + 0.2 * alpha[idx] - y[idx]
print(f"max |Ky alpha - y| on 64 rows = {np.max(np.abs(res)):.3e}")
t = time.time(); mu, vf, vd = fit.predict(par, Xs); print(f"predict m={m}: {time.time()-t:.3f}s (device {fit.last_device_ms():.2f} ms) varf in [{vf.min():.4f},{vf.max():.4f}]")
Ks = fit.cov(X, Xs[:32], 0.8, [1.7 * np.sqrt(d)], "SqExp")
print(f"predict mean vs K*'alpha (32 cols): {np.max(np.abs(mu[:32] - Ks.T @ alpha)):.3e}")
