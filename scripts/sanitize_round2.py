"""Round-2 code paths for compute-sanitizer (memcheck): a batch that takes the two-level Cholesky with the merged publisher CTA,
48 x 48 tiles for the symmetric products, per-warp k-ranges, the fused W o K epilogue; device L-BFGS through the multi-device
entry point with handle reuse; metrics; synthetic generation; nearPD."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from gpsurvival_b200 import gp, synth, _lib
rng = np.random.default_rng(0)
n, d, B = 285, 200, 80
pbs = [synth.make_problem(n, 4, d, n // 2, 300 + b) for b in range(2)]
X = np.stack([pbs[b % 2]["trainingData"] for b in range(B)])
y = np.log(np.stack([pbs[b % 2]["trainingTargets"] for b in range(B)]))
ev = np.stack([pbs[b % 2]["events"] for b in range(B)])
fit = gp.Fit(0, batch=B)
fit.set_options("ARD", "Zero", "noiseCorrVec", "intended")
fit.set_data(X, y, ev)
fit.set_noise_corr(np.full((B, int(np.sum(ev[0] == 0))), np.log(0.2)))
par = np.concatenate([np.log([0.2, 0.8]), np.log(1.7 * np.sqrt(d)) * np.ones(d)])[None, :] + 0.05 * rng.standard_normal((B, 2 + d))
for _ in range(3):
    f, g = fit.nlml_grad(par)
print("batch ok", f[0], np.isfinite(g).all())
res = fit.metrics(rng.standard_normal((B, 50)), np.abs(rng.standard_normal((B, 50))) + 0.1, (rng.random((B, 50)) < 0.7).astype(float))
print("metrics ok", res["c.index"][0])
fit.close()
# folds through the multi-device entry point, twice (idle handle reused), then released
pb = synth.make_problem(90, 5, 3, 36, 77)
Xp, yp, evp = pb["trainingData"], np.log(pb["trainingTargets"]), pb["events"]
cens, died = np.where(evp == 0)[0], np.where(evp == 1)[0]
tr = np.array([rng.permutation(np.concatenate([rng.permutation(cens)[:20], rng.permutation(died)[:40]])) for _ in range(6)], dtype=np.int32)
te = np.array([rng.permutation(np.setdiff1d(np.arange(90), r))[:20] for r in tr], dtype=np.int32)
start = np.stack([np.concatenate([np.log([0.2, 0.8]), np.log(1.7 * np.sqrt(3)) + 0.1 * rng.standard_normal(3)]) for _ in range(6)])
for _ in range(2):
    r = gp.fit_folds_multi([0], Xp, yp, evp, tr, te, start, cov_form="ARD", noise_corr="noiseCorrVec", method="L-BFGS", maxit=4,
                           tolerance=1e300, max_count=6, survival=True)
print("folds ok", r["objective"][0], _lib.load().gps_multi_release())
# nearPD on a singular Ky, synthetic generation
one = gp.Fit(0)
Xd = np.vstack([pb["trainingData"][:40], pb["trainingData"][:40]])
one.set_data(Xd, np.concatenate([yp[:40], yp[:40]]), np.concatenate([evp[:40], evp[:40]]))
print("nearPD ok", one.nlml(np.array([-800.0, 0.0, 0.0])), one.near_pd_used())
out = one.synth_generate(7, 64, 3, n_censor=20)
print("synth ok", type(out).__name__)
one.close()
