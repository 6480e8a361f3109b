"""Strong scaling of BASELINE.json configs[3] through the C ABI's multi-device entry point (gps_fit_folds_multi): 512 GPS3 ARD
fits = 16 restarts x 32 folds of ONE Yuan-shaped data set, from ONE host process, on 1 / 2 / 4 / 8 GPUs of the box.
    python scripts/sweep_multi.py [total] [method] [ndev ...]
Prints one JSON line per device count; no collective anywhere (LPT partition in the library, one host thread per device)."""
import json, os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from gpsurvival_b200 import gp, synth

total = int(sys.argv[1]) if len(sys.argv) > 1 else 512
method = sys.argv[2] if len(sys.argv) > 2 else "L-BFGS"
have = torch.cuda.device_count()
counts = [int(a) for a in sys.argv[3:]] or [n for n in (1, 2, 4, 8) if n <= have]
cfg = synth.CONFIGS["C4"]
restarts = 16
folds = total // restarts
pb = synth.make_problem(cfg.n + cfg.m, 4, cfg.d, int(round((cfg.n + cfg.m) * cfg.c / cfg.n)), 20160916 + 3)
Xp, yp, evp = pb["trainingData"], np.log(pb["trainingTargets"]), pb["events"]
rng = np.random.default_rng(9)
cens, died = np.where(evp == 0)[0], np.where(evp == 1)[0]
tr, te = [], []
for f in range(folds):
    rows = np.concatenate([rng.permutation(cens)[:cfg.c], rng.permutation(died)[:cfg.n - cfg.c]])
    rest = np.setdiff1d(np.arange(Xp.shape[0]), rows)
    tr.append(rng.permutation(rows))
    te.append(rng.permutation(rest)[:cfg.m])
tr = np.array([tr[t % folds] for t in range(total)], dtype=np.int32)
te = np.array([te[t % folds] for t in range(total)], dtype=np.int32)
starts = np.stack([gp._flatten(synth.start_log_hyp(cfg.d, cfg.cov_form, rng), cfg.d, "Zero", cfg.noise_corr, None) for _ in range(total)])
base = None
ref_obj = None
for nd in counts:
    devs = list(range(nd))
    times = []
    for rep in range(3):                      # first pass: handles, buffers and graphs are created (and, on devices not yet touched,
        t0 = time.perf_counter()              # the context and the kernels are loaded); later passes reuse the idle handle per device
        res = gp.fit_folds_multi(devs, Xp, yp, evp, tr, te, starts, cov_form=cfg.cov_form, noise_corr=cfg.noise_corr,
                                 method=method, maxit=10, tolerance=1e300, max_count=6, survival=True)
        dt = time.perf_counter() - t0
        times.append(dt)
    obj = np.asarray(res["objective"], dtype=np.float64)
    if ref_obj is None:
        ref_obj = obj.copy()
    same = bool(np.array_equal(obj, ref_obj))
    if base is None:
        base = total / dt
    print(json.dumps({"metric": "gps3_fits_per_sec", "n_gpus": nd, "value": total / dt, "unit": "fits/s", "seconds": dt, "seconds_first_call": times[0], "seconds_all_calls": times, "fits": total,
                      "speedup_vs_%dgpu" % counts[0]: (total / dt) / base, "scaling": "strong",
                      "lost_fits": int(np.sum(res["status"] != 0)), "objectives_bitwise_equal_to_first_run": same,
                      "workload": "C4: %d GPS3 ARD fits (%d restarts x %d folds of one n=%d+%d, d=%d data set), %s, 6 GPS rounds x maxit 10; "
                                  "gps_fit_folds_multi from one host process, LPT partition, no collective" % (total, restarts, folds, cfg.n, cfg.m, cfg.d, method)}),
          flush=True)
