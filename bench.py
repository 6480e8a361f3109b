#!/usr/bin/env python
"""bench.py — GP logML + gradient evaluations / s on BASELINE.json configs[1]
(C2: GPS3, ARD, n=2000, d=20, 50 % censoring), one process per GPU.

    python bench.py --gpus N --steps K --warmup W            # our CUDA path
    python bench.py --impl reference --steps K --warmup W    # the reference algorithm on host cores

One "step" = one pass of the hot path over one BATCH of independent fits: `--batch` (default 256: 74 GB of the
180 GB of HBM; the diagonal-block phases of the Cholesky are latency-bound, so a larger batch amortises them — the same
code does 3 077 / 3 190 / 3 230 evaluations/s at 64 / 128 / 256 fits per step)
C2-shaped GPS3/ARD fits (distinct synthetic data sets and start hyper-parameters — restarts /
CV folds / repetitions, which north_star batches per GPU) each get one evaluation of
ForOptimisationLog + OptimisationGradientLog (NLML and its gradient w.r.t. all 22
hyper-parameters).  The unit of the metric stays the single evaluation: value = batch * K / time.
A lone fit is latency-bound by the sequential pivot chain of its Cholesky; `single_fit` reports it.

* value  : evaluations/s, inputs resident in HBM, device-timed with CUDA events on the
           library's stream (gps_time_eval: K back-to-back graph replays; nothing is cached
           between replays: each recomputes Gram, Cholesky, inverse and gradient), max over ranks.
* e2e    : the same metric through the public API with HOST buffers (pinned, X column-major per fit
           as the ABI and R store it): every step uploads X, y, events, the GPS3 noise-correction
           vectors and par for the whole batch, evaluates, and reads all NLMLs and gradients back.
* roofline: the evaluation graph (one launch = `batch` evaluations, each 3 n^2 d + n^3
           algorithmic FP64 flops) against the FP64 tensor (DMMA) peak measured in
           this run with cublasDgemm 8192^3 (MEASURED_PEAKS.json carries no FP64 figure).
* cpu_baseline: the reference's algorithm restated in numpy (oracle/, as-written operation
           sequence) on the host cores — R itself is not installed in the image.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

from gpsurvival_b200 import synth   # noqa: E402

METRIC = "gp_logml_grad_evals_per_sec"
UNIT = "evals/s"
CFG = synth.CONFIGS["C2"]


def workload(rank: int, fit: int = 0):
    pb = synth.make_config_problem(CFG, fit_index=1000 * rank + fit)
    X = pb["trainingData"]
    y = np.log(pb["trainingTargets"])
    ev = pb["events"]
    lnc = np.full(int(np.sum(ev == 0)), np.log(0.2))
    par0 = synth.start_par(CFG, d=X.shape[1])
    return X, y, ev, lnc, par0


def batch_workload(rank: int, batch: int, distinct: int = 4):
    """`batch` independent fits; `distinct` different synthetic data sets are generated (host time)
    and cycled, every fit gets its own hyper-parameter vector."""
    sets = [workload(rank, f) for f in range(min(distinct, batch))]
    X = np.stack([sets[i % len(sets)][0] for i in range(batch)])
    y = np.stack([sets[i % len(sets)][1] for i in range(batch)])
    ev = np.stack([sets[i % len(sets)][2] for i in range(batch)])
    lnc = np.stack([sets[i % len(sets)][3] for i in range(batch)])
    return X, y, ev, lnc, sets[0][4]


def step_pars(par0, count, seed):
    rng = np.random.default_rng(seed)
    return par0[None, :] + 0.05 * rng.standard_normal((count, par0.size))


def measure_extra_config(name, B, local, reps):
    """One more BASELINE.json shape on the driver's clock (north_star's target shapes): `B` fits of config `name`
    batched on this GPU, NLML+gradient evaluations device-timed like the headline, with per-stage rates and its own
    clock sample."""
    from gpsurvival_b200 import gp
    global CFG
    saved = CFG
    CFG = synth.CONFIGS[name]
    try:
        X, y, ev, lnc, par0 = batch_workload(0, B)
        n, d = X.shape[1], X.shape[2]
        fit = gp.Fit(local, batch=B)
        fit.set_options(CFG.cov_form, "Zero", CFG.noise_corr, "intended")
        fit.set_data(X, y, ev)
        fit.set_noise_corr(lnc)
        rng = np.random.default_rng(23)
        pars = par0[None, None, :] + 0.05 * rng.standard_normal((4, B, par0.size))
        for w in range(3):
            fit.nlml_grad(pars[w])
        sampler = ClockSampler(local)
        sampler.start()
        ms = fit.time_eval(kind=1, reps=reps)
        clocks = sampler.stop()
        fit.time_stages()
        st = fit.time_stages()
        fl = algorithmic_flops(n, d, ard=True)
        n3 = float(n) ** 3 / 3.0
        out = {"workload": "%s: %d fits of n=%d d=%d, GPS3 ARD NLML+gradient" % (name, B, n, d), "evals_per_sec": B * 1e3 / ms,
               "ms_per_step": ms, "tflops": B * fl / (ms * 1e-3) / 1e12, "clocks": clocks,
               "stages_ms": {"gram": st[0], "cholesky": st[1], "tri_inverse": st[2], "kinv": st[3], "wk_grad": st[4]},
               "stage_tflops": {"gram (n^2 d)": B * float(n) * n * d / (st[0] * 1e-3) / 1e12,
                                "cholesky (n^3/3)": B * n3 / (st[1] * 1e-3) / 1e12,
                                "tri_inverse (n^3/3)": B * n3 / (st[2] * 1e-3) / 1e12,
                                "kinv (n^3/3)": B * n3 / (st[3] * 1e-3) / 1e12,
                                "W o K + gradient contraction (2 n^2 (d+1))": B * 2.0 * n * n * (d + 1) / (st[4] * 1e-3) / 1e12}}
        fit.close()
        return out
    finally:
        CFG = saved


def measure_extra_c5(local, n=16384, d=50, m=4096):
    """BASELINE.json configs[4]: one large GPS1 / SqExp fit, n = 16384 — one NLML+gradient evaluation with per-stage rates
    and one prediction at m = 4096 points (inputs are random: sampling GP targets at this n on the host would dominate)."""
    from gpsurvival_b200 import gp
    rng = np.random.default_rng(4)
    X = np.asfortranarray(rng.standard_normal((n, d)))
    Xs = np.asfortranarray(rng.standard_normal((m, d)))
    y = rng.standard_normal(n)
    ev = (rng.random(n) > 0.5).astype(np.float64)
    par = np.log([0.2, 0.8, 1.7 * np.sqrt(d)])
    fit = gp.Fit(local)
    fit.set_options("SqExp", "Zero", False)
    fit.set_data(X, y, ev)
    for i in range(2):
        fit.nlml_grad(par + 1e-3 * i)
    sampler = ClockSampler(local)
    sampler.start()
    ms = fit.time_eval(kind=1, reps=3)
    clocks = sampler.stop()
    st = fit.time_stages()
    n3 = float(n) ** 3 / 3.0
    fit.finalize(par)
    fit.predict(par, Xs)
    pred_ms = fit.last_device_ms()
    pred_flops = 2.0 * n * m * d + float(n) * n * m + 4.0 * n * m          # SURVEY.md §8d "Predict flops"
    out = {"workload": "C5: one GPS1 SqExp fit n=%d d=%d; NLML+gradient evaluation and a prediction at m=%d" % (n, d, m),
           "eval_ms": ms, "tflops": (float(n) * n * d + float(n) ** 3) / (ms * 1e-3) / 1e12, "clocks": clocks,
           "stages_ms": {"gram": st[0], "cholesky": st[1], "tri_inverse": st[2], "kinv": st[3], "wk_grad": st[4]},
           "stage_tflops": {"cholesky (n^3/3)": n3 / (st[1] * 1e-3) / 1e12, "tri_inverse (n^3/3)": n3 / (st[2] * 1e-3) / 1e12,
                            "kinv (n^3/3)": n3 / (st[3] * 1e-3) / 1e12},
           "predict_m4096": {"device_ms": pred_ms, "tflops": pred_flops / (pred_ms * 1e-3) / 1e12,
                             "flops": "2 n m d + n^2 m + 4 n m (cross-covariance, V = L^-1 K*, mean, variance)"}}
    fit.close()
    return out


def measure_extra_c4_sweep(local, total=512, restarts=16, method="L-BFGS"):
    """BASELINE.json configs[3] on one GPU: 512 GPS3 fits = 16 restarts x 32 folds of ONE Yuan-shaped data set, through
    gps_fit_folds_multi (the pool crosses the bus once, fits are row-index views; sub-batches sized to free memory)."""
    from gpsurvival_b200 import gp
    cfg = synth.CONFIGS["C4"]
    folds = total // restarts
    pb = synth.make_problem(cfg.n + cfg.m, 4, cfg.d, int(round((cfg.n + cfg.m) * cfg.c / cfg.n)), 20160916 + 3)
    Xp, yp, evp = pb["trainingData"], np.log(pb["trainingTargets"]), pb["events"]
    rng = np.random.default_rng(9)
    cens, died = np.where(evp == 0)[0], np.where(evp == 1)[0]
    tr, te = [], []
    for f in range(folds):                       # every fold: cfg.c censored + (n - c) uncensored training rows, the rest tests
        rows = np.concatenate([rng.permutation(cens)[:cfg.c], rng.permutation(died)[:cfg.n - cfg.c]])
        rest = np.setdiff1d(np.arange(Xp.shape[0]), rows)
        tr.append(rng.permutation(rows))
        te.append(rng.permutation(rest)[:cfg.m])
    tr = np.array([tr[t % folds] for t in range(total)], dtype=np.int32)
    te = np.array([te[t % folds] for t in range(total)], dtype=np.int32)
    starts = np.stack([gp._flatten(synth.start_log_hyp(cfg.d, cfg.cov_form, rng), cfg.d, "Zero", cfg.noise_corr, None)
                       for _ in range(total)])
    kw = dict(cov_form=cfg.cov_form, noise_corr=cfg.noise_corr, method=method, maxit=10, tolerance=1e300, max_count=6, survival=True)
    # first call: creates the handle (30 GB of per-fit buffers at 512 fits), captures its graphs, loads the kernels of this
    # shape; the library keeps that handle idle for the next call (gps_multi_release frees it), which is the one timed —
    # the steady state of a session that runs sweep after sweep (scripts/sweep_multi.py reports both the same way)
    t0 = time.perf_counter()
    gp.fit_folds_multi([local], Xp, yp, evp, tr, te, starts, **kw)
    dt_first = time.perf_counter() - t0
    later = []
    for _ in range(2):
        t0 = time.perf_counter()
        res = gp.fit_folds_multi([local], Xp, yp, evp, tr, te, starts, **kw)
        later.append(time.perf_counter() - t0)
    dt = min(later)                              # host-driven lock-step loop: run-to-run spread of a few tenths of a second
    from gpsurvival_b200 import _lib
    _lib.load().gps_multi_release()
    return {"workload": "C4: %d GPS3 ARD fits (%d restarts x %d folds of one n=%d+%d, d=%d data set), %s, 6 GPS rounds x maxit 10"
                        % (total, restarts, folds, cfg.n, cfg.m, cfg.d, method),
            "fits_per_sec": total / dt, "seconds": dt, "lost_fits": int(np.sum(res["status"] != 0)),
            "seconds_first_call": dt_first, "seconds_later_calls": later, "batched_evaluations": int(res["ticks"]),
            "h2d_note": "pool uploaded once per sub-batch; fits, censored rows and test folds are row-index views"}


CONFIG_NOTES = {"C2": "BASELINE.json configs[1]", "C3": "BASELINE.json configs[2], Tothill shape",
                "C4": "BASELINE.json configs[3], Yuan shape"}


def config_dict(B, world, n, d, nhyp):
    return {"workload": "%s: GPS3 ARD NLML+gradient, n=%d d=%d c=%d (%s); one step = "
                        "one evaluation of each of %d independent fits batched per GPU"
                        % (CFG.name, n, d, CFG.c, CONFIG_NOTES.get(CFG.name, ""), B),
            "n": int(n), "d": int(d), "hyperparameters": int(nhyp), "fits_per_gpu": int(B),
            "partition": f"independent fits, {B} per GPU x {world} GPUs, no collective",
            "l2": "per-step working set %d fits x (5 n x n matrices + X and X') = %.1f GB >> 126 MB L2 (no flush needed)"
                  % (B, B * (5 * n * n + 2 * n * d) * 8 / 1e9)}


def measured_traffic(B):
    """DRAM bytes of one step from the committed ncu capture of the same workload (bench.py cannot run
    under ncu itself); None when no capture exists for this batch size."""
    import glob
    names = sorted((os.path.basename(f) for f in glob.glob(os.path.join(ROOT, "profiles", "r*_traffic_%s_batch*.json" % CFG.name.lower()))),
                   reverse=True)                      # rNx names sort by round, then letter: newest first
    for name in names:
        try:
            t = json.load(open(os.path.join(ROOT, "profiles", name)))
            if int(t["fits_per_step"]) == int(B):
                return float(t["traffic_bytes_per_step"]), name
        except (OSError, ValueError, KeyError):
            continue
    return None, None


def algorithmic_flops(n, d, ard=True):
    return (3.0 if ard else 1.0) * n * n * d + float(n) ** 3


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled during the timed region (profiling guide)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.idx = gpu_index
        self.rows = []
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-i", str(self.idx), "-lms", "100"], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[1]))
                mx.append(float(r[2]))
            except (ValueError, IndexError):
                continue
            for k, nm in enumerate(names):
                if len(r) > 5 + k and r[5 + k].lower().startswith("active"):
                    reasons.add(nm)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": float(max(mx)) if mx else None,
                "samples": len(sm), "reasons": sorted(reasons)}


def measure_fp64_peak(device):
    """cublasDgemm 8192^3 (via torch.matmul, float64): best of 5 — the FP64 tensor-peak denominator."""
    import torch
    n = 8192
    a = torch.randn(n, n, dtype=torch.float64, device=device)
    b = torch.randn(n, n, dtype=torch.float64, device=device)
    torch.matmul(a, b)
    torch.cuda.synchronize(device)
    best = 1e30
    for _ in range(5):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        torch.matmul(a, b)
        e1.record()
        torch.cuda.synchronize(device)
        best = min(best, e0.elapsed_time(e1))
    del a, b
    torch.cuda.empty_cache()
    return 2.0 * n ** 3 / (best * 1e-3) / 1e12


def cpu_eval(par, X, y, ev, lnc):
    """One NLML + gradient evaluation the way the reference computes it (numpy restatement)."""
    from oracle import gp_oracle as orc
    kw = dict(cov_form=CFG.cov_form, mean_form="Zero", noise_corr=CFG.noise_corr, log_hyp_noise_corr=lnc)
    f = orc.nlml(par, X, y, ev, as_written_ops=True, **kw)
    g = orc.nlml_grad(par, X, y, ev, **kw)
    return f, g


def host_cores():
    try:
        return len(os.sched_getaffinity(0))
    except AttributeError:
        return os.cpu_count() or 1


def pin_blas_threads(nthreads=None):
    """Set the BLAS pool of THIS process (torchrun exports OMP_NUM_THREADS=1, a plain shell does not: the arm must not
    depend on how it was launched) and return the count the pool really has afterwards."""
    nthreads = nthreads or host_cores()
    try:
        from threadpoolctl import threadpool_info, threadpool_limits
        threadpool_limits(limits=nthreads, user_api="blas")
        th = [p.get("num_threads", 1) for p in threadpool_info() if p.get("user_api") == "blas"]
        return max(th) if th else 1
    except Exception:
        return 1


_WORKER = r"""
import os, sys, time, json
import numpy as np
sys.path.insert(0, sys.argv[1])
from oracle import gp_oracle as orc
z = np.load(sys.argv[2])
kw = dict(cov_form=str(z["cov_form"]), mean_form="Zero", noise_corr=str(z["noise_corr"]), log_hyp_noise_corr=z["lnc"])
warm, cnt, budget = int(sys.argv[3]), int(sys.argv[4]), float(sys.argv[5])
def ev(p):
    orc.nlml(p, z["X"], z["y"], z["ev"], as_written_ops=True, **kw)
    orc.nlml_grad(p, z["X"], z["y"], z["ev"], **kw)
for w in range(warm):
    ev(z["pars"][w % len(z["pars"])])
print("READY", flush=True)
sys.stdin.readline()                      # all workers start their timed region together
t0 = time.perf_counter(); done = 0
while done < cnt and (done < 1 or time.perf_counter() - t0 < budget):
    ev(z["pars"][(warm + done) % len(z["pars"])]); done += 1
print(json.dumps({"evals": done, "seconds": time.perf_counter() - t0}), flush=True)
"""


def cpu_concurrent(X, y, ev, lnc, pars, workers, warm, cnt, budget_s):
    """The reference algorithm with ONE independent fit per host core (one single-threaded process each) — how a
    host would run restarts / folds / repetitions side by side.  Returns (evals/s over all workers, workers, evals)."""
    import tempfile
    tmp = tempfile.mkdtemp(prefix="gps_cpu_")
    path = os.path.join(tmp, "w.npz")
    np.savez(path, X=X, y=y, ev=ev, lnc=lnc, pars=np.asarray(pars), cov_form=CFG.cov_form, noise_corr=str(CFG.noise_corr))
    env = dict(os.environ, OMP_NUM_THREADS="1", OPENBLAS_NUM_THREADS="1", MKL_NUM_THREADS="1")
    procs = [subprocess.Popen([sys.executable, "-c", _WORKER, ROOT, path, str(warm), str(cnt), str(budget_s)], env=env,
                              stdin=subprocess.PIPE, stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
             for _ in range(workers)]
    try:
        for p in procs:
            if p.stdout.readline().strip() != "READY":
                raise RuntimeError("cpu worker failed to start")
        t0 = time.perf_counter()
        for p in procs:
            p.stdin.write("go\n")
            p.stdin.flush()
        res = [json.loads(p.stdout.readline()) for p in procs]
        wall = time.perf_counter() - t0
    finally:
        for p in procs:
            try:
                p.stdin.close()
                p.wait(timeout=5)
            except Exception:
                p.kill()
        try:
            os.remove(path)
            os.rmdir(tmp)
        except OSError:
            pass
    total = sum(r["evals"] for r in res)
    return total / wall, workers, total


def run_reference(args):
    """The reference's algorithm (numpy restatement, as-written operation sequence) on the host cores, two ways:
    (a) one process with the BLAS pool over all cores, (b) one single-threaded process per core, each evaluating its
    own fit — how a host runs independent restarts / folds side by side.  `value` is the better of the two (all the
    host threads the path can use); both are reported.  A step of (b) = one evaluation on every core at once."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    X, y, ev, lnc, par0 = workload(0)
    cores = host_cores()
    blas = pin_blas_threads(cores)
    pars = step_pars(par0, max(args.steps, 3) + args.warmup + 1, 7)
    t_w = time.perf_counter()
    for w in range(max(args.warmup, 1)):
        cpu_eval(pars[w], X, y, ev, lnc)
    per = (time.perf_counter() - t_w) / max(args.warmup, 1)
    steps_a = max(3, min(args.steps, int(45.0 / per)))          # bounded: ~45 s for arm (a)
    t0 = time.perf_counter()
    for s in range(steps_a):
        cpu_eval(pars[args.warmup + s], X, y, ev, lnc)
    dt_a = time.perf_counter() - t0
    val_a = steps_a / dt_a
    try:
        val_b, workers, total_b = cpu_concurrent(X, y, ev, lnc, pars, cores, min(max(args.warmup, 1), 2), args.steps, 75.0)
    except Exception as exc:    # a host that cannot spawn workers still reports arm (a)
        val_b, workers, total_b = 0.0, 0, 0
        print("cpu_concurrent failed: %r" % (exc,), file=sys.stderr)
    if val_b >= val_a:
        val, used, steps = val_b, workers, max(1, total_b // max(workers, 1))
        ms_step = 1e3 * workers / val_b
        sample = (f"{total_b} full evaluations of the C2 workload (n={X.shape[0]}, d={X.shape[1]}): {workers} single-threaded "
                  f"processes side by side, one fit each; one step = one evaluation on every core")
    else:
        val, used, steps, ms_step = val_a, blas, steps_a, 1e3 * dt_a / steps_a
        sample = f"{steps_a} full evaluations of the C2 workload (n={X.shape[0]}, d={X.shape[1]}), one process, {blas} BLAS threads"
    line = {
        "impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus, "steps": steps,
        "warmup": args.warmup, "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": dict(config_dict(args.batch, args.gpus, X.shape[0], X.shape[1], par0.size),
                       reference_step="the host evaluates the fits of a batch on its cores; a timed step here is a bounded "
                                      "sample of the batch (see cpu_baseline.sample), value is evaluations/s over the whole host"),
        "cpu_baseline": {"value": val, "unit": UNIT, "cores": used, "kind": "port", "sample": sample,
                         "host_cores": cores,
                         "one_process_blas": {"value": val_a, "threads": blas, "evaluations": steps_a},
                         "concurrent_fits": {"value": val_b, "processes": workers, "evaluations": total_b},
                         "note": "reference algorithm restated in numpy (as-written op sequence); R unavailable in image"},
        "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)
    return 0


def run_gpu(args):
    import torch
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    dist = None
    if world > 1:
        import torch.distributed as dist
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        torch.cuda.set_device(local)
        # NCCL prints its version banner on stdout when the communicator is created; stdout must carry
        # exactly one JSON line, so stdout points at stderr until the first collective has run
        sys.stdout.flush()
        saved_fd = os.dup(1)
        os.dup2(2, 1)
        try:
            dist.init_process_group("nccl", device_id=torch.device("cuda", local))
            dist.barrier()
            torch.cuda.synchronize()
        finally:
            sys.stdout.flush()
            os.dup2(saved_fd, 1)
            os.close(saved_fd)
    device = torch.device("cuda", local)
    torch.cuda.set_device(device)

    from gpsurvival_b200 import gp
    B = args.batch
    X, y, ev, lnc, par0 = batch_workload(rank, B)
    n, d = X.shape[1], X.shape[2]
    flops = algorithmic_flops(n, d, ard=True)

    fit = gp.Fit(local, batch=B)
    fit.set_options(CFG.cov_form, "Zero", CFG.noise_corr, "intended")
    fit.set_data(X, y, ev)
    fit.set_noise_corr(lnc)
    K, W = args.steps, max(args.warmup, 3)
    rng = np.random.default_rng(11 + rank)
    pars = par0[None, None, :] + 0.05 * rng.standard_normal((K + W, B, par0.size))
    for w in range(W):                                   # warm-up: direct run, graph capture, replay
        fit.nlml_grad(pars[w])

    def barrier():
        torch.cuda.synchronize(device)
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize(device)

    def max_over_ranks(v):
        if dist is None:
            return v
        t = torch.tensor([v], dtype=torch.float64, device=device)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    # ---- FP64 tensor peak of this GPU (roofline denominator), rank 0 only to keep the run short
    peak_tf = measure_fp64_peak(device) if rank == 0 else 0.0

    # ---- device-resident throughput: K graph replays, CUDA events on the library's stream --------
    sampler = ClockSampler(local)
    barrier()
    sampler.start()
    l0 = fit.launch_count()
    ms_per_step = fit.time_eval(kind=1, reps=K)
    launches = fit.launch_count() - l0
    barrier()
    dev_ms = max_over_ranks(ms_per_step * K)
    clocks = sampler.stop()
    value = world * B * K / (dev_ms * 1e-3)

    # ---- end to end through the public API with host buffers --------------------------------------
    # inputs live in PINNED host memory, X in the layout the C ABI (and R) uses: every fit column-major
    def pinned(a):
        t = torch.empty(a.shape, dtype=torch.float64, pin_memory=True)
        v = t.numpy()
        v[...] = a
        return t, v
    keep = [pinned(np.ascontiguousarray(np.transpose(X, (0, 2, 1)))), pinned(y), pinned(ev), pinned(lnc), pinned(pars)]
    Xp, yp, evp, lncp, parsp = (k[1] for k in keep)
    fit.set_data(Xp, yp, evp, colmajor=True)          # untimed: registers nothing, only warms the path
    barrier()
    t0 = time.perf_counter()
    for s in range(K):
        fit.set_data(Xp, yp, evp, colmajor=True)
        fit.set_noise_corr(lncp)
        f, g = fit.nlml_grad(parsp[W + s])
    torch.cuda.synchronize(device)
    e2e_s = max_over_ranks(time.perf_counter() - t0)
    barrier()
    e2e_val = world * B * K / e2e_s
    h2d = 8 * (X.size + y.size + ev.size + lnc.size + B * par0.size)
    d2h = 8 * B * (1 + par0.size) + 4 * B

    line = None
    if rank == 0 and args.quick:          # developer A/B runs: the device-timed value and the stage timers, nothing else
        fit.time_stages()
        stages = fit.time_stages()
        one = gp.Fit(local)
        one.set_options(CFG.cov_form, "Zero", CFG.noise_corr, "intended")
        one.set_data(X[0], y[0], ev[0])
        one.set_noise_corr(lnc[0])
        for w in range(3):
            one.nlml_grad(pars[w, 0])
        single_ms, single_nlml_ms = one.time_eval(kind=1, reps=20), one.time_eval(kind=0, reps=20)
        one.close()
        print(json.dumps({"quick": True, "config": CFG.name, "batch": B, "value": value, "ms_per_step": dev_ms / K,
                          "tflops": B * flops / (ms_per_step * 1e-3) / 1e12, "peak_tf": peak_tf, "e2e": e2e_val,
                          "stages_ms": [float(v) for v in stages[:5]], "single_ms": single_ms,
                          "single_nlml_ms": single_nlml_ms, "launches_per_step": launches // max(K, 1),
                          "env": {k: v for k, v in os.environ.items() if k.startswith("GPS_")}}), flush=True)
    elif rank == 0:
        fit.time_stages()                 # first pass pays one-time costs of kernel variants only this (ungrouped) path uses
        stages = fit.time_stages()
        gemm_ms = fit.bench_gemm(8192, 8192, 8192, 3)
        achieved = B * flops / (ms_per_step * 1e-3) / 1e12
        traffic, traffic_file = measured_traffic(B)
        n3 = float(n) ** 3 / 3.0

        def tf(fl, ms):
            return B * fl / (ms * 1e-3) / 1e12 if ms > 0 else None
        # per-stage algorithmic rates of the same step (CUDA events between the stages, gps_time_stages)
        stage_rates = {
            "gram (dmma + covariance epilogue; n^2 d)": tf(float(n) * n * d, stages[0]),
            "cholesky (panel_kernel + dmma trailing updates; n^3/3)": tf(n3, stages[1]),
            "triangular inverse (dmma, U and Linv; n^3/3)": tf(n3, stages[2]),
            "Kinv = U U' (dmma_ws_kernel, the longest launch; n^3/3)": tf(n3, stages[3]),
            "W o K + gradient contraction (2 n^2 (d+1))": tf(2.0 * n * n * (d + 1), stages[4]),
        }
        # a lone fit (latency-bound): same workload, batch of one
        one = gp.Fit(local)
        one.set_options(CFG.cov_form, "Zero", CFG.noise_corr, "intended")
        one.set_data(X[0], y[0], ev[0])
        one.set_noise_corr(lnc[0])
        for w in range(3):
            f1, g1 = one.nlml_grad(pars[w, 0])
        single_ms = one.time_eval(kind=1, reps=20)
        single_nlml_ms = one.time_eval(kind=0, reps=20)
        one.close()
        # full GPS3 fits / s: fixed protocol of SURVEY.md §8d (6 outer GPS rounds, the reference's minimum,
        # x one optim() run of maxitSurvival = 10 + test and training predictions), B fits in lock-step
        fits = {}
        try:
            rs = np.random.default_rng(5)
            starts = [synth.start_log_hyp(d, CFG.cov_form, rs) for _ in range(B)]
            par_start = np.stack([gp._flatten(s0, d, "Zero", CFG.noise_corr, None) for s0 in starts])
            Xtest = X[:, :500, :]
            # a non-positive-definite point is rejected (+Inf) in this throughput protocol: repairing it the reference's way
            # (nearPD, a full eigen-decomposition per Higham iteration) costs seconds at n = 2000, on the GPU as in R
            fit.set_near_pd(False)
            # Nelder-Mead needs one evaluation per simplex vertex (hyper-parameters + 1) before it moves: only sensible at d = 20
            for method in (("Nelder-Mead", "BFGS", "L-BFGS") if CFG.name == "C2" else ("BFGS", "L-BFGS")):
                tt = time.perf_counter()
                rr = fit.fit_batch(X, y, ev, par_start, Xtest, method=method, maxit=10, tolerance=1e300, max_count=6,
                                   survival=True)                      # whole fits behind the C ABI (gps_fit_batch)
                dtf = time.perf_counter() - tt
                fits[method] = {"fits_per_sec": B / dtf, "seconds": dtf, "batched_evaluations": int(rr["ticks"]),
                                "rounds": int(rr["count"][0]), "mean_objective": float(np.mean(rr["objective"]))}
            fit.set_near_pd(True)
            fit.set_options(CFG.cov_form, "Zero", CFG.noise_corr, "intended")
            fit.set_data(X, y, ev)
            fit.set_noise_corr(lnc)
            f, g = fit.nlml_grad(pars[W + K - 1])
        except Exception as exc:        # never lose the headline line because of the extra measurement
            fits = {"error": repr(exc)}
        # north_star's other target shapes on the same clock (each with its own clock sample); never lose the headline
        extras = {}
        fit.close()
        fit = None
        if not args.no_extras:
            order = os.environ.get("BENCH_EXTRAS_ORDER", "C3x32,C4x256,C5,C4_sweep_512_fits").split(",")   # developer knob
            fns = {"C3x32": lambda: measure_extra_config("C3", 32, local, 10),
                   "C4x256": lambda: measure_extra_config("C4", 256, local, 5),
                   "C5": lambda: measure_extra_c5(local),
                   "C4_sweep_512_fits": lambda: measure_extra_c4_sweep(local)}
            for key, fn in ((k, fns[k.split("#")[0]]) for k in order):
                try:
                    extras[key] = fn()
                except Exception as exc:
                    extras[key] = {"error": repr(exc)}
        # bounded CPU sample of the same workload (about 10-30 s of host work)
        blas_threads = pin_blas_threads()
        t0 = time.perf_counter()
        cnt = 0
        f_cpu = g_cpu = None
        while cnt < 3 or (time.perf_counter() - t0 < 10.0 and cnt < 40):
            f_cpu, g_cpu = cpu_eval(pars[W + K - 1, 0], X[0], y[0], ev[0], lnc[0])
            cnt += 1
        cpu_val = cnt / (time.perf_counter() - t0)
        try:      # the honest whole-host reading: one single-threaded process per core, one fit each
            conc_val, conc_workers, conc_total = cpu_concurrent(X[0], y[0], ev[0], lnc[0], pars[W:, 0], host_cores(), 1, 2, 12.0)
        except Exception as exc:
            conc_val, conc_workers, conc_total = None, 0, 0
            print("cpu_concurrent failed: %r" % (exc,), file=sys.stderr)
        parity = {"nlml_rel": abs(f[0] - f_cpu) / abs(f_cpu),
                  "grad_rel": float(np.max(np.abs(g[0] - g_cpu)) / np.max(np.abs(g_cpu)))}
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": W,
            "ms_per_step": dev_ms / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic",
            "config": config_dict(B, world, n, d, par0.size),
            "e2e": {"value": e2e_val, "unit": UNIT, "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h),
                    "ms_per_step": 1e3 * e2e_s / K},
            "gpu_launches": int(launches),
            "clocks": clocks,
            "roofline": {"bound": "tensor", "achieved": achieved, "peak": peak_tf, "unit": "TFLOP/s",
                         "frac": achieved / peak_tf if peak_tf else None, "traffic": traffic,
                         "traffic_source": ("profiles/%s: ncu dram__bytes_read.sum + dram__bytes_write.sum summed over "
                                            "the launches of one step (bytes)" % traffic_file) if traffic_file else None,
                         "stage_tflops": stage_rates,
                         "dmma_issue_peak": {"tflops": 37.1, "frac": achieved / 37.1,
                                             "source": "register-only mma.sync.m8n8k4.f64 issue rate measured on B200 "
                                                       "(lab/dmma_lab.cu, profiles/r1d_dmma_lab.log); cublasDgemm reaches 35.9"},
                         "kernel": "evaluation graph (1 launch = %d NLML+gradient evaluations, %d kernel nodes)"
                                   % (B, launches // max(K, 1)),
                         "algorithmic_flops_per_launch": B * flops,
                         "peak_source": "cublasDgemm 8192^3 via torch.matmul(float64), best of 5, measured in this run "
                                        "(MEASURED_PEAKS.json has no FP64 figure)"},
            "dmma_gemm_8192": {"tflops": 2.0 * 8192 ** 3 / (gemm_ms * 1e-3) / 1e12,
                               "frac_of_cublas_dgemm": (2.0 * 8192 ** 3 / (gemm_ms * 1e-3) / 1e12) / peak_tf if peak_tf else None},
            "stages_ms": {"gram": stages[0], "cholesky": stages[1], "tri_inverse": stages[2], "kinv": stages[3],
                          "wk_grad": stages[4], "total": stages[7]},
            "gps3_fits": dict(fits, protocol="gps_fit_batch (C ABI): 6 GPS rounds x optim(maxitSurvival=10) + test/train "
                                             "predictions, %d fits in lock-step on one GPU (SURVEY.md §8d); non-positive-definite "
                                             "points rejected as +Inf (gps_set_near_pd off: one nearPD repair at n = 2000 costs seconds)" % B),
            "single_fit": {"nlml_grad_ms": single_ms, "nlml_ms": single_nlml_ms, "evals_per_sec": 1e3 / single_ms,
                           "tflops": flops / (single_ms * 1e-3) / 1e12},
            "cpu_baseline": {"value": cpu_val, "unit": UNIT, "cores": blas_threads, "kind": "port",
                             "sample": f"{cnt} full evaluations of one fit of the same {CFG.name} workload, one process, "
                                       f"{blas_threads} BLAS threads",
                             "host_cores": host_cores(),
                             "concurrent_fits": {"value": conc_val, "processes": conc_workers, "evaluations": conc_total,
                                                 "sample": "one single-threaded process per host core, one independent fit each"},
                             "note": "reference algorithm restated in numpy (as-written op sequence); R unavailable in image"},
            "parity_vs_oracle": parity,
            "extras": extras,
        }
    if fit is not None:
        fit.close()
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()
    if line is not None:
        print(json.dumps(line), flush=True)
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=30)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--batch", type=int, default=0, help="independent fits evaluated per step on each GPU "
                                                         "(default: 256 for C2, 32 for C3, 256 for C4)")
    ap.add_argument("--no-extras", action="store_true", help="skip the C3 / C4 / C5 extras of the JSON line")
    ap.add_argument("--quick", action="store_true", help="developer A/B runs: device-timed value + stage timers only")
    ap.add_argument("--config", default="C2", choices=["C2", "C3", "C4"],
                    help="workload shape; the default C2 is the configuration BASELINE.json's metric is quoted on")
    args = ap.parse_args()
    global CFG
    CFG = synth.CONFIGS[args.config]
    if args.batch <= 0:
        args.batch = {"C2": 256, "C3": 32, "C4": 256}[args.config]
    if args.impl == "reference":
        return run_reference(args)
    return run_gpu(args)


if __name__ == "__main__":
    sys.exit(main())
