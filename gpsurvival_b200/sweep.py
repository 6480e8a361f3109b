"""Sweep of independent whole fits over the GPUs of one node (BASELINE.json configs[3]).

    python -m gpsurvival_b200.sweep --config C4 --fits 512 --method BFGS            # 1 GPU
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 \
           -m gpsurvival_b200.sweep --config C4 --fits 512 --method BFGS            # 8 GPUs

One process per GPU.  The fits (restarts x CV folds of a C4-shaped data set: distinct synthetic data
sets, perturbed start hyper-parameters) are partitioned over the ranks by estimated cost
(`partition.partition`), each rank runs its shard as lock-step batches on its own GPU
(`batchfit.apply_gp_batch`), and the per-fit results are gathered on the host — no collective on the
data path.  Prints one JSON line: fits/s over the slowest rank.
"""
from __future__ import annotations

import argparse
import json
import os
import time

import numpy as np


def main(argv=None):
    ap = argparse.ArgumentParser()
    ap.add_argument("--config", default="C4")
    ap.add_argument("--fits", type=int, default=512)
    ap.add_argument("--batch", type=int, default=128, help="fits advanced in lock-step per GPU")
    ap.add_argument("--method", default="BFGS", choices=["BFGS", "Nelder-Mead"])
    ap.add_argument("--rounds", type=int, default=6)
    ap.add_argument("--distinct", type=int, default=4, help="distinct synthetic data sets generated per rank")
    ap.add_argument("--driver", default="c", choices=["c", "python"],
                    help="c: whole fits behind the C ABI (gps_fit_batch); python: the lock-step host driver (batchfit)")
    args = ap.parse_args(argv)

    import torch
    from . import batchfit, gp, partition, synth

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    dist = None
    if world > 1:
        import torch.distributed as dist
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        torch.cuda.set_device(local)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    cfg = synth.CONFIGS[args.config]
    costs = [partition.estimated_cost(cfg.n, cfg.d, cfg.cov_form == "ARD")] * args.fits
    mine = partition.shard_indices(costs, rank, world)

    base = [synth.make_config_problem(cfg, fit_index=1000 * rank + i) for i in range(min(args.distinct, max(len(mine), 1)))]
    d = base[0]["trainingData"].shape[1]
    rng = np.random.default_rng(7 + rank)
    params = dict(modelType="survival", noiseCorr=cfg.noise_corr, optimType=args.method, maxit=100, maxitSurvival=10,
                  meanFuncForm="Zero", covFuncForm=cfg.cov_form, tolerance=1e300, maxCount=args.rounds,
                  burnIn=False, imposePriors=False)

    def barrier():
        torch.cuda.synchronize(local)
        if dist is not None:
            dist.barrier()

    results = {}
    fit = None
    barrier()
    t0 = time.perf_counter()
    for lo in range(0, len(mine), args.batch):
        idx = mine[lo:lo + args.batch]
        if fit is None or fit.batch != len(idx):
            if fit is not None:
                fit.close()
            fit = gp.Fit(local, batch=len(idx))
        probs = [base[i % len(base)] for i in idx]
        starts = [synth.start_log_hyp(d, cfg.cov_form, rng) for _ in idx]
        if args.driver == "python":
            res = batchfit.apply_gp_batch(probs, dict(params, logHypStart=starts), fit=fit)
            for i, r in zip(idx, res):
                results[i] = {"objective": r["objective"], "logMarLik": r["logMarLik"], "rounds": r["count"],
                              "pred_mean": float(np.mean(r["funcMeanPred"]))}
        else:
            fit.set_options(cfg.cov_form, "Zero", cfg.noise_corr, "intended")
            X = np.stack([p["trainingData"] for p in probs])
            y = np.stack([np.log(p["trainingTargets"]) for p in probs])
            ev = np.stack([p["events"] for p in probs])
            Xt = np.stack([p["testData"] for p in probs])
            par0 = np.stack([gp._flatten(s0, d, "Zero", cfg.noise_corr, None) for s0 in starts])
            rr = fit.fit_batch(X, y, ev, par0, Xt, method=args.method, maxit=10, tolerance=1e300, max_count=args.rounds,
                               survival=True)
            for k, i in enumerate(idx):
                results[i] = {"objective": float(rr["objective"][k]), "logMarLik": float(rr["logMarLik"][k]),
                              "rounds": int(rr["count"][k]), "pred_mean": float(np.mean(rr["funcMeanPred"][k]))}
    torch.cuda.synchronize(local)
    mine_s = time.perf_counter() - t0
    barrier()
    full = partition.gather_results(results, world)          # host-side concatenation: the only exchange
    total_s = time.perf_counter() - t0
    if dist is not None:
        tt = torch.tensor([mine_s], dtype=torch.float64, device=f"cuda:{local}")
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        mine_s = float(tt.item())
    if rank == 0:
        assert sorted(full) == list(range(args.fits))
        print(json.dumps({"metric": "gps3_fits_per_sec", "value": args.fits / mine_s, "unit": "fits/s", "n_gpus": world,
                          "fits": args.fits, "seconds": mine_s, "seconds_incl_gather": total_s,
                          "config": {"workload": f"{args.config}: n={cfg.n} d={cfg.d} c={cfg.c} {cfg.cov_form} "
                                                 f"{cfg.noise_corr}; {args.rounds} GPS rounds x optim({args.method}, maxit 10) "
                                                 f"+ test/train predictions; {args.batch} fits in lock-step per GPU; "
                                                 f"driver: {'gps_fit_batch (C ABI)' if args.driver == 'c' else 'batchfit (Python)'}",
                                     "partition": f"{args.fits} independent fits over {world} GPUs, no collective"},
                          "mean_objective": float(np.mean([v["objective"] for v in full.values()]))}), flush=True)
    if fit is not None:
        fit.close()
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    raise SystemExit(main())
