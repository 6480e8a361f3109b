"""CPU tests of the multi-GPU host logic: the partitioner, and the N > 1 path with two `gloo`
processes on CPU (each rank evaluates its shard with the CPU oracle standing in for its GPU, the
results are gathered on the host — the only cross-rank step of the design)."""
import os
import socket
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

from gpsurvival_b200 import partition as part


def test_partition_is_exact_balanced_and_deterministic():
    rng = np.random.default_rng(0)
    costs = [part.estimated_cost(int(n), int(d)) for n, d in zip(rng.integers(50, 500, 97), rng.integers(1, 3000, 97))]
    for world in (1, 2, 4, 8):
        shards = part.partition(costs, world)
        assert sorted(i for s in shards for i in s) == list(range(97))
        assert shards == part.partition(costs, world)
        loads = [sum(costs[i] for i in s) for s in shards]
        assert max(loads) - min(loads) <= max(costs)            # LPT bound
    assert part.partition([], 3) == [[], [], []]
    assert part.partition([1.0], 4) == [[0], [], [], []]
    eq = part.partition([1.0] * 512, 8)                          # C4: 512 equal fits -> 64 per GPU
    assert [len(s) for s in eq] == [64] * 8
    with pytest.raises(ValueError):
        part.partition([1.0], 0)


def test_group_by_shape():
    g = part.group_by_shape([(400, 2000, 200), (400, 2000, 200), (285, 10000, 142), (400, 2000, 200)])
    assert g == {(400, 2000, 200): [0, 1, 3], (285, 10000, 142): [2]}


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    import torch.distributed as dist
    from gpsurvival_b200 import partition as part, synth
    from oracle import gp_oracle as orc
    dist.init_process_group("gloo", rank=rank, world_size=world)
    shapes = [(30 + 7 * i, 2 + i % 3) for i in range(9)]
    costs = [part.estimated_cost(n, d) for n, d in shapes]

    def evaluate(idx):
        out = {}
        for i in idx:
            n, d = shapes[i]
            pb = synth.make_problem(n, 4, d, n // 3, 500 + i)
            par = np.log([0.2, 0.8, 1.7 * np.sqrt(d)])
            out[i] = (rank, orc.nlml(par, pb["trainingData"], np.log(pb["trainingTargets"]), pb["events"]))
        return out

    full = part.run_sharded(costs, evaluate, rank, world)
    dist.barrier()
    dist.destroy_process_group()
    q.put((rank, {k: v for k, v in full.items()}))


def test_two_rank_gloo_sharded_sweep():
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    got = dict(q.get(timeout=240) for _ in procs)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert sorted(got) == [0, 1]
    assert got[0].keys() == got[1].keys() and sorted(got[0]) == list(range(9))
    for k in got[0]:
        assert got[0][k] == got[1][k]                              # every rank holds the same gathered table
    owners = {k: v[0] for k, v in got[0].items()}
    assert set(owners.values()) == {0, 1}                          # both ranks did work
    # the sharded sweep equals the serial one
    from gpsurvival_b200 import synth
    from oracle import gp_oracle as orc
    for i in (0, 8):
        n, d = 30 + 7 * i, 2 + i % 3
        pb = synth.make_problem(n, 4, d, n // 3, 500 + i)
        ref = orc.nlml(np.log([0.2, 0.8, 1.7 * np.sqrt(d)]), pb["trainingData"], np.log(pb["trainingTargets"]), pb["events"])
        assert got[0][i][1] == ref


def test_batch_bfgs_follows_sequential_vmmin():
    """The lock-step batched BFGS reproduces, fit by fit, the sequential restatement of R's vmmin
    (itself pinned on R's documented example in test_oracle.py)."""
    from scipy.optimize import rosen, rosen_der
    from gpsurvival_b200 import optim
    from gpsurvival_b200.batch_bfgs import batch_bfgs
    rng = np.random.default_rng(3)
    for n in (2, 6):
        x0 = np.vstack([np.r_[-1.2, 1.0, np.zeros(n - 2)]] + [rng.uniform(-1.5, 1.5, n) for _ in range(5)])
        calls = {"n": 0}

        def evaluate(P):
            calls["n"] += 1
            return np.array([rosen(p) for p in P]), np.array([rosen_der(p) for p in P])

        for maxit in (7, 200):
            res, ticks = batch_bfgs(evaluate, x0, maxit=maxit, device="cpu")
            for i in range(x0.shape[0]):
                ref = optim.bfgs(rosen, rosen_der, x0[i], maxit=maxit)
                assert tuple(res[i]["counts"]) == tuple(ref["counts"]), (n, maxit, i)
                assert res[i]["convergence"] == ref["convergence"]
                assert abs(res[i]["value"] - ref["value"]) <= 1e-9 * max(1.0, abs(ref["value"]))
                assert np.allclose(res[i]["par"], ref["par"], rtol=1e-8, atol=1e-10)
    r, _ = batch_bfgs(evaluate, np.array([[-1.2, 1.0]]), maxit=100, device="cpu")
    assert tuple(r[0]["counts"]) == (110, 43) and abs(r[0]["value"] - 9.594956e-18) < 5e-24    # R's example(optim)
    # inactive fits are left alone; non-finite values are rejected points, not errors
    def ev2(P):
        f = np.array([rosen(p) if p[0] < 1.05 else np.inf for p in P])
        return f, np.array([rosen_der(p) for p in P])
    r2, _ = batch_bfgs(ev2, x0[:3, :2] if x0.shape[1] >= 2 else x0, maxit=30, active=[True, False, True], device="cpu")
    assert r2[1] is None and r2[0] is not None and np.isfinite(r2[0]["value"])


def test_c_abi_lpt_partitioner_balances_and_is_deterministic(lib):
    """gps_partition_lpt (the host partitioner behind gps_fit_batch_multi): every job assigned, loads balanced to within the
    largest job (the LPT guarantee), identical jobs split evenly, same answer on every call."""
    import numpy as np
    from gpsurvival_b200 import gp
    rng = np.random.default_rng(0)
    cost = rng.uniform(1.0, 10.0, 200)
    for parts in (1, 2, 4, 8):
        a = gp.partition_lpt(cost, parts)
        assert a.min() >= 0 and a.max() < parts and np.array_equal(a, gp.partition_lpt(cost, parts))
        loads = np.array([cost[a == g].sum() for g in range(parts)])
        assert loads.max() - loads.min() <= cost.max() + 1e-9
    even = gp.partition_lpt(np.full(512, 3.0), 8)
    assert np.all(np.bincount(even, minlength=8) == 64)
