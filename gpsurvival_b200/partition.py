"""Partitioning of independent fits over GPUs.

The hot path shards ACROSS fits only — optimiser restarts, CV folds, repetitions, model variants
are fully independent (e.g. the nReps loops of runSyntheticExp1.R:208-213) — so the data path has
no collective: each rank evaluates its own shard on its own GPU, and the only exchange is the
host-side gathering of per-fit results (a few KB per fit) at the end.  A single large fit stays
on one GPU.

Pure host logic; torch.distributed is used only for the final gather when it is initialised.
"""
from __future__ import annotations

from typing import Callable, Dict, List, Sequence


def estimated_cost(n: int, d: int, ard: bool = True) -> float:
    """Algorithmic flops of one NLML + gradient evaluation."""
    return (3.0 if ard else 1.0) * n * n * d + float(n) ** 3


def partition(costs: Sequence[float], world: int) -> List[List[int]]:
    """Longest-processing-time-first assignment of problem indices to `world` ranks: problems
    sorted by decreasing cost (ties by index), each given to the currently least-loaded rank
    (ties to the lowest rank).  Deterministic; every index appears exactly once; each shard is
    returned in increasing index order."""
    if world < 1:
        raise ValueError("world must be >= 1")
    order = sorted(range(len(costs)), key=lambda i: (-float(costs[i]), i))
    load = [0.0] * world
    shards: List[List[int]] = [[] for _ in range(world)]
    for i in order:
        r = min(range(world), key=lambda k: (load[k], k))
        shards[r].append(i)
        load[r] += float(costs[i])
    return [sorted(s) for s in shards]


def shard_indices(costs: Sequence[float], rank: int, world: int) -> List[int]:
    return partition(costs, world)[rank]


def group_by_shape(shapes: Sequence[tuple]) -> Dict[tuple, List[int]]:
    """Indices of problems that can share one batched handle (identical (n, d, n_censored))."""
    groups: Dict[tuple, List[int]] = {}
    for i, s in enumerate(shapes):
        groups.setdefault(tuple(s), []).append(i)
    return groups


def gather_results(local: Dict[int, object], world: int) -> Dict[int, object]:
    """Host-side concatenation of per-fit results from all ranks (identity when world == 1)."""
    if world == 1:
        return dict(local)
    import torch.distributed as dist
    if not dist.is_initialized():
        raise RuntimeError("torch.distributed is not initialised")
    parts = [None] * world
    dist.all_gather_object(parts, local)
    out: Dict[int, object] = {}
    for p in parts:
        for k, v in p.items():
            if k in out:
                raise RuntimeError(f"problem {k} was evaluated by two ranks")
            out[k] = v
    return out


def run_sharded(costs: Sequence[float], evaluate: Callable[[List[int]], Dict[int, object]], rank: int,
                world: int) -> Dict[int, object]:
    """Evaluate the shard of this rank with `evaluate(indices) -> {index: result}` and return the
    results of ALL problems on every rank."""
    mine = shard_indices(costs, rank, world)
    local = evaluate(mine)
    if sorted(local) != mine:
        raise RuntimeError("evaluate() must return exactly the indices it was given")
    full = gather_results(local, world)
    if sorted(full) != list(range(len(costs))):
        raise RuntimeError("gathered results do not cover every problem exactly once")
    return full
