"""Host-side logic of the root-parallel sharding (north_star subsystem 4).  Pure torch/numpy so that
it runs under gloo on CPU as well as under NCCL: which trees a rank owns, which root each tree belongs to,
and the one exchange step (all-reduce of the per-root statistics)."""
from __future__ import annotations

from dataclasses import dataclass

import numpy as np
import torch


@dataclass
class Plan:
    rank: int
    world: int
    trees_per_rank: int
    n_roots: int
    tree_id_base: int          # global id of this rank's tree 0 (also its Philox stream offset)
    group: np.ndarray          # int32 [trees_per_rank]: root id of every local tree


def plan(rank: int, world: int, trees_per_rank: int, n_roots: int) -> Plan:
    if trees_per_rank % n_roots:
        raise ValueError("trees_per_rank must be a multiple of n_roots")
    per = trees_per_rank // n_roots
    group = (np.arange(trees_per_rank) // per).astype(np.int32)
    return Plan(rank, world, trees_per_rank, n_roots, rank * trees_per_rank, group)


def segmented_sum(na: torch.Tensor, total: torch.Tensor, group: torch.Tensor, n_roots: int):
    """Reference implementation of isla_merge_root_stats for host tensors (CPU tests)."""
    A = na.shape[1]
    mna = torch.zeros((n_roots, A), dtype=torch.int64).index_add_(0, group.long(), na.to(torch.int64))
    mtot = torch.zeros((n_roots, A), dtype=torch.float64).index_add_(0, group.long(), total.to(torch.float64))
    return mna, mtot


def all_reduce_root_stats(mna: torch.Tensor, mtot: torch.Tensor, group=None) -> None:
    """The single exchange step: dense [n_roots, A] (na:int64, total:float64) summed over ranks, in place."""
    import torch.distributed as dist
    dist.all_reduce(mna, op=dist.ReduceOp.SUM, group=group)
    dist.all_reduce(mtot, op=dist.ReduceOp.SUM, group=group)


def pick(mna: torch.Tensor, mtot: torch.Tensor) -> torch.Tensor:
    """q = total/na per root and its arg-max (end of fit(), mcts_hash.py:42-50) on the merged statistics."""
    q = torch.where(mna > 0, mtot / mna.clamp(min=1).to(torch.float64), torch.full_like(mtot, -float("inf")))
    return q.argmax(dim=1).to(torch.int32)
