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
    n_roots: int               # distinct roots of the WHOLE job
    tree_id_base: int          # global id of this rank's tree 0 (also its Philox stream offset)
    group: np.ndarray          # int32 [trees_per_rank]: LOCAL root index of every local tree (ascending)
    local_roots: int = 0       # roots this rank holds trees of
    root_offset: int = 0       # global id of local root 0
    scaling: str = "weak"


def plan(rank: int, world: int, trees_per_rank: int, n_roots: int) -> Plan:
    """Weak scaling: every rank searches trees_per_rank trees of its own, trees_per_rank / n_roots for each of the
    job's n_roots roots (every rank holds every root)."""
    if trees_per_rank % n_roots:
        raise ValueError("trees_per_rank must be a multiple of n_roots")
    per = trees_per_rank // n_roots
    group = (np.arange(trees_per_rank) // per).astype(np.int32)
    return Plan(rank, world, trees_per_rank, n_roots, rank * trees_per_rank, group, n_roots, 0, "weak")


def plan_total(rank: int, world: int, total_trees: int, n_roots: int) -> Plan:
    """Strong scaling (BASELINE config 5: "65 536 trees sharded across 1/2/4/8"): the job has total_trees trees, tree i
    belongs to root i // (total_trees / n_roots) and to rank i // (total_trees / world).  A rank therefore holds either
    n_roots / world whole roots or a 1 / (world / n_roots) share of one root."""
    if total_trees % world or total_trees % n_roots:
        raise ValueError("total_trees must be a multiple of the world size and of n_roots")
    tpr, per = total_trees // world, total_trees // n_roots
    if tpr % per and per % tpr:
        raise ValueError("a rank must hold whole roots or a whole fraction of one root")
    first = rank * tpr
    gl = (np.arange(first, first + tpr) // per)
    root_offset = int(gl[0])
    group = (gl - root_offset).astype(np.int32)
    return Plan(rank, world, tpr, n_roots, first, group, int(group[-1]) + 1, root_offset, "strong")


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


def all_reduce_packed(packed: torch.Tensor, group=None) -> None:
    """The single exchange step of the packed form: float64 [n_roots, A, 2] = (na, total) summed over ranks, in place
    (ranks contribute zeros for roots they hold no tree of, so the same call is merge and gather)."""
    import torch.distributed as dist
    dist.all_reduce(packed, op=dist.ReduceOp.SUM, group=group)


def pick(mna: torch.Tensor, mtot: torch.Tensor, seed: int = 0) -> torch.Tensor:
    """q = total/na per root and a UNIFORMLY RANDOM arg-max among exact ties (end of fit(), mcts_hash.py:42-50:
    `np.random.choice(max_children)`) on the merged statistics.  The draw comes from a generator seeded with `seed` on
    the tensors' device, so every rank of a job (same seed, same merged statistics) picks the same action."""
    q = torch.where(mna > 0, mtot / mna.clamp(min=1).to(torch.float64), torch.full_like(mtot, -float("inf")))
    ties = q == q.max(dim=1, keepdim=True).values
    count = ties.sum(dim=1)
    g = torch.Generator(device=q.device).manual_seed(int(seed))
    u = torch.rand(q.shape[0], generator=g, device=q.device, dtype=torch.float64)
    k = torch.minimum((u * count).floor().to(torch.int64), count - 1)          # position among the tied maxima
    hit = ties & (ties.cumsum(dim=1) == (k + 1).unsqueeze(1))
    return hit.to(torch.int8).argmax(dim=1).to(torch.int32)


# ---------------------------------------------------------------- continuous-action agents (apw / dpw)
# Every tree samples its own root actions, so the root children of the trees of one root are LISTS that differ per
# tree: the merge is a union, not a dense sum -- locally over a rank's trees, then one all-gather.

def merge_root_children(actions: torch.Tensor, na: torch.Tensor, total: torch.Tensor):
    """Union of root-children lists: rows of `actions` [n, action_dim] that are bit-equal (the centre / bounds genetic
    expansions add to every tree, duplicates of a sample) are summed; order = first appearance.  Returns the merged
    (actions, na:int64, total:float64)."""
    actions = actions.reshape(len(na), -1).to(torch.float64)
    if len(na) == 0:
        return actions, na.to(torch.int64), total.to(torch.float64)
    uniq, inverse = torch.unique(actions, dim=0, return_inverse=True)
    first = torch.full((len(uniq),), len(na), dtype=torch.int64, device=actions.device)
    first = first.scatter_reduce(0, inverse, torch.arange(len(na), device=actions.device), reduce="amin")
    order = torch.argsort(first)
    mna = torch.zeros(len(uniq), dtype=torch.int64, device=actions.device).index_add_(0, inverse, na.to(torch.int64))
    mtot = torch.zeros(len(uniq), dtype=torch.float64, device=actions.device).index_add_(0, inverse, total.to(torch.float64))
    return uniq[order], mna[order], mtot[order]


def all_gather_root_children(actions: torch.Tensor, na: torch.Tensor, total: torch.Tensor, group=None):
    """The exchange step for continuous-action agents: every rank contributes its (already locally merged) root children
    (actions [n_r, action_dim], na [n_r], total [n_r]); all ranks receive the union in rank order, merged again."""
    import torch.distributed as dist
    world = dist.get_world_size(group)
    actions = actions.reshape(len(na), -1).to(torch.float64)
    ad = actions.shape[1]
    cnt = torch.tensor([len(na), ad], dtype=torch.int64, device=actions.device)
    cnts = [torch.zeros_like(cnt) for _ in range(world)]
    dist.all_gather(cnts, cnt, group=group)
    ad = int(max(int(c[1]) for c in cnts))              # a rank without children still needs the row width
    m = int(max(int(c[0]) for c in cnts))
    packed = torch.zeros((max(m, 1), ad + 2), dtype=torch.float64, device=actions.device)
    if len(na):
        packed[: len(na), :ad] = actions
        packed[: len(na), ad] = na.to(torch.float64)    # exact: visit counts are far below 2**53
        packed[: len(na), ad + 1] = total.to(torch.float64)
    parts = [torch.zeros_like(packed) for _ in range(world)]
    dist.all_gather(parts, packed, group=group)
    rows = torch.cat([parts[r][: int(cnts[r][0])] for r in range(world)], dim=0)
    return merge_root_children(rows[:, :ad], rows[:, ad].to(torch.int64), rows[:, ad + 1])


def pick_continuous(actions: torch.Tensor, na: torch.Tensor, total: torch.Tensor, seed: int = 0):
    """arg-max q over the merged children, exact ties broken uniformly (mcts_action_...:38-44); returns (index, action row)."""
    q = total / na.clamp(min=1).to(torch.float64)
    tied = torch.nonzero(q == q.max()).reshape(-1)
    g = torch.Generator(device=q.device).manual_seed(int(seed))
    i = int(tied[int(torch.randint(len(tied), (1,), generator=g, device=q.device).item())].item()) if len(tied) > 1 else int(tied[0].item())
    return i, actions[i]
