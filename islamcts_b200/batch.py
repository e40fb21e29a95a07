"""Root-parallel batched search: many independent vanilla trees per GPU, root statistics merged per root.

This is the public call for BASELINE config C5 ("65 536 independent CartPole trees x nsim=1000 sharded
across 1/2/4/8 B200 with NCCL root-statistics merge") and north_star subsystem (4).  The reference's only
parallel construct is joblib over whole experiments (islaMcts/experiment_runner.py:305-306); here the
unit of sharding is the tree:

    strong scaling (total_trees=...):  tree i -> rank i // (total / world), root i // (total / n_roots);
    weak scaling (trees_per_rank=...): every rank searches trees_per_rank trees, trees_per_rank / n_roots per root;
    per step: ONE kernel sums the trees of each root on the rank in a fixed order, straight from the root blocks
    (isla_engine_root_merge), ONE all-reduce(sum) of the packed [n_roots, A, 2] (na, total) float64 array crosses
    NVLink (ranks contribute zeros for roots they hold no tree of: merge and gather in one call), and every rank takes
    q = total/na and the uniformly random arg-max among exact ties locally (mcts_hash.py:42-50).

No communication happens during the search itself.
"""
from __future__ import annotations

import numpy as np
import torch

from . import _lib, sharding
from . import engine as E


class RootParallelSearch:
    def __init__(self, env: str | int = "CartPole-v1", trees_per_rank: int | None = None, n_sim: int = 1000, C: float = 1.0,
                 gamma: float = 0.99, max_depth: int = 500, n_roots: int | None = None, seed: int = 0,
                 precision: int = _lib.PRECISION_F32, budget_mode: int = _lib.BUDGET_TO_MAX_DEPTH,
                 device="cuda:0", process_group=None, rank: int = 0, world_size: int = 1, total_trees: int | None = None):
        self.env_id = E.ENV_NAMES[env] if isinstance(env, str) else int(env)
        self.device = torch.device(device)
        self.rank, self.world = int(rank), int(world_size)
        self.pg = process_group
        self.n_sim = int(n_sim)
        self.seed = int(seed)
        self.A = E.N_ACTIONS[self.env_id]
        self.S = E.STATE_DIM[self.env_id]
        if total_trees is not None:
            # strong scaling: the job's trees are split over the ranks
            self.n_roots = int(n_roots) if n_roots else int(total_trees)
            self.plan = sharding.plan_total(self.rank, self.world, int(total_trees), self.n_roots)
        else:
            # weak scaling: every rank holds trees_per_rank / n_roots trees of each of the n_roots roots
            trees_per_rank = 65536 if trees_per_rank is None else int(trees_per_rank)
            self.n_roots = int(n_roots) if n_roots else trees_per_rank
            self.plan = sharding.plan(self.rank, self.world, trees_per_rank, self.n_roots)
        self.trees = self.plan.trees_per_rank
        self.engine = E.Engine(self.env_id, agent=_lib.AGENT_VANILLA, n_trees=self.trees, sim_capacity=self.n_sim,
                               max_depth=max_depth, C_=C, gamma=gamma, budget_mode=budget_mode, precision=precision,
                               kernel=_lib.KERNEL_THREAD_PER_TREE, seed=seed, tree_id_base=self.plan.tree_id_base,
                               device=self.device)
        # local root index of every local tree, and the job-wide packed statistics [n_roots, A, 2]
        self.group = torch.as_tensor(self.plan.group, device=self.device).contiguous()
        self._packed = torch.zeros((self.n_roots, self.A, 2), dtype=torch.float64, device=self.device)
        lo = self.plan.root_offset
        self._local = self._packed[lo: lo + self.plan.local_roots]
        self._out_host = None
        self._fits = 0

    # roots_of_groups: [n_roots, S]; each group's trees all start from that root
    def expand_roots(self, roots_of_groups: torch.Tensor) -> torch.Tensor:
        r = roots_of_groups.to(self.device, torch.float64)
        return r[self.plan.root_offset + self.group.long()].contiguous()

    def merge_message_bytes(self) -> int:
        """Bytes one rank contributes to the all-reduce of a step."""
        return self.n_roots * self.A * 16

    def search_device(self, tree_roots_dev: torch.Tensor):
        """Device-resident pass: roots [trees, S] float64 on the GPU -> merged (na int64, total float64) [n_roots, A]."""
        eng = self.engine
        eng.set_roots(tree_roots_dev)
        eng.run(self.n_sim)
        if self.world > 1 and self.plan.local_roots != self.n_roots:
            self._packed.zero_()                      # roots this rank holds no tree of contribute zeros
        eng.root_merge(self.plan.local_roots, self._local)
        if self.world > 1:
            sharding.all_reduce_packed(self._packed, self.pg)
        return self._packed[..., 0].to(torch.int64), self._packed[..., 1]

    def pick(self, mna, mtot):
        self._fits += 1
        return sharding.pick(mna, mtot, seed=self.seed * 1000003 + self._fits)

    def fit(self, roots_host: torch.Tensor):
        """End-to-end call with HOST buffers: roots [n_roots, S] float64 (pinned) -> dict of host tensors."""
        if self._out_host is None:
            self._out_host = dict(
                action=torch.empty(self.n_roots, dtype=torch.int32).pin_memory(),
                na=torch.empty((self.n_roots, self.A), dtype=torch.int64).pin_memory(),
                total=torch.empty((self.n_roots, self.A), dtype=torch.float64).pin_memory())
        g = roots_host.to(self.device, non_blocking=True)
        mna, mtot = self.search_device(self.expand_roots(g))
        act = self.pick(mna, mtot)
        self._out_host["action"].copy_(act, non_blocking=True)
        self._out_host["na"].copy_(mna, non_blocking=True)
        self._out_host["total"].copy_(mtot, non_blocking=True)
        torch.cuda.current_stream(self.device).synchronize()
        return self._out_host

    def h2d_bytes(self) -> int:
        return self.n_roots * self.S * 8

    def d2h_bytes(self) -> int:
        return self.n_roots * (4 + self.A * 16)
