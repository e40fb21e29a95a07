"""Profiling driver for the secondary kernels: C2 as a batch of 296 trees (CTA engine, rollout-dominated) and the
standalone rollout kernels."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from islamcts_b200 import _lib
from islamcts_b200 import engine as E
from islamcts_b200.engine import Engine

g = torch.Generator().manual_seed(7)
eng = Engine(_lib.ENV_CARTPOLE, n_trees=296, sim_capacity=1000, max_depth=500, C_=1.0, gamma=0.99, kernel=_lib.KERNEL_CTA_PER_TREE,
             rollouts_per_leaf=4096, seed=1)
roots = (torch.rand((296, 4), generator=g, dtype=torch.float64) - 0.5) * 0.1
eng.set_roots(roots).run(200)
print(eng.counters())
one = Engine(_lib.ENV_CARTPOLE, n_trees=1, sim_capacity=1000, max_depth=500, C_=1.0, gamma=0.99, kernel=_lib.KERNEL_CTA_PER_TREE,
             rollouts_per_leaf=4096, seed=1)
one.set_roots(roots[:1]).run(300)
for env_id, budget in ((0, 500), (1, 200), (2, 500), (3, 1000)):
    n = 1 << 24 if env_id in (0, 3) else 1 << 21
    E.rollout(env_id, None, n=n, budget=budget, gamma=0.99, seed=1234, sim=0)
torch.cuda.synchronize()
# round 2: the APW widening wave (C3), the tree-parallel DPW search (C4, 64 simulations in flight) and the small-batch launch
# of the thread-per-tree kernel (8192 trees = the per-GPU shard of the 8-GPU strong-scaling run)
c3 = Engine(_lib.ENV_PENDULUM, agent=_lib.AGENT_APW, n_trees=1, sim_capacity=10000, max_depth=200, C_=1.0, gamma=0.99, alpha1=0.8,
            k1=60.0, kernel=_lib.KERNEL_CTA_PER_TREE, seed=1)
c3.set_roots(torch.tensor([[1.0, 0.3]], dtype=torch.float64)).run(10000)
c4 = Engine(_lib.ENV_MOUNTAINCAR, agent=_lib.AGENT_DPW, n_trees=1, sim_capacity=10000, max_depth=500, C_=1.0, gamma=0.99, alpha1=0.5,
            k1=1.0, alpha2=0.5, k2=1.0, kernel=_lib.KERNEL_CTA_PER_TREE, seed=1, tree_parallel=64, virtual_value=-5.0)
c4.set_roots(torch.tensor([[-0.5, 0.0]], dtype=torch.float64)).run(10000)
sm = Engine(_lib.ENV_CARTPOLE, n_trees=8192, sim_capacity=1000, max_depth=500, C_=1.0, gamma=0.99, kernel=_lib.KERNEL_THREAD_PER_TREE, seed=1)
sm.set_roots((torch.rand((8192, 4), generator=g, dtype=torch.float64) - 0.5) * 0.1).run(1000)
torch.cuda.synchronize()
print(c3.counters(), c4.counters(), sm.counters())
