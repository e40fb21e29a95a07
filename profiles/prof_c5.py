"""Profiling driver: a reduced C5 pass (same kernel, same 65 536 trees, fewer simulations) for ncu."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from islamcts_b200.batch import RootParallelSearch

nsim = int(sys.argv[1]) if len(sys.argv) > 1 else 200
trees = int(sys.argv[2]) if len(sys.argv) > 2 else 65536
s = RootParallelSearch("CartPole-v1", trees_per_rank=trees, n_sim=nsim, n_roots=256, seed=0)
g = torch.Generator().manual_seed(1234)
roots = ((torch.rand((256, 4), generator=g, dtype=torch.float64) - 0.5) * 0.1)
tr = s.expand_roots(roots)
for _ in range(2):
    s.search_device(tr)
torch.cuda.synchronize()
print(s.engine.counters())
