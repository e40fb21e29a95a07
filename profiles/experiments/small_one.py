import sys, torch
sys.path.insert(0, ".")
from islamcts_b200.engine import Engine
nt = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
g = torch.Generator().manual_seed(7)
roots = (torch.rand((nt, 4), generator=g, dtype=torch.float64) - 0.5) * 0.1
eng = Engine(0, n_trees=nt, sim_capacity=1000, max_depth=500, gamma=0.99, kernel=1, seed=1)
eng.set_roots(roots).run(1000); torch.cuda.synchronize()
print(eng.counters())
