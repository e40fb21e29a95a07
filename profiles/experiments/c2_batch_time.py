import sys, torch
sys.path.insert(0, ".")
from islamcts_b200.engine import Engine
g = torch.Generator().manual_seed(7)
for nt, R in ((296, 4096), (296, 512), (1184, 64), (148, 4096)):
    roots = (torch.rand((nt, 4), generator=g, dtype=torch.float64) - 0.5) * 0.1
    eng = Engine(0, n_trees=nt, sim_capacity=1000, max_depth=500, C_=1.0, gamma=0.99, kernel=2, rollouts_per_leaf=R, seed=1)
    eng.set_roots(roots).run(200); torch.cuda.synchronize()
    best = 1e9
    for _ in range(3):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        eng.set_roots(roots); c0 = eng.counters()["env_steps"]
        e0.record(); eng.run(200); e1.record(); torch.cuda.synchronize(); best = min(best, e0.elapsed_time(e1))
        steps = eng.counters()["env_steps"] - c0
    print("trees", nt, "R", R, "ms %.2f" % best, "sims/s %.3fM" % (nt * 200 / best / 1e3), "env-steps/s %.1fG" % (steps / best / 1e6))
