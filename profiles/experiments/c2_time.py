import sys, torch
sys.path.insert(0, ".")
from islamcts_b200.engine import Engine
g = torch.Generator().manual_seed(7)
roots = (torch.rand((1, 4), generator=g, dtype=torch.float64) - 0.5) * 0.1
for R, cs in ((4096, 0), (4096, 1), (512, 0), (32, 0), (1, 0)):
    eng = Engine(0, n_trees=1, sim_capacity=1000, max_depth=500, C_=1.0, gamma=0.99, kernel=2, rollouts_per_leaf=R, seed=1, cluster_ctas=cs)
    eng.set_roots(roots).run(1000); torch.cuda.synchronize()
    best = 1e9
    for _ in range(3):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); eng.set_roots(roots).run(1000); e1.record(); torch.cuda.synchronize(); best = min(best, e0.elapsed_time(e1))
    c = eng.counters()
    print("R", R, "cluster", cs, "ms %.2f" % best, "us/sim %.1f" % best, "sims/s %.0f" % (1000 / best * 1e3), "env-steps/s %.2fG" % (c["env_steps"] / best / 1e6), "levels/sim %.0f" % (c["levels"] / 1000))
