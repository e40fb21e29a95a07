import sys, torch
sys.path.insert(0, ".")
from islamcts_b200.engine import Engine
g = torch.Generator().manual_seed(7)
for env_id, S, md, gam, nsim in ((0, 4, 500, 0.99, 1000), (3, 1, 1000, 1.0, 100)):
    for nt in (1, 8, 32):
        roots = (torch.rand((nt, S), generator=g, dtype=torch.float64) - 0.5) * 0.1 if env_id == 0 else torch.zeros((nt, 1), dtype=torch.float64)
        res = []
        for kern in (1, 2):
            eng = Engine(env_id, n_trees=nt, sim_capacity=nsim, max_depth=md, gamma=gam, C_=1.0 if env_id == 0 else 0.5, kernel=kern, seed=1)
            eng.set_roots(roots).run(nsim); torch.cuda.synchronize()
            best = 1e9
            for _ in range(5):
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record(); eng.set_roots(roots).run(nsim); e1.record(); torch.cuda.synchronize(); best = min(best, e0.elapsed_time(e1))
            res.append(best)
        print("env", env_id, "trees", nt, "nsim", nsim, "TPT %.3f ms  CTA %.3f ms" % tuple(res), flush=True)
