import sys, torch
sys.path.insert(0, ".")
from islamcts_b200.engine import Engine
g = torch.Generator().manual_seed(7)
for nt in (8192, 65536):
    roots = (torch.rand((nt, 4), generator=g, dtype=torch.float64) - 0.5) * 0.1
    for bm in (0, 1):
        eng = Engine(0, n_trees=nt, sim_capacity=1000, max_depth=500, gamma=0.99, kernel=1, seed=1, budget_mode=bm)
        eng.set_roots(roots).run(1000); torch.cuda.synchronize()
        best = 1e9
        for _ in range(3):
            eng.set_roots(roots); torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); eng.run(1000); e1.record(); torch.cuda.synchronize(); best = min(best, e0.elapsed_time(e1))
        c = eng.counters()
        print("trees", nt, "budget_mode", bm, "%.2f ms  %.1fM sims/s  levels/sim %.1f  rollout steps/sim %.1f" % (best, nt * 1000 / best / 1e3, c["levels"] / c["sims"], c["rollout_steps"] / c["sims"]), flush=True)
        del eng; torch.cuda.empty_cache()
