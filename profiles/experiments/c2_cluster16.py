import sys, torch
sys.path.insert(0, ".")
from islamcts_b200.engine import Engine
g = torch.Generator().manual_seed(7)
roots = (torch.rand((1, 4), generator=g, dtype=torch.float64) - 0.5) * 0.1
ref = None
for cs, bt in ((0, 0), (8, 512), (8, 256), (8, 128), (16, 256), (16, 128), (16, 64), (4, 512), (4, 256)):
    try:
        eng = Engine(0, n_trees=1, sim_capacity=1000, max_depth=500, C_=1.0, gamma=0.99, kernel=2, rollouts_per_leaf=4096, seed=1, cluster_ctas=cs, block_threads=bt)
        eng.set_roots(roots).run(1000); torch.cuda.synchronize()
        best = 1e9
        for _ in range(3):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); eng.set_roots(roots).run(1000); e1.record(); torch.cuda.synchronize(); best = min(best, e0.elapsed_time(e1))
        na, tot = eng.root_stats()
        key = (na.cpu().numpy().tolist(), tot.cpu().numpy().tolist())
        if ref is None: ref = key
        print("cluster", cs, "threads", bt, "us/sim %.2f" % best, "sims/s %.0f" % (1e6 / best), "same" if key == ref else "DIFFERENT", eng.counters()["faults"], flush=True)
    except Exception as e:
        print("cluster", cs, "threads", bt, "failed:", str(e)[:200], flush=True)
