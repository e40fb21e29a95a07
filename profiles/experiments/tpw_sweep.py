import sys, torch
sys.path.insert(0, ".")
from islamcts_b200.engine import Engine
g = torch.Generator().manual_seed(7)
for nt in (4096, 16384, 32768, 65536, 131072):
    roots = (torch.rand((nt, 4), generator=g, dtype=torch.float64) - 0.5) * 0.1
    out = []
    for tpw in (2, 4, 8, 16, 32):
        eng = Engine(0, n_trees=nt, sim_capacity=1000, max_depth=500, gamma=0.99, kernel=1, seed=1, tuning=tpw << 8)
        eng.set_roots(roots).run(1000); torch.cuda.synchronize()
        best = 1e9
        for _ in range(2):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); eng.set_roots(roots).run(1000); e1.record(); torch.cuda.synchronize(); best = min(best, e0.elapsed_time(e1))
        out.append("tpw%d %.0fM" % (tpw, nt * 1000 / best / 1e3))
        del eng; torch.cuda.empty_cache()
    print("trees", nt, " ".join(out), flush=True)
