"""Round-2 A/B driver: C5-shaped launches (CartPole, nsim 1000) for several batch sizes and trees-per-warp settings.
usage: python profiles/experiments/r2_sweep.py [trees:tpw,tpw,... ...]   (tpw 0 = the launch's own choice)"""
import sys, torch
sys.path.insert(0, ".")
from islamcts_b200.engine import Engine
spec = sys.argv[1:] or ["8192:0,2,4,8", "16384:0,4,8", "65536:0,16"]
g = torch.Generator().manual_seed(7)
for item in spec:
    nt, tpws = item.split(":")
    nt = int(nt)
    roots = (torch.rand((nt, 4), generator=g, dtype=torch.float64) - 0.5) * 0.1
    out = []
    for tpw in [int(x) for x in tpws.split(",")]:
        eng = Engine(0, n_trees=nt, sim_capacity=1000, max_depth=500, gamma=0.99, kernel=1, seed=1, tuning=tpw << 8)
        eng.set_roots(roots).run(1000); torch.cuda.synchronize()
        best = 1e9
        for _ in range(3):
            eng.set_roots(roots); torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); eng.run(1000); e1.record(); torch.cuda.synchronize(); best = min(best, e0.elapsed_time(e1))
        c = eng.counters()
        out.append("tpw%d %.1fM (%.2f ms)" % (tpw, nt * 1000 / best / 1e3, best))
        assert c["faults"] == 0, c
        del eng; torch.cuda.empty_cache()
    print("trees", nt, " | ".join(out), flush=True)
