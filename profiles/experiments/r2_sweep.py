"""Round-2 A/B driver: C5-shaped launches (CartPole, nsim 1000) for several batch sizes and launch shapes.
usage: python profiles/experiments/r2_sweep.py [trees:shape,shape,... ...]
shape = <trees per warp> (lock-step kernel; 0 = the launch's own choice) | f<trees per CTA>x<worker warps> (worker / owner kernel)
        | auto (whatever the library picks)"""
import sys, torch
sys.path.insert(0, ".")
from islamcts_b200.engine import Engine
spec = sys.argv[1:] or ["8192:auto", "16384:auto", "65536:auto"]
g = torch.Generator().manual_seed(7)
for item in spec:
    nt, shapes = item.split(":")
    nt = int(nt)
    roots = (torch.rand((nt, 4), generator=g, dtype=torch.float64) - 0.5) * 0.1
    out = []
    for sh in shapes.split(","):
        if sh == "auto":
            tuning = 0
        else:
            flow = False
            tpc, _, w = sh.lstrip("f").partition("x")
            tuning = (int(tpc or 0) << 8) | (int(w or 0) << 16)
        eng = Engine(0, n_trees=nt, sim_capacity=1000, max_depth=500, gamma=0.99, kernel=1, seed=1, tuning=tuning)
        eng.set_roots(roots).run(1000); torch.cuda.synchronize()
        best = 1e9
        for _ in range(3):
            eng.set_roots(roots); torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); eng.run(1000); e1.record(); torch.cuda.synchronize(); best = min(best, e0.elapsed_time(e1))
        c = eng.counters()
        out.append("%s %.1fM (%.2f ms)" % (sh, nt * 1000 / best / 1e3, best))
        assert c["faults"] == 0, c
        del eng; torch.cuda.empty_cache()
    print("trees", nt, " | ".join(out), flush=True)
