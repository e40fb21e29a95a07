import os, sys, time, torch
sys.path.insert(0, ".")
from islamcts_b200.batch import RootParallelSearch
for trees in (16384, 32768, 65536, 131072):
    s = RootParallelSearch("CartPole-v1", trees_per_rank=trees, n_sim=1000, n_roots=256, seed=0)
    g = torch.Generator().manual_seed(1234)
    roots = ((torch.rand((256, 4), generator=g, dtype=torch.float64) - 0.5) * 0.1)
    tr = s.expand_roots(roots)
    for _ in range(2): s.search_device(tr)
    torch.cuda.synchronize()
    ts = []
    for _ in range(3):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); s.search_device(tr); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    print("trees", trees, "ms", min(ts), "sims/s %.1fM" % (trees * 1000 / min(ts) / 1e3), flush=True)
    del s
    torch.cuda.empty_cache()
