import sys, torch
sys.path.insert(0, ".")
from islamcts_b200 import _lib
from islamcts_b200.engine import Engine
dev = torch.device("cuda:0")
x = torch.randn(8192, 8192, device=dev)
for _ in range(20):
    x = x @ x * 1e-4
torch.cuda.synchronize()
for nt in (1, 1024, 65536):
    for tun in (0, 2):
        eng = Engine(_lib.ENV_FROZENLAKE, n_trees=nt, sim_capacity=100, max_depth=1000, C_=0.5, gamma=1.0, kernel=_lib.KERNEL_THREAD_PER_TREE, seed=1, device=dev, tuning=tun)
        roots = torch.zeros((nt, 1), dtype=torch.float64, device=dev)
        eng.set_roots(roots).run(100); torch.cuda.synchronize()
        best = 1e9
        for _ in range(5):
            eng.set_roots(roots); torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); eng.run(100); e1.record(); torch.cuda.synchronize(); best = min(best, e0.elapsed_time(e1))
        print("FrozenLake trees", nt, "tuning", tun, "%.3f ms  %.1f M sims/s" % (best, nt * 100 / best / 1e3), eng.counters()["faults"], flush=True)
