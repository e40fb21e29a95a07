import os, sys, torch
sys.path.insert(0, ".")
from islamcts_b200 import _lib
from islamcts_b200.engine import Engine
dev = "cuda:0"
pose = torch.tensor([[-5.70993649, 25.49759526, 2.5, 0.0, 3.0]], dtype=torch.float64)
def timeit(fn, reps=2):
    fn(); torch.cuda.synchronize()
    best = 1e9
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize(); best = min(best, e0.elapsed_time(e1))
    return best * 1e-3
for tuning in (0, 16):
    for nt in (148, 1024, 4096):
        eng = Engine(_lib.ENV_CURVE, agent=_lib.AGENT_VANILLA, n_trees=nt, sim_capacity=1000, max_depth=500, C_=11.0, gamma=0.95,
                     action_grid=(5, 19), kernel=2, seed=1, device=dev, tuning=tuning)
        roots = pose.repeat(nt, 1)
        t = timeit(lambda: eng.set_roots(roots).run(1000))
        print("curve vanilla tuning", tuning, "trees", nt, "ms %.1f" % (t * 1e3), "sims/s %.2fM" % (nt * 1000 / t / 1e6), flush=True)
    for nt in (100, 1024):
        eng = Engine(_lib.ENV_CARTPOLE, n_trees=nt, sim_capacity=1000, max_depth=500, gamma=0.99, kernel=2, seed=1, device=dev, tuning=tuning)
        roots = (torch.rand((nt, 4), dtype=torch.float64) - 0.5) * 0.1
        t = timeit(lambda: eng.set_roots(roots).run(1000))
        print("cartpole cta tuning", tuning, "trees", nt, "ms %.1f" % (t * 1e3), "sims/s %.2fM" % (nt * 1000 / t / 1e6), flush=True)
    eng = Engine(_lib.ENV_PENDULUM, agent=_lib.AGENT_APW, n_trees=1024, sim_capacity=1000, max_depth=200, gamma=0.99, alpha1=0.5, k1=3.0, kernel=2, seed=1, device=dev, tuning=tuning)
    roots = torch.tensor([[1.0, 0.3]], dtype=torch.float64).repeat(1024, 1)
    t = timeit(lambda: eng.set_roots(roots).run(1000))
    print("apw pendulum tuning", tuning, "trees 1024 ms %.1f" % (t * 1e3), "sims/s %.2fM" % (1024 * 1000 / t / 1e6), flush=True)
