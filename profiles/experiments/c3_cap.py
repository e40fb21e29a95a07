import sys, torch
sys.path.insert(0, ".")
from islamcts_b200 import _lib
from islamcts_b200.engine import Engine
dev = torch.device("cuda:0")
def timeit(fn, reps=3):
    fn(); torch.cuda.synchronize()
    best = 1e9
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize(); best = min(best, e0.elapsed_time(e1))
    return best
roots = torch.tensor([[1.0, 0.3]], dtype=torch.float64)
for cap in (10000, 10002, 20000, 40000):
    eng = Engine(_lib.ENV_PENDULUM, agent=_lib.AGENT_APW, n_trees=1, sim_capacity=cap, max_depth=200, C_=1.0, gamma=0.99, alpha1=0.8, k1=60.0, seed=1, device=dev)
    t_set = timeit(lambda: eng.set_roots(roots))
    t = timeit(lambda: eng.set_roots(roots).run(10000))
    print("cap", cap, "set_roots %.3f ms" % t_set, "set+run %.3f ms" % t, eng.counters())
