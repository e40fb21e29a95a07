import sys, time, torch
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
    return best * 1e-3
roots = torch.tensor([[1.0, 0.3]], dtype=torch.float64)
for bt in (32, 0, 512):
    for bm in (_lib.BUDGET_TO_MAX_DEPTH, _lib.BUDGET_ZERO):
        eng = Engine(_lib.ENV_PENDULUM, agent=_lib.AGENT_APW, n_trees=1, sim_capacity=10000, max_depth=200, C_=1.0, gamma=0.99, alpha1=0.8,
                     k1=60.0, kernel=_lib.KERNEL_CTA_PER_TREE, rollouts_per_leaf=1, budget_mode=bm, seed=1, device=dev, block_threads=bt)
        t = timeit(lambda: eng.set_roots(roots).run(10000))
        c = eng.counters()
        print("C3 block_threads", bt, "budget_mode", bm, "%.2f ms  %.0f sims/s  env-steps/s %.3g" % (t * 1e3, 10000 / t, c["env_steps"] / t), c["faults"])
