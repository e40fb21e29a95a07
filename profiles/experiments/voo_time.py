import torch, time
from islamcts_b200 import _lib
from islamcts_b200.engine import Engine
dev = "cuda:0"
pose = torch.tensor([[-5.70993649, 25.49759526, 2.5, 0.0, 3.0]], dtype=torch.float64)
def timeit(fn, reps=3):
    fn(); torch.cuda.synchronize()
    best = 1e9
    for _ in range(reps):
        t0 = time.perf_counter(); fn(); torch.cuda.synchronize(); best = min(best, time.perf_counter() - t0)
    return best
for nt in (32, 128, 256, 512):
    eng = Engine(_lib.ENV_CURVE, agent=_lib.AGENT_APW, n_trees=1, sim_capacity=100, max_depth=500, C_=11.837, gamma=0.4, alpha1=0.0,
                 k1=95.0, expansion=_lib.EXPAND_VOO, epsilon=0.5, voo_n_samples=100, voo_max_try=1000,
                 kernel=_lib.KERNEL_CTA_PER_TREE, seed=1, device=dev, block_threads=nt)
    t = timeit(lambda: eng.set_roots(pose).run(100))
    print("voo example_curve nt", nt, "ms", t * 1e3, "sims/s", 100 / t, eng.counters())
# sweep.py-like voo: n_sample 5
eng = Engine(_lib.ENV_CURVE, agent=_lib.AGENT_APW, n_trees=1, sim_capacity=1000, max_depth=500, C_=11.0, gamma=0.95, alpha1=0.5,
             k1=3.0, expansion=_lib.EXPAND_VOO, epsilon=0.7, voo_n_samples=5, voo_max_try=1000, kernel=2, seed=1, device=dev)
t = timeit(lambda: eng.set_roots(pose).run(1000))
print("voo sweep defaults ms", t * 1e3, "sims/s", 1000 / t)
eng = Engine(_lib.ENV_PENDULUM, agent=_lib.AGENT_APW, n_trees=1, sim_capacity=1000, max_depth=200, C_=1.0, gamma=0.99, alpha1=0.5,
             k1=3.0, expansion=_lib.EXPAND_VOO, epsilon=0.7, voo_n_samples=5, voo_max_try=1000, kernel=2, seed=1, device=dev)
t = timeit(lambda: eng.set_roots(torch.tensor([[1.0, 0.3]], dtype=torch.float64)).run(1000))
print("voo pendulum ms", t * 1e3, "sims/s", 1000 / t)
