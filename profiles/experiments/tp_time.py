import sys, torch
sys.path.insert(0, ".")
from islamcts_b200 import _lib
from islamcts_b200.engine import Engine
dev = torch.device("cuda:0")
def timeit(fn, reps=2):
    fn(); torch.cuda.synchronize()
    best = 1e9
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize(); best = min(best, e0.elapsed_time(e1))
    return best * 1e-3
roots = torch.tensor([[-0.5, 0.0]], dtype=torch.float64)
for B in (1, 8, 32, 64):
    eng = Engine(_lib.ENV_MOUNTAINCAR, agent=_lib.AGENT_DPW, n_trees=1, sim_capacity=10000, max_depth=500, C_=1.0, gamma=0.99,
                 alpha1=0.5, k1=1.0, alpha2=0.5, k2=1.0, kernel=_lib.KERNEL_CTA_PER_TREE, seed=1, device=dev, tree_parallel=B, virtual_value=-5.0)
    t = timeit(lambda: eng.set_roots(roots).run(10000))
    a, na, tot = eng.root_children(0)
    print("C4 tree_parallel", B, "%.1f ms  %.0f sims/s" % (t * 1e3, 10000 / t), "root children", len(a), "faults", eng.counters()["faults"],
          "best q %.3f" % (tot / na).max())
cp = torch.tensor([[0.01, 0.0, 0.02, 0.0]], dtype=torch.float64)
for B in (1, 16, 32):
    eng = Engine(_lib.ENV_CARTPOLE, agent=_lib.AGENT_VANILLA, n_trees=1, sim_capacity=20000, max_depth=500, C_=1.0, gamma=0.99,
                 kernel=_lib.KERNEL_CTA_PER_TREE, rollouts_per_leaf=32, seed=1, device=dev, tree_parallel=B, virtual_value=0.0)
    t = timeit(lambda: eng.set_roots(cp).run(1000))
    print("CartPole 32 rollouts/leaf tree_parallel", B, "%.1f ms  %.0f sims/s" % (t * 1e3, 1000 / t), eng.counters()["faults"])
