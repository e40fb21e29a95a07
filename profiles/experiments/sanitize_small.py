"""Small runs of every kernel family for compute-sanitizer (memcheck)."""
import sys
sys.path.insert(0, ".")
import numpy as np, torch
from islamcts_b200 import _lib
from islamcts_b200.engine import Engine, env_step, rollout
from islamcts_b200.episodes import EpisodeBatch

rng = np.random.default_rng(0)
# thread-per-tree: CartPole (deep, cooperative phase) and FrozenLake (shallow), partial warps
for env_id, S in ((0, 4), (3, 1)):
    e = Engine(env_id, n_trees=100, sim_capacity=300, max_depth=200, gamma=0.99, kernel=1, seed=3)
    roots = np.zeros((100, S)) if env_id == 3 else rng.uniform(-0.05, 0.05, (100, 4))
    e.set_roots(roots).run(150); e.run(150)
    print("tpt", env_id, e.counters(), e.depth_stats(3)[0])
    e.best(); e.root_stats()
# CTA engine: every agent / expansion, R = 1 and 64
pose = np.array([[-5.70993649, 25.49759526, 2.5, 0.0, 3.0]])
cases = [dict(env_id=0, agent=0), dict(env_id=1, agent=1, alpha1=0.5, k1=2.0), dict(env_id=1, agent=1, alpha1=0.5, k1=2.0, expansion=1, epsilon=0.4, voo_n_samples=6, voo_max_try=9),
         dict(env_id=1, agent=1, alpha1=0.5, k1=2.0, expansion=2, epsilon=0.5), dict(env_id=1, agent=1, alpha1=0.5, k1=2.0, expansion=3, epsilon=0.5, voo_n_samples=4),
         dict(env_id=0, agent=2, alpha1=0.3, k1=0.4), dict(env_id=2, agent=3, alpha1=0.4, k1=1.0, alpha2=0.2, k2=0.5),
         dict(env_id=4, agent=1, alpha1=0.5, k1=3.0), dict(env_id=4, agent=1, alpha1=0.3, k1=4.0, expansion=1, epsilon=0.5, voo_n_samples=10, voo_max_try=12),
         dict(env_id=4, agent=0, action_grid=(5, 7))]
for kw in cases:
    for R in (1, 64):
        kw2 = dict(kw); env_id = kw2.pop("env_id")
        e = Engine(env_id, n_trees=3, sim_capacity=120, max_depth=30, gamma=0.95, C_=2.0, kernel=2, seed=5, rollouts_per_leaf=R, **kw2)
        S = e.state_dim
        roots = np.tile(pose, (3, 1)) if env_id == 4 else rng.uniform(-0.05, 0.05, (3, S))
        e.set_roots(roots).run(120)
        e.best(); e.root_children(1); e.depth_stats(2)
        if e.counters()["faults"]:
            print("faults in", kw, R, [e.tree_meta(i)["fault"] for i in range(3)])
# leaf batches larger than the CTA (rollouts handed out dynamically), tree-parallel batches
for env_id, kw in ((0, dict(agent=0)), (3, dict(agent=0)), (1, dict(agent=1, alpha1=0.5, k1=2.0))):
    e = Engine(env_id, n_trees=2, sim_capacity=60, max_depth=40, gamma=0.95, kernel=2, seed=9, rollouts_per_leaf=200, block_threads=64, **kw)
    e.set_roots(np.zeros((2, e.state_dim)) if env_id == 3 else rng.uniform(-0.05, 0.05, (2, e.state_dim))).run(60)
    e.best()
    assert e.counters()["faults"] == 0
e = Engine(2, n_trees=1, sim_capacity=400, max_depth=60, gamma=0.99, kernel=2, seed=2, agent=3, alpha1=0.5, k1=1.0, alpha2=0.5, k2=1.0, tree_parallel=16, virtual_value=-5.0)
e.set_roots(np.array([[-0.5, 0.0]])).run(400); e.best()
print("cta ok")
ns, r, t = env_step(4, np.tile(pose, (64, 1)), rng.uniform([-5, -30], [5, 30], (64, 2)), precision=1)
for env_id in range(5):
    rollout(env_id, None, n=5000, budget=50, gamma=0.9, seed=1)
    rollout(env_id, None, n=70001, budget=7, gamma=0.9, seed=2)
rollout(4, None, n=5000, budget=50, gamma=0.9, seed=1, action_grid=(5, 7))
eb = EpisodeBatch("CartPole-v1", n_episodes=20, n_sim=60, max_steps=5, seed=1, max_depth=30)
print(eb.run(rng.uniform(-0.05, 0.05, (20, 4)))["n_steps"])
torch.cuda.synchronize()
print("done")
