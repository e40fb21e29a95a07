import sys, time, torch, numpy as np
sys.path.insert(0, ".")
from islamcts_b200 import envs, _lib
from islamcts_b200.engine import Engine
from islamcts_b200.agents.mcts_hash import MctsHash
from islamcts_b200.agents.parameters.mcts_parameters import MctsParameters
from islamcts_b200.utils.action_selection_functions import continuous_default_policy, ucb1
env = envs.make("CartPole-v1", api=4, seed=3); obs = env.reset()
def T(): torch.cuda.synchronize(); return time.perf_counter()
for rep in range(3):
    t0 = T()
    p = MctsParameters(root_data=obs, env=env, n_sim=1000, C=1, action_selection_fn=ucb1, gamma=0.99, rollout_selection_fn=continuous_default_policy, max_depth=500, n_actions=2, seed=rep)
    ag = MctsHash(p); t1 = T()
    a = ag.fit(); t2 = T()
    n = len(ag.root.actions); d = ag.root.actions[int(a)].get_depth_max(0); t3 = T()
    obs2, r, done, _ = env.step(a); t4 = T()
    print("agent %.2f ms | fit %.2f | root metrics %.2f | env.step %.2f" % ((t1-t0)*1e3, (t2-t1)*1e3, (t3-t2)*1e3, (t4-t3)*1e3))
# inside fit
root = np.asarray(env.state, dtype=np.float64)
for rep in range(3):
    t0 = T(); eng = Engine(0, n_trees=1, sim_capacity=4000, max_depth=500, gamma=0.99, seed=rep); t1 = T()
    eng.set_roots(root[None, :]); t2 = T(); eng.run(1000); t3 = T(); c = eng.counters(); t4 = T(); rk, act = eng.best(); x = act.cpu(); t5 = T(); rc = eng.root_children(0); t6 = T()
    print("Engine() %.2f | set_roots %.2f | run %.2f | counters %.2f | best %.2f | root_children %.2f" % tuple((b-a)*1e3 for a, b in ((t0,t1),(t1,t2),(t2,t3),(t3,t4),(t4,t5),(t5,t6))))
for cap in (1000, 4000):
    eng = Engine(0, n_trees=1, sim_capacity=cap, max_depth=500, gamma=0.99, seed=1)
    eng.set_roots(root[None, :]).run(1000); t0 = T(); eng.set_roots(root[None, :]).run(1000); t1 = T()
    print("capacity", cap, "run 1000 sims %.2f ms" % ((t1-t0)*1e3), "pool_in_smem?", eng.layout.pool_stride)
