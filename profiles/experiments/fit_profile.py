import sys, time, cProfile, pstats, io
sys.path.insert(0, ".")
import numpy as np, torch
from islamcts_b200 import envs
from islamcts_b200.agents import MctsApw
from islamcts_b200.agents.parameters.pw_parameters import PwParameters
from islamcts_b200.utils.action_selection_functions import continuous_default_policy, ucb1
env = envs.make("Pendulum-v1", api=4, seed=3)
obs = env.reset()
def make():
    return MctsApw(PwParameters(root_data=obs, env=env.unwrapped, n_sim=10000, C=1.0, action_selection_fn=ucb1, gamma=0.99,
                                rollout_selection_fn=continuous_default_policy, max_depth=200, n_actions=None, alpha=0.8, k=60,
                                action_expansion_function=continuous_default_policy, seed=5))
for _ in range(3):
    make().fit()
torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(10):
    make().fit()
print("ms per fit", (time.perf_counter() - t0) * 100)
pr = cProfile.Profile(); pr.enable()
for _ in range(10):
    make().fit()
pr.disable()
s = io.StringIO(); pstats.Stats(pr, stream=s).sort_stats("cumulative").print_stats(25); print(s.getvalue()[:4500])
