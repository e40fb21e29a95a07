"""The drop-in Python API end to end on the device: README example, return types, proxies, statistics."""
import numpy as np
import pytest

from oracle import coracle

pytestmark = pytest.mark.gpu


def test_readme_example_frozenlake_reaches_goal():
    """README.md:99-121 of the reference, verbatim flow, with the device-backed env instead of gym."""
    from islamcts_b200 import envs
    from islamcts_b200.agents.mcts import Mcts
    from islamcts_b200.agents.parameters.mcts_parameters import MctsParameters
    from islamcts_b200.utils.action_selection_functions import continuous_default_policy, ucb1
    np.random.seed(0)
    real_env = envs.make("FrozenLake-v1").unwrapped
    observation, _ = real_env.reset()
    parameters = MctsParameters(C=0.5, n_sim=100, root_data=observation, env=real_env, action_selection_fn=ucb1,
                                max_depth=1000, gamma=1, rollout_selection_fn=continuous_default_policy,
                                n_actions=real_env.action_space.n)
    done, steps, reward = False, 0, 0.0
    while not done and steps < 30:
        parameters.env = real_env
        agent = Mcts(param=parameters)
        action = agent.fit()
        assert isinstance(action, np.int64) and 0 <= action < 4
        observation, reward, done, _, _ = real_env.step(action)
        parameters.root_data = observation
        steps += 1
    assert done and reward == 1.0 and observation == 15      # the agent reaches the goal


def test_mctshash_cartpole_contract():
    from islamcts_b200 import envs
    from islamcts_b200.agents.mcts_hash import MctsHash
    from islamcts_b200.agents.parameters.mcts_parameters import MctsParameters
    from islamcts_b200.utils.action_selection_functions import continuous_default_policy, ucb1
    env = envs.make("CartPole-v1", api=4, seed=3)
    obs = env.reset()
    p = MctsParameters(root_data=obs, env=env, n_sim=300, C=1, action_selection_fn=ucb1, gamma=0.99,
                       rollout_selection_fn=continuous_default_policy, max_depth=500, n_actions=2, seed=11)
    agent = MctsHash(p)
    agent.root.data = obs                                   # sweep.py:189 assigns this before fit()
    a = agent.fit()
    assert isinstance(a, np.int64) and a in (0, 1)
    assert agent.q_values.shape == (2,) and agent.q_values[int(np.argmax(agent.q_values))] == agent.q_values.max()
    root = agent.root
    assert root.ns == 300 and set(root.actions) == {0, 1}
    assert sum(c.na for c in root.actions.values()) == 300
    assert root.total == pytest.approx(sum(c.total for c in root.actions.values()), rel=1e-12)
    node = root.actions[int(a)]
    assert node.q_value == pytest.approx(agent.q_values.max())
    assert node.get_depth_max(0) >= 2 and node.get_depth_mean(0, True) >= 1
    # invariants of SURVEY 8(a) on the whole proxy tree: non-root non-terminal ns = 1 + sum na ; na = sum child ns
    stack = [(root, True)]
    while stack:
        s, is_root = stack.pop()
        if not s.terminal and s.actions:
            assert s.ns == (0 if is_root else 1) + sum(c.na for c in s.actions.values())
        for c in s.actions.values():
            assert c.na == sum(ch.ns for ch in c.children.values())
            stack.extend((ch, False) for ch in c.children.values())
    agent.fit()                                             # calling fit() twice continues the same tree
    assert agent.root.ns == 600
    # the oracle with the same Philox key builds the same root statistics (f32 env: identical here)
    t = coracle.Tree(coracle.make_config(0, n_actions=2, max_depth=500, C_=1.0, gamma=0.99, seed=11), env.state)
    t.run(600)
    ea, ena, etot = t.root_children()
    np.testing.assert_array_equal(agent.root_statistics["na"], ena)


def test_pw_agents_return_types_and_keys():
    from islamcts_b200 import envs
    from islamcts_b200.agents import MctsApw, MctsDpw, MctsSpw
    from islamcts_b200.agents.parameters.dpw_parameters import DpwParameters
    from islamcts_b200.agents.parameters.pw_parameters import PwParameters
    from islamcts_b200.utils.action_selection_functions import continuous_default_policy, ucb1, voo
    common = dict(n_sim=400, C=1, action_selection_fn=ucb1, gamma=0.95, rollout_selection_fn=continuous_default_policy,
                  max_depth=40, seed=5)
    env = envs.make("Pendulum-v1", api=4, seed=1)
    obs = env.reset()
    for ae in (continuous_default_policy, voo(0.7, continuous_default_policy, 5)):
        ag = MctsApw(PwParameters(root_data=obs, env=env, n_actions=None, alpha=0.5, k=1, action_expansion_function=ae, **common))
        a = ag.fit()
        assert isinstance(a, np.ndarray) and a.shape == (1,) and -2 <= a[0] <= 2
        assert a.tobytes() in ag.root.actions or ae is not continuous_default_policy   # sweep.py:196 lookup
        assert len(ag.root.actions) == len(ag.q_values) and ag.root.ns == 400
    env = envs.make("MountainCarContinuous-v0", api=4, seed=2)
    obs = env.reset()
    ag = MctsDpw(DpwParameters(root_data=obs, env=env, n_actions=None, alphaApw=0.5, kApw=1, alphaSpw=0.5, kSpw=1,
                               state_variable="state", **common))
    a = ag.fit()
    assert a.dtype == np.float32 and a.tobytes() in ag.root.actions
    keys = list(ag.root.actions)
    assert keys == sorted(keys)                            # dpw re-orders root.actions by key bytes (:30)
    env = envs.make("CartPole-v1", api=4, seed=3)
    obs = env.reset()
    ag = MctsSpw(PwParameters(root_data=obs, env=env, n_actions=2, alpha=0.3, k=0.4, action_expansion_function=None,
                              state_variable="state", **common))
    idx = ag.fit()
    assert isinstance(idx, int) and idx in (0, 1)


def test_root_q_statistics_agree_with_oracle_within_ci():
    """north_star: root Q estimates (an RNG-independent statistic) agree within a stated confidence interval.
    256 device trees (fp32 env, Philox key A) vs 256 oracle trees (float64 env, key B): per-action mean root Q
    within 3.5 standard errors (99.95% two-sided) and mean visit share within 3.5 s.e."""
    from islamcts_b200.engine import Engine
    n_trees, n_sim = 256, 200
    root = coracle.env_reset(0, 123, 0)
    roots = np.tile(root[None, :], (n_trees, 1))
    eng = Engine(0, n_trees=n_trees, sim_capacity=n_sim, max_depth=500, C_=1.0, gamma=0.99, precision=0, seed=1001)
    eng.set_roots(roots).run(n_sim)
    na, tot = eng.root_stats()
    na, tot = na.cpu().numpy(), tot.cpu().numpy()
    cfg = coracle.make_config(0, n_actions=2, max_depth=500, C_=1.0, gamma=0.99, seed=2002)
    ena, etot, _ = coracle.run_many(cfg, roots, n_sim, n_threads=8)
    for a in range(2):
        qg, qc = tot[:, a] / na[:, a], etot[:, a] / ena[:, a]
        se = np.sqrt(qg.var(ddof=1) / n_trees + qc.var(ddof=1) / n_trees)
        assert abs(qg.mean() - qc.mean()) < 3.5 * se, (a, qg.mean(), qc.mean(), se)
        sg, sc = na[:, a] / n_sim, ena[:, a] / n_sim
        se = np.sqrt(sg.var(ddof=1) / n_trees + sc.var(ddof=1) / n_trees)
        assert abs(sg.mean() - sc.mean()) < 3.5 * se + 1e-9


def test_ensemble_fit_merges_root_statistics():
    from islamcts_b200 import envs
    from islamcts_b200.agents.mcts_hash import MctsHash
    from islamcts_b200.agents.parameters.mcts_parameters import MctsParameters
    from islamcts_b200.utils.action_selection_functions import continuous_default_policy, ucb1
    env = envs.make("CartPole-v1", api=4, seed=8)
    obs = env.reset()
    p = MctsParameters(root_data=obs, env=env, n_sim=100, C=1, action_selection_fn=ucb1, gamma=0.99,
                       rollout_selection_fn=continuous_default_policy, max_depth=200, n_actions=2, seed=4, n_trees=2048)
    ag = MctsHash(p)
    a = ag.fit()
    assert a in (0, 1) and ag.root_statistics["na"].sum() == 2048 * 100


def test_curve_env_sweep_flow_apw_and_vanilla():
    """sweep.py:174-199 / sweep_vanilla.py:176-200 of the reference on its own CurveEnv: a fresh agent per real step,
    ``agent.root.actions[action.tobytes()]`` / ``agent.root.actions[action]`` look-ups, depth metrics, env.step."""
    from islamcts_b200.agents import MctsApw, MctsHash
    from islamcts_b200.agents.parameters.mcts_parameters import MctsParameters
    from islamcts_b200.agents.parameters.pw_parameters import PwParameters
    from islamcts_b200.environments import CurveEnv, DiscreteCurveEnv
    from islamcts_b200.utils.action_selection_functions import continuous_default_policy, ucb1
    common = dict(n_sim=300, C=11, action_selection_fn=ucb1, gamma=0.95, rollout_selection_fn=continuous_default_policy,
                  max_depth=30, seed=21, precision="f64")
    # --- APW on the continuous box
    env = CurveEnv()
    obs = env.reset()
    for step in range(3):
        p = PwParameters(root_data=obs, env=env.unwrapped, n_actions=None, alpha=0.5, k=3,
                         action_expansion_function=continuous_default_policy, **common)
        agent = MctsApw(p)
        agent.root.data = obs
        action = agent.fit()
        assert action.dtype == np.float64 and action.shape == (2,)
        assert -5 <= action[0] <= 5 and -30 <= action[1] <= 30
        node = agent.root.actions[action.tobytes()]
        assert node.q_value == pytest.approx(agent.q_values.max())
        assert node.get_depth_max(0) >= 0 and len(agent.root.actions) == len(agent.q_values)
        # same Philox key -> the oracle builds the same root (float64 env arithmetic on both sides)
        t = coracle.Tree(coracle.make_config(4, agent=1, max_depth=30, C_=11.0, gamma=0.95, alpha1=0.5, k1=3.0, seed=21),
                         [*env.car.state, env.n_step])
        t.run(300)
        ea, ena, etot = t.root_children()
        np.testing.assert_array_equal(agent.root_statistics["na"], ena)
        np.testing.assert_array_equal(agent.root_statistics["action"][:, 0], ea)
        np.testing.assert_array_equal(agent.root_statistics["action"][:, 1], t.root_action2)
        np.testing.assert_allclose(agent.root_statistics["total"], etot, rtol=1e-9)
        n_before = env.n_step
        obs, reward, done, extra = env.step(np.array([-5.0, -30.0]))   # brake + turn right keeps the car on the track
        assert extra is None and env.n_step == n_before + 1 and not done
        eobs, er, _, _ = coracle.env_step(4, [*p.root_data, n_before], [-5.0, -30.0])
        np.testing.assert_allclose(obs, eobs[:4], rtol=1e-13)
        assert reward == pytest.approx(er, rel=1e-13)
    env.reset()
    assert env.n_step == 3                    # reset() re-creates the car but keeps the step counter (curve_env.py:179-182)
    # --- vanilla MctsHash on the discrete grid (sweep_vanilla.py:177 uses [n_accelerations, n_angles])
    denv = DiscreteCurveEnv([5, 7])
    obs = denv.reset()
    for a in (0, 0, 0):
        obs, _, done, _ = denv.step(a)
    assert not done and denv.actions[0] == (-5.0, -30.0) and denv.action_space.n == 35
    p = MctsParameters(root_data=obs, env=denv.unwrapped, n_actions=denv.action_space.n, **common)
    agent = MctsHash(p)
    agent.root.data = obs
    action = agent.fit()
    assert isinstance(action, np.int64) and 0 <= action < 35
    node = agent.root.actions[action]
    assert node.q_value == pytest.approx(agent.q_values.max())
    assert agent.root.ns == 300 and len(agent.root.actions) == 35
    t = coracle.Tree(coracle.make_config(4, agent=0, n_actions=35, grid0=5, grid1=7, max_depth=30, C_=11.0, gamma=0.95, seed=21),
                     [*denv.car.state, denv.n_step])
    t.run(300)
    ea, ena, etot = t.root_children()
    np.testing.assert_array_equal(agent.root_statistics["na"], ena)
    np.testing.assert_array_equal(agent.root_statistics["action"], ea)
    obs2, reward, done, _ = denv.step(action)
    assert obs2.shape == (4,)
    # root-parallel ensemble over the 35 grid actions
    p8 = MctsParameters(root_data=obs, env=denv.unwrapped, n_actions=35, n_trees=8, **common)
    a8 = MctsHash(p8).fit()
    assert isinstance(a8, np.int64) and 0 <= a8 < 35


def test_reference_curve_env_objects_are_recognised():
    """A maintainer passes the REFERENCE's env object as param.env; only attributes are read (car.state, n_step,
    action_space, actions).  Stand-ins with exactly those attributes are used here (the reference is not on the GPU box)."""
    import itertools
    import types
    from islamcts_b200.agents import MctsHash
    from islamcts_b200.agents.parameters.mcts_parameters import MctsParameters
    from islamcts_b200.utils.action_selection_functions import continuous_default_policy, ucb1
    class DiscreteCurveEnv:     # same class name as the reference's
        pass
    env = DiscreteCurveEnv()
    env.car = types.SimpleNamespace(state=np.array([-5.70993649, 25.49759526, 2.5, 0.0]))
    env.n_step = 3
    env.actions = list(itertools.product(np.linspace(-5.0, 5.0, 3), np.linspace(-30.0, 30.0, 4)))
    env.action_space = types.SimpleNamespace(n=12)
    p = MctsParameters(root_data=env.car.state, env=env, n_sim=200, C=5, action_selection_fn=ucb1, gamma=0.9,
                       rollout_selection_fn=continuous_default_policy, max_depth=20, n_actions=12, seed=2, precision="f64")
    agent = MctsHash(p)
    a = agent.fit()
    t = coracle.Tree(coracle.make_config(4, agent=0, n_actions=12, grid0=3, grid1=4, max_depth=20, C_=5.0, gamma=0.9, seed=2),
                     [*env.car.state, 3.0])
    t.run(200)
    ea, ena, etot = t.root_children()
    np.testing.assert_array_equal(agent.root_statistics["na"], ena)
    rc, rank, eact = t.best()
    # the final arg-max draw continues the tree stream on both sides
    assert int(a) == int(eact)
    env.actions[5] = (0.123, 4.0)                   # not the linspace grid any more -> refuse, never guess
    with pytest.raises(NotImplementedError):
        MctsHash(p).fit()


def test_root_proxies_serve_sweep_metrics_from_device_analytics():
    """sweep.py:191-200 reads len(root.actions), root.actions[action] and its depth metrics every real step: those come
    from two small device arrays; the full tree is decoded only when something deeper is touched, and agrees."""
    from islamcts_b200 import envs
    from islamcts_b200.agents import MctsApw, MctsHash
    from islamcts_b200.agents.abstract_mcts import _RootActionProxy
    from islamcts_b200.agents.parameters.mcts_parameters import MctsParameters
    from islamcts_b200.agents.parameters.pw_parameters import PwParameters
    from islamcts_b200.utils.action_selection_functions import continuous_default_policy, ucb1
    common = dict(n_sim=500, C=1, action_selection_fn=ucb1, gamma=0.95, rollout_selection_fn=continuous_default_policy,
                  max_depth=60, seed=4)
    cases = []
    env = envs.make("CartPole-v1", api=4, seed=1); obs = env.reset()
    cases.append(MctsHash(MctsParameters(root_data=obs, env=env, n_actions=2, **common)))
    env = envs.make("Pendulum-v1", api=4, seed=1); obs = env.reset()
    cases.append(MctsApw(PwParameters(root_data=obs, env=env, n_actions=None, alpha=0.4, k=2,
                                      action_expansion_function=continuous_default_policy, **common)))
    for agent in cases:
        action = agent.fit()
        root = agent.root
        assert agent._full_proxy is None                       # nothing decoded yet
        key = action.tobytes() if isinstance(action, np.ndarray) else action
        node = root.actions[key]
        assert isinstance(node, _RootActionProxy)
        dmax, dmean, q = node.get_depth_max(0), node.get_depth_mean(0, True), node.q_value
        assert agent._full_proxy is None and len(root.actions) == len(agent.q_values)
        assert q == pytest.approx(agent.q_values.max())
        full = agent._full_root()                              # the reference's recursion on the decoded tree
        for k, n in root.actions.items():
            assert n.get_depth_max(0) == full.actions[k].get_depth_max(0)
            assert n.get_depth_mean(0, True) == pytest.approx(full.actions[k].get_depth_mean(0, True), rel=1e-12)
            assert n.get_depth_max(3) == full.actions[k].get_depth_max(3)
            assert n.na == full.actions[k].na and n.total == full.actions[k].total
            assert sorted(n.get_depth_mean(0)) == sorted(full.actions[k].get_depth_mean(0))
        assert root.ns == 500 and dmax >= 1 and dmean >= 1


def test_apw_ensemble_unions_the_trees_root_children():
    """n_trees > 1 for a continuous-action agent: every tree samples its own root actions; fit() returns the arg-max over
    the union of the trees' root children (equal actions summed)."""
    from islamcts_b200 import envs
    from islamcts_b200.agents import MctsApw
    from islamcts_b200.agents.parameters.pw_parameters import PwParameters
    from islamcts_b200.utils.action_selection_functions import continuous_default_policy, genetic_policy2, ucb1
    env = envs.make("Pendulum-v1", api=4, seed=2); obs = env.reset()
    p = PwParameters(root_data=obs, env=env, n_sim=200, C=2, action_selection_fn=ucb1, gamma=0.95, max_depth=30,
                     rollout_selection_fn=continuous_default_policy, n_actions=None, alpha=0.5, k=2,
                     action_expansion_function=genetic_policy2(0.5, continuous_default_policy), n_trees=6, seed=9)
    agent = MctsApw(p)
    a = agent.fit()
    st = agent.root_statistics
    assert a.shape == (1,) and st["na"].sum() == 6 * 200
    assert len(np.unique(st["action"])) == len(st["action"])           # merged: no duplicates left
    per_tree = sum(len(agent._engine.root_children(i)[1]) for i in range(6))
    assert len(st["action"]) <= per_tree - 3 * 5                       # centre and both bounds exist in all 6 trees: summed
    best = int(np.argmax(st["total"] / st["na"]))
    assert float(a[0]) == float(st["action"][best, 0])


def test_example_lake_with_grid_policy_rollouts():
    """examples/example_lake.py:56-97: grid_policy over the Manhattan-distance table as rollout_selection_fn.  Mcts at HEAD
    never rolls out (depth_mode='reference_head'); in the budgeted mode the heuristic guides the rollouts."""
    from islamcts_b200 import envs
    from islamcts_b200.agents.mcts import Mcts
    from islamcts_b200.agents.parameters.mcts_parameters import MctsParameters
    from islamcts_b200.utils.action_selection_functions import grid_policy, ucb1
    from tests.golden_util import lake_manhattan_table
    distances = {s: list(lake_manhattan_table()[s]) for s in range(16)}
    rollout_fn = grid_policy(distances, 4)
    real_env = envs.make("FrozenLake-v1").unwrapped
    observation, _ = real_env.reset()
    done, steps, reward = False, 0, 0.0
    while not done and steps < 20:
        p = MctsParameters(C=0.5, n_sim=100, root_data=observation, env=real_env, action_selection_fn=ucb1, max_depth=1000,
                           gamma=1, rollout_selection_fn=rollout_fn, n_actions=4, depth_mode="budgeted", seed=steps)
        agent = Mcts(param=p)
        action = agent.fit()
        observation, reward, done, _, _ = real_env.step(action)
        steps += 1
    assert done and reward == 1.0 and steps == 6            # the shortest path to the goal
    from islamcts_b200.agents.mcts_hash import MctsHash
    cart = envs.make("CartPole-v1", api=4, seed=1); obs = cart.reset()
    with pytest.raises(NotImplementedError):
        MctsHash(MctsParameters(C=1, n_sim=10, root_data=obs, env=cart, action_selection_fn=ucb1, max_depth=10, gamma=0.9,
                                rollout_selection_fn=rollout_fn, n_actions=2)).fit()


def test_slippery_frozenlake_through_the_agent_api():
    """gym.make("FrozenLake-v1", is_slippery=True): fit() clones the env generator's next draws (every simulation of the
    reference steps a pickled clone) and the search equals the oracle fed with the same draws."""
    import copy
    from islamcts_b200 import envs
    from islamcts_b200.agents.mcts_hash import MctsHash
    from islamcts_b200.agents.parameters.mcts_parameters import MctsParameters
    from islamcts_b200.utils.action_selection_functions import continuous_default_policy, ucb1
    env = envs.make("FrozenLake-v1", api=4, seed=12, is_slippery=True)
    obs = env.reset()
    total_slips = 0
    for step in range(4):
        p = MctsParameters(C=0.5, n_sim=150, root_data=obs, env=env, action_selection_fn=ucb1, max_depth=25, gamma=0.95,
                           rollout_selection_fn=continuous_default_policy, n_actions=4, seed=40 + step, precision="f64")
        noise = copy.deepcopy(env.np_random).random(700)
        agent = MctsHash(p)
        a = agent.fit()
        t = coracle.Tree(coracle.make_config(3, n_actions=4, max_depth=25, C_=0.5, gamma=0.95, seed=40 + step), [float(env.s)])
        assert t.set_noise(noise) == 0 and t.run(150) == 0
        ea, ena, etot = t.root_children()
        np.testing.assert_array_equal(agent.root_statistics["na"], ena)
        np.testing.assert_allclose(agent.root_statistics["total"], etot, rtol=1e-12)
        before = env.s
        obs, r, done, _ = env.step(a)
        det = envs.make("FrozenLake-v1", api=4); det.reset(); det.s = before
        total_slips += int(det.step(a)[0] != obs)
        if done:
            break
    assert 0 <= total_slips <= 4


def test_spw_agent_on_slippery_lake_steps_the_live_env():
    """MctsSpw on FrozenLake is_slippery=True through the drop-in classes: the search consumes the env's own generator
    (the reference does not clone the env in spw / dpw), action nodes get several children keyed by observation, and
    after fit() the env's generator stands where the reference's would: advanced by the env.step calls of the search."""
    import copy
    from islamcts_b200 import envs
    from islamcts_b200.agents import MctsSpw
    from islamcts_b200.agents.parameters.pw_parameters import PwParameters
    from islamcts_b200.utils.action_selection_functions import continuous_default_policy, ucb1
    env = envs.make("FrozenLake-v1", api=4, seed=5, is_slippery=True)
    obs = env.reset()
    before = copy.deepcopy(env.np_random)
    p = PwParameters(root_data=obs, env=env.unwrapped, n_sim=300, C=0.5, action_selection_fn=ucb1, gamma=0.95,
                     rollout_selection_fn=continuous_default_policy, max_depth=20, n_actions=4, alpha=0.5, k=1,
                     action_expansion_function=None, depth_mode="reference_head", seed=3, state_variable="s")
    ag = MctsSpw(p)
    a = ag.fit()
    assert 0 <= int(a) < 4 and len(ag.q_values) == 4
    root = ag._full_root()
    fan = [len(act.children) for act in root.actions.values()]
    assert max(fan) >= 2                                           # several next states under one root action
    assert root.ns == 300            # (an spw action node's na does not count the visits that resample and descend, :126-130)
    steps = ag._engine.counters()["env_steps"]
    assert steps >= 300
    before.random(int(steps))
    assert env.np_random.random() == before.random()               # the live generator moved by exactly the search's steps


def test_proxy_children_are_keyed_like_the_reference():
    """action_node.children is a dict keyed by observation.tobytes() in the *Hash agents (mcts_hash.py:97) and by the
    observation itself in Mcts (agents/mcts.py:83-84); child.data is that observation (round-1 advisor finding: the
    proxies keyed children by position)."""
    from islamcts_b200 import envs
    from islamcts_b200.agents import Mcts, MctsHash
    from islamcts_b200.agents.parameters.mcts_parameters import MctsParameters
    from islamcts_b200.utils.action_selection_functions import continuous_default_policy, ucb1
    env = envs.make("CartPole-v1", api=4, seed=2)
    obs = env.reset()
    ag = MctsHash(MctsParameters(root_data=obs, env=env.unwrapped, n_sim=200, C=1, action_selection_fn=ucb1, gamma=0.95,
                                 rollout_selection_fn=continuous_default_policy, max_depth=50, n_actions=2, seed=4))
    a = ag.fit()
    node = ag.root.actions[int(a)]
    nxt_obs, _, _, _ = env.step(a)                                   # the real env, float64 dynamics
    (key, child), = node.children.items()
    assert isinstance(key, bytes) and child.data.dtype == np.float32 and key == child.data.tobytes()
    np.testing.assert_allclose(child.data, nxt_obs, rtol=1e-5, atol=1e-6)      # same observation up to float32 env rounding
    lake = envs.make("FrozenLake-v1", api=5, seed=0)
    o, _ = lake.reset()
    m = Mcts(MctsParameters(root_data=o, env=lake.unwrapped, n_sim=60, C=0.5, action_selection_fn=ucb1, gamma=1.0,
                            rollout_selection_fn=continuous_default_policy, max_depth=50, n_actions=4, seed=1, depth_mode="budgeted"))
    m.fit()
    right = m.root.actions[2]
    assert list(right.children.keys()) == [1] and right.children[1].data == 1   # RIGHT from cell 0 leads to cell 1
