"""TEST INFRASTRUCTURE ONLY -- freeze reference runs as golden fixtures.

Runs the reference's OWN agents (imported from /root/reference through
oracle/ref_loader.py, unmodified) on the float64 env restatement
(oracle/pyenvs.py) with every random draw recorded (oracle/trace.py), and writes
``tests/golden/<name>.npz`` holding the configuration, the decision trace, the
resulting tree in canonical DFS form, the root statistics and the action
``fit()`` returned.  The C oracle and the CUDA engine replay the trace and must
reproduce the tree: integers bit-exact, float64 totals to 1e-9 relative.

Usage (build container only; /root/reference does not exist on the GPU box):
    python -m oracle.gen_golden            # writes all fixtures
"""
from __future__ import annotations

import json
import os
import sys

import numpy as np

from oracle import pyenvs, ref_loader
from oracle import trace as T

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")

# name -> scenario.  agent: reference class key in ref_loader.load(); budget: "to_max_depth" | "zero"
# describes what that class does AT HEAD.
SCENARIOS = {
    # C2-like: the only agent with working depth-budgeted rollouts
    "hash_cartpole": dict(agent="MctsHash", env="CartPole-v1", n_sim=150, C=1.0, gamma=0.99, max_depth=500,
                          budget="to_max_depth", env_seed=0, rng_seed=1),
    # small max_depth + large C: exercises budget exhaustion, wide exploration and many UCB ties
    "hash_cartpole_shallow": dict(agent="MctsHash", env="CartPole-v1", n_sim=200, C=8.0, gamma=0.9, max_depth=6,
                                  budget="to_max_depth", env_seed=3, rng_seed=2),
    # C1: vanilla Mcts on FrozenLake (zero-length rollouts at HEAD)
    "mcts_lake": dict(agent="Mcts", env="FrozenLake-v1", n_sim=100, C=0.5, gamma=1.0, max_depth=1000,
                      budget="zero", env_seed=0, rng_seed=3),
    "mcts_lake_long": dict(agent="Mcts", env="FrozenLake-v1", n_sim=600, C=0.5, gamma=0.95, max_depth=1000,
                           budget="zero", env_seed=0, rng_seed=4),
    # MctsHash on FrozenLake (obs wrapped as np.int64 so that .tobytes() exists): integer env + real rollouts
    "hash_lake": dict(agent="MctsHash", env="FrozenLake-v1", n_sim=300, C=0.5, gamma=0.9, max_depth=30,
                      budget="to_max_depth", env_seed=0, rng_seed=5),
    # C3-like: APW on Pendulum; k=60, alpha=0.8 widens on every simulation (tree depth 1)
    "apw_pendulum_c3": dict(agent="MctsApw", env="Pendulum-v1", n_sim=200, C=1.0, gamma=0.99, max_depth=200,
                            alpha=0.8, k=60, budget="zero", env_seed=0, rng_seed=6),
    # narrow widening so that UCB fires and the tree gets deep
    "apw_pendulum_narrow": dict(agent="MctsApw", env="Pendulum-v1", n_sim=300, C=2.0, gamma=0.95, max_depth=200,
                                alpha=0.4, k=1, budget="zero", env_seed=1, rng_seed=7),
    "apw_voo_pendulum": dict(agent="MctsApw", env="Pendulum-v1", n_sim=120, C=2.0, gamma=0.95, max_depth=200,
                             alpha=0.5, k=2, budget="zero", env_seed=2, rng_seed=8,
                             expansion="voo", epsilon=0.5, voo_n_samples=3, voo_max_try=50),
    # APW with the genetic2 expansion (--ae genetic2): centre / bounds first, then epsilon-mix of random and the mean
    # of the two lowest-q actions
    "apw_genetic2_pendulum": dict(agent="MctsApw", env="Pendulum-v1", n_sim=160, C=2.0, gamma=0.95, max_depth=200,
                                  alpha=0.6, k=2, budget="zero", env_seed=3, rng_seed=13, expansion="genetic2", epsilon=0.6),
    # APW with the first genetic expansion (--ae genetic, n_sample points between the two lowest-q actions)
    "apw_genetic_pendulum": dict(agent="MctsApw", env="Pendulum-v1", n_sim=170, C=2.0, gamma=0.95, max_depth=200,
                                 alpha=0.6, k=2, budget="zero", env_seed=4, rng_seed=18, expansion="genetic", epsilon=0.6,
                                 voo_n_samples=5),
    "apw_genetic_curve": dict(agent="MctsApw", env="CurveEnv", n_sim=200, C=11.0, gamma=0.9, max_depth=50, alpha=0.45, k=3,
                              budget="zero", env_seed=0, rng_seed=19, pre_steps=[[-5.0, -30.0], [-5.0, -30.0], [-4.0, -25.0]],
                              expansion="genetic", epsilon=0.5, voo_n_samples=4),
    # CurveEnv is the reference's OWN environment (islaMcts/environments/curve_env.py) -- the env sweep.py runs.
    # APW on the continuous 2-D action box, after two real steps so that the hidden step counter is > 0
    "apw_curve": dict(agent="MctsApw", env="CurveEnv", n_sim=250, C=11.0, gamma=0.99, max_depth=50, alpha=0.5, k=3,
                      budget="zero", env_seed=0, rng_seed=14, pre_steps=[[-5.0, -30.0], [-5.0, -30.0], [-4.0, -25.0]]),
    # examples/example_curve.py:32-45 runs APW + voo on CurveEnv (C=11.837, gamma=0.4, 100 voo samples): 2-D Voronoi sampling
    "apw_voo_curve": dict(agent="MctsApw", env="CurveEnv", n_sim=200, C=11.837, gamma=0.4, max_depth=50, alpha=0.3, k=4,
                          budget="zero", env_seed=0, rng_seed=16, pre_steps=[[-5.0, -30.0], [-5.0, -30.0], [-4.0, -25.0]],
                          expansion="voo", epsilon=0.5, voo_n_samples=12, voo_max_try=40),
    "apw_genetic2_curve": dict(agent="MctsApw", env="CurveEnv", n_sim=220, C=11.0, gamma=0.9, max_depth=50, alpha=0.45, k=3,
                               budget="zero", env_seed=0, rng_seed=17, pre_steps=[[-5.0, -30.0], [-5.0, -30.0], [-4.0, -25.0]],
                               expansion="genetic2", epsilon=0.5),
    # vanilla MctsHash on DiscreteCurveEnv([5 accelerations, 7 angles]) (sweep_vanilla.py:177), real rollouts
    "hash_curve_discrete": dict(agent="MctsHash", env="DiscreteCurveEnv", n_sim=300, C=5.0, gamma=0.95, max_depth=12,
                                budget="to_max_depth", env_seed=0, rng_seed=15, grid=[5, 7], pre_steps=[0, 0, 0]),
    # MctsHash on FrozenLake with the heuristic rollout policy of examples/example_lake.py:22-66: grid_policy over the
    # Manhattan distance of the cell each action leads to (ties broken by np.random.choice)
    "hash_lake_grid_policy": dict(agent="MctsHash", env="FrozenLake-v1", n_sim=220, C=0.5, gamma=0.95, max_depth=14,
                                  budget="to_max_depth", env_seed=5, rng_seed=20, rollout="grid"),
    # FrozenLake is_slippery=True: my_deepcopy clones the env WITH its generator for every simulation
    # (utils/mcts_utils.py:8-16), so the k-th step of every simulation draws the same noise; Mcts (zero rollouts at HEAD)
    # and MctsHash (budgeted rollouts)
    "mcts_lake_slippery": dict(agent="Mcts", env="FrozenLake-v1", n_sim=250, C=0.5, gamma=0.95, max_depth=30, budget="zero",
                               env_seed=7, rng_seed=22, slippery=True),
    "hash_lake_slippery": dict(agent="MctsHash", env="FrozenLake-v1", n_sim=220, C=0.5, gamma=0.95, max_depth=16,
                               budget="to_max_depth", env_seed=8, rng_seed=23, slippery=True),
    # SPW on the slippery lake steps the LIVE env (no clone): every env.step draws fresh noise, so an action node gets
    # several children -- widening for real (k=1), and resampling among them with weights na / ns (k=0.5, alpha=0.2)
    "spw_lake_slippery": dict(agent="MctsSpw", env="FrozenLake-v1", n_sim=300, C=0.5, gamma=0.95, max_depth=30, alpha=0.5, k=1,
                              budget="zero", env_seed=11, rng_seed=24, slippery=True),
    "spw_lake_slippery_resample": dict(agent="MctsSpw", env="FrozenLake-v1", n_sim=400, C=0.7, gamma=0.9, max_depth=30, alpha=0.2,
                                       k=0.5, budget="zero", env_seed=12, rng_seed=25, slippery=True),
    # SPW (with the state_variable shim) on CartPole: k=1 always re-expands; k=0.4 resamples and descends
    "spw_cartpole_k1": dict(agent="MctsSpw", env="CartPole-v1", n_sim=150, C=1.0, gamma=0.99, max_depth=500,
                            alpha=0.5, k=1, budget="zero", env_seed=0, rng_seed=9),
    "spw_cartpole_descend": dict(agent="MctsSpw", env="CartPole-v1", n_sim=200, C=1.0, gamma=0.9, max_depth=500,
                                 alpha=0.3, k=0.4, budget="zero", env_seed=4, rng_seed=10),
    # C4-like: DPW (state_variable + index-UCB shims) on MountainCarContinuous
    "dpw_mountaincar_c4": dict(agent="MctsDpw", env="MountainCarContinuous-v0", n_sim=400, C=1.0, gamma=0.99,
                               max_depth=500, alpha1=0.5, k1=1, alpha2=0.5, k2=1, budget="zero", env_seed=0,
                               rng_seed=11),
    "dpw_mountaincar_descend": dict(agent="MctsDpw", env="MountainCarContinuous-v0", n_sim=300, C=0.5, gamma=0.95,
                                    max_depth=500, alpha1=0.4, k1=1, alpha2=0.2, k2=0.5, budget="zero",
                                    env_seed=5, rng_seed=12),
}


def lake_manhattan_knowledge():
    """The prior knowledge of examples/example_lake.py:22-66: for every cell and action, the Manhattan distance to the goal
    (3, 3) of the cell the action leads to (moves off the grid stay put).  key = row * 4 + col."""
    def dist(r, c, a):
        nr, nc = (r, c - 1) if a == 0 else (r + 1, c) if a == 1 else (r, c + 1) if a == 2 else (r - 1, c)
        if nc < 0 or nc > 3 or nr < 0 or nr > 3:
            nr, nc = r, c
        return abs(3 - nc) + abs(3 - nr)
    return {r * 4 + c: [dist(r, c, a) for a in range(4)] for c in range(4) for r in range(4)}


class _LakeHashEnv(pyenvs.FrozenLakeEnv):
    """FrozenLake whose observation is np.int64 (has .tobytes()) so MctsHash can key it."""

    def step(self, a):
        out = super().step(a)
        return (np.int64(out[0]),) + tuple(out[1:])


def run_scenario(name: str, sc: dict):
    R = ref_loader.load()
    env_name = sc["env"]
    agent_key = sc["agent"]
    api = 5 if agent_key == "Mcts" else 4  # mcts.py:79-82 indexes vals[0..2]; the others unpack 4 (mcts_hash.py:93)
    if env_name in ("CurveEnv", "DiscreteCurveEnv"):
        env = R.CurveEnv() if env_name == "CurveEnv" else R.DiscreteCurveEnv(sc["grid"])
        env.action_space.seed(sc["env_seed"])
        obs = env.reset()
        for a in sc.get("pre_steps", []):
            obs, _, done, _ = env.step(np.asarray(a, dtype=np.float64) if env_name == "CurveEnv" else int(a))
            assert not done, "pre_steps must keep the car on the track"
        root_state = np.array([*env.car.state, float(env.n_step)], dtype=np.float64)
    elif env_name == "FrozenLake-v1" and agent_key in ("MctsHash", "MctsSpw"):
        env = _LakeHashEnv(api=4, seed=sc["env_seed"], is_slippery=sc.get("slippery", False))
        obs = env.reset()
        root_state = pyenvs.env_state_vector(env)
    elif env_name == "FrozenLake-v1" and sc.get("slippery"):
        env = pyenvs.FrozenLakeEnv(api=api, seed=sc["env_seed"], is_slippery=True)
        obs = env.reset()[0]
        root_state = pyenvs.env_state_vector(env)
    else:
        env = pyenvs.make(env_name, api=api, seed=sc["env_seed"])
        obs = env.reset()
        if api == 5:
            obs = obs[0]
        root_state = pyenvs.env_state_vector(env)
    tr = T.Trace()
    env.action_space = T.RecordingSpace(env.action_space, tr)
    n_actions = getattr(env.action_space, "n", None)
    root_data = obs
    if env_name == "FrozenLake-v1" and agent_key == "MctsHash":
        root_data = np.array([0, 0])  # fit() indexes root_data[0], root_data[1] (mcts_hash.py:37-38)
    if env_name == "FrozenLake-v1" and agent_key == "MctsSpw":
        root_data = np.int64(obs)     # poked into env.s by fit() (mcts_state_...:24-25)
    rollout_fn = R.continuous_default_policy
    if sc.get("rollout") == "grid":
        rollout_fn = R.asf.grid_policy(lake_manhattan_knowledge(), 4)
    common = dict(root_data=root_data, env=env, n_sim=sc["n_sim"], C=sc["C"], action_selection_fn=R.ucb1,
                  gamma=sc["gamma"], rollout_selection_fn=rollout_fn, max_depth=sc["max_depth"],
                  n_actions=n_actions, x_values=[], y_values=[], depths=[])
    if agent_key in ("Mcts", "MctsHash"):
        param = R.MctsParameters(**common)
    elif agent_key == "MctsApw":
        if sc.get("expansion") == "voo":
            ae = R.voo(sc["epsilon"], R.continuous_default_policy, sc["voo_n_samples"], sc["voo_max_try"])
        elif sc.get("expansion") == "genetic2":
            ae = R.asf.genetic_policy2(sc["epsilon"], R.continuous_default_policy)
        elif sc.get("expansion") == "genetic":
            ae = R.asf.genetic_policy(sc["epsilon"], R.continuous_default_policy, sc["voo_n_samples"])
        else:
            ae = R.continuous_default_policy
        param = R.PwParameters(**common, alpha=sc["alpha"], k=sc["k"], action_expansion_function=ae)
    elif agent_key == "MctsSpw":
        param = R.SpwParams(**common, alpha=sc["alpha"], k=sc["k"], action_expansion_function=None,
                            state_variable="s" if env_name == "FrozenLake-v1" else "state")
    elif agent_key == "MctsDpw":
        common["action_selection_fn"] = R.ucb1_index
        param = R.DpwParams(**common, alphaApw=sc["alpha1"], kApw=sc["k1"], alphaSpw=sc["alpha2"], kSpw=sc["k2"],
                            state_variable="state")
    else:
        raise KeyError(agent_key)
    agent = getattr(R, agent_key)(param)
    # what every simulation's cloned env will draw, step by step (slippery FrozenLake)
    noise = None
    if sc.get("slippery"):
        import copy
        # (SPW steps the live env: its draws run on through the whole fit(), one per tree level at HEAD)
        noise = copy.deepcopy(env.np_random).random(sc["max_depth"] + 4 if agent_key != "MctsSpw" else sc["n_sim"] * 64)
    fit_error = ""
    steps0 = pyenvs.STEP_COUNTER[0]
    with T.recording(tr, sc["rng_seed"]):
        try:
            action = agent.fit()
        except ValueError as e:
            # mcts_double_...:41 decodes a 4-byte float32 key with dtype=float (8 bytes): HEAD bug.
            fit_error = str(e)
            action = None
    kinds, vals = tr.arrays()
    assert kinds[-1] == T.K_CHOICE, "fit() ends with the arg-max tie-break pick"
    ser = T.serialise_reference_tree(agent.root)
    ch = list(agent.root.actions.values())
    root_action = np.array([float(np.asarray(c.data, dtype=np.float64).reshape(-1)[0]) for c in ch])
    root_action2 = np.array([float(np.asarray(c.data, dtype=np.float64).reshape(-1)[1]) if np.size(c.data) > 1 else 0.0 for c in ch])
    root_na = np.array([c.na for c in ch], dtype=np.int64)
    root_total = np.array([c.total for c in ch], dtype=np.float64)
    if action is None:
        q = root_total / root_na
        best = np.flatnonzero(q == q.max())[int(vals[-1])]
        action_val = root_action[best]
    elif agent_key == "MctsSpw":
        action_val = float(root_action[int(action)])  # fit returns an INDEX into the sorted actions (:38)
    else:
        action_val = float(np.asarray(action, dtype=np.float64).reshape(-1)[0])
    # the sweep loop's per-step metrics (sweep.py:199-200): the reference's own get_depth_max / get_depth_mean
    # (agents/abstract_mcts.py:120-142,190-209) on every root action node
    root_depth_max = np.array([c.get_depth_max(0) for c in ch], dtype=np.int32)
    root_depth_mean = np.array([float(c.get_depth_mean(0, True)) for c in ch], dtype=np.float64)
    out = dict(
        config=np.array(json.dumps(dict(sc, name=name, n_actions=n_actions, fit_error=fit_error))),
        root_state=root_state,
        trace_kind=kinds, trace_value=vals,
        q_values=np.asarray(agent.q_values, dtype=np.float64),
        root_action=root_action, root_action2=root_action2, root_na=root_na, root_total=root_total,
        root_depth_max=root_depth_max, root_depth_mean=root_depth_mean,
        # (the live-env table of SPW is cut to what the fit() consumed)
        **({"noise": noise if agent_key != "MctsSpw" else noise[: int(pyenvs.STEP_COUNTER[0] - steps0) + 8]} if noise is not None else {}),
        fit_action=np.float64(action_val),
        env_steps=np.int64(pyenvs.STEP_COUNTER[0] - steps0) if env_name not in ("CurveEnv", "DiscreteCurveEnv") else np.int64(-1),
        **{f"tree_{k}": v for k, v in ser.items()},
    )
    return out


def curve_env_vectors(n_walks=400, seed=77):
    """Step vectors of the reference's OWN CurveEnv (islaMcts/environments/curve_env.py:121-177):
    random walks from reset() and from random on-track poses, recording (state+n_step, action) ->
    (next state, reward, terminal).  Pins the env arithmetic of the oracle / device CurveEnv."""
    R = ref_loader.load()
    rng = np.random.default_rng(seed)
    S, A, NS, RW, TM = [], [], [], [], []
    for w in range(n_walks):
        env = R.CurveEnv()
        env.reset()
        if w % 2:
            x = rng.uniform(-11.0, 9.0)
            lo, hi = env.lower_bound(x), env.higher_bound(x)
            env.car.position[:] = (x, rng.uniform(lo, hi))
            env.car.velocity = rng.uniform(0.0, 8.0)
            env.car.angle = rng.uniform(-60.0, 100.0) % 360
            env.car.state = np.array([*env.car.position, env.car.velocity, env.car.angle])
            env.n_step = int(rng.integers(0, 30))
        for _ in range(12):
            s = np.array([*env.car.state, float(env.n_step)])
            a = rng.uniform([-5.0, -30.0], [5.0, 30.0]) if w % 4 < 2 else rng.uniform([-5.0, -30.0], [0.0, 0.0])
            ns, r, done, _ = env.step(a)
            S.append(s); A.append(a); NS.append(np.array([*ns, float(env.n_step)])); RW.append(float(r)); TM.append(bool(done))
            if done:
                break
    # DiscreteCurveEnv index -> (acceleration, angle) table (curve_env.py:180-195)
    d = R.DiscreteCurveEnv([5, 7])
    out = dict(state=np.array(S), action=np.array(A), next_state=np.array(NS), reward=np.array(RW),
               terminal=np.array(TM), grid=np.array([5, 7]), grid_actions=np.array(d.actions, dtype=np.float64))
    os.makedirs(os.path.join(GOLDEN_DIR, "env"), exist_ok=True)
    path = os.path.join(GOLDEN_DIR, "env", "curve_env_steps.npz")
    np.savez_compressed(path, **out)
    print(f"curve_env_steps: {len(S)} steps, terminal {np.mean(TM):.2f}, wins {(np.array(RW) > 100).sum()} -> {path}")


def main(names=None):
    os.makedirs(GOLDEN_DIR, exist_ok=True)
    if not names or "curve_env_steps" in names:
        curve_env_vectors()
    for name, sc in SCENARIOS.items():
        if names and name not in names:
            continue
        out = run_scenario(name, sc)
        path = os.path.join(GOLDEN_DIR, name + ".npz")
        np.savez_compressed(path, **out)
        print(f"{name}: trace={len(out['trace_kind'])} records={len(out['tree_kind'])} "
              f"max_depth={out['tree_depth'].max()} root_children={len(out['root_na'])} "
              f"fit_action={float(out['fit_action']):.6g} size={os.path.getsize(path)}B")


if __name__ == "__main__":
    main(sys.argv[1:])
