#!/usr/bin/env python
"""CLI mirror of the reference's sweep.py: same flag names, types and defaults (sweep.py:22-47 of the reference;
--n_angles/--n_accelerations of sweep_vanilla.py:45-48 size DiscreteCurveEnv like sweep_vanilla.py:177).
--environment is effective here (the reference parses and ignores it and always builds CurveEnv, sweep.py:173-175):
a gym id, "CurveEnv" or "DiscreteCurveEnv".  wandb / plotting are out of scope; results are printed as JSON lines."""
from __future__ import annotations

import argparse
import json
import time

import numpy as np


def argument_parser():
    p = argparse.ArgumentParser(formatter_class=argparse.ArgumentDefaultsHelpFormatter)
    p.add_argument('--environment', default=None, type=str,
                   help='The environment to run: a gym id, "CurveEnv" or "DiscreteCurveEnv".  Default (like the reference, which '
                        'parses this flag, ignores it and always builds its CurveEnv, sweep.py:22,175): CurveEnv, or DiscreteCurveEnv '
                        'for the discrete-action algorithms (sweep_vanilla.py:177)')
    p.add_argument('--algorithm', default="vanilla", type=str, help='The algorithm to run')
    p.add_argument('--nsim', default=1000, type=int, help='The number of simulation the algorithm will run')
    p.add_argument('--c', default=1, type=float, help='exploration-exploitation factor')
    p.add_argument('--as', default="ucb", type=str, help='the function to select actions during simulations')
    p.add_argument('--ae', default="random", type=str, help='the function to select actions to add to the tree')
    p.add_argument('--gamma', default=0.5, type=float, help='discount factor')
    p.add_argument('--rollout', default='random', type=str, help='the function to select actions during rollout')
    p.add_argument('--max_depth', default=500, type=int, help='max number of steps during rollout')
    p.add_argument('--n_actions', default=10, type=int, help='number of actions for vanilla mcts')
    p.add_argument('--alpha1', default=0, type=float, help='alpha')
    p.add_argument('--alpha2', default=0.1, type=float, help='alpha2')
    p.add_argument('--k1', default=1, type=int, help='k1')
    p.add_argument('--k2', default=1, type=int, help='k2')
    p.add_argument('--epsilon', default=0.7, type=float, help='epsilon value for epsilon greedy strategies')
    p.add_argument('--genetic_default', default="random", type=str, help='the default policy for the genetic algorithm')
    p.add_argument('--n_sample', default=5, type=int, help='the number of samples taken by the genetic algorithm or by voo')
    p.add_argument('--n_episodes', default=100, type=int, help='the number of episodes for each experiment')
    p.add_argument('--group', default="", type=str, help='the group of the run')
    p.add_argument('--n_angles', default=5, type=int)
    p.add_argument('--n_accelerations', default=19, type=int)
    # build-only flags
    p.add_argument('--n_trees', default=1, type=int, help='root-parallel ensemble size')
    p.add_argument('--rollouts_per_leaf', default=1, type=int)
    p.add_argument('--tree_parallel', default=1, type=int,
                   help='simulations of ONE tree in flight at once (virtual loss); 1 = the reference\'s sequential loop')
    p.add_argument('--virtual_value', default=0.0, type=float, help='provisional return credited to a leaf whose rollouts are in flight')
    p.add_argument('--seed', default=None, type=int)
    p.add_argument('--depth_mode', default="budgeted", choices=["budgeted", "reference_head"])
    p.add_argument('--precision', default="f32", choices=["f32", "f64"])
    p.add_argument('--max_steps', default=100, type=int, help='real steps per episode (the reference stops at 100, sweep.py:206)')
    p.add_argument('--timeout_reward', default=-1000.0, type=float,
                   help='reward booked for the step that reaches --max_steps (the reference: -1000, sweep.py:206-209)')
    p.add_argument('--reuse_tree', action='store_true',
                   help='--lockstep only: keep the subtree below the action played instead of a fresh tree per real step (vanilla)')
    p.add_argument('--lockstep', action='store_true',
                   help='advance all --n_episodes episodes together on the device (islamcts_b200.episodes.EpisodeBatch)')
    return p


def get_function(name, args):
    from islamcts_b200.utils.action_selection_functions import (continuous_default_policy, genetic_policy, genetic_policy2,
                                                                ucb1, voo)
    if name == "ucb":
        return ucb1
    if name == "random":
        return continuous_default_policy
    if name == "voo":
        return voo(epsilon=args["epsilon"], default_policy=get_function(args["genetic_default"], args), n_samples=args["n_sample"])
    if name == "genetic":
        return genetic_policy(epsilon=args["epsilon"], default_policy=get_function(args["genetic_default"], args),
                              n_samples=args["n_sample"])
    if name == "genetic2":
        return genetic_policy2(epsilon=args["epsilon"], default_policy=get_function(args["genetic_default"], args))
    raise NotImplementedError(f"--{name}: unknown function name")


def create_agent(args, env, observation, seed=None):
    from islamcts_b200.agents import Mcts, MctsApw, MctsDpw, MctsHash, MctsSpw
    from islamcts_b200.agents.parameters.dpw_parameters import DpwParameters
    from islamcts_b200.agents.parameters.mcts_parameters import MctsParameters
    from islamcts_b200.agents.parameters.pw_parameters import PwParameters
    n_act = getattr(env.action_space, "n", None)
    common = dict(root_data=observation, env=env, n_sim=args["nsim"], C=args["c"], action_selection_fn=get_function(args["as"], args),
                  gamma=args["gamma"], rollout_selection_fn=get_function(args["rollout"], args), max_depth=args["max_depth"],
                  n_actions=n_act, x_values=None, y_values=None, depths=None, n_trees=args["n_trees"],
                  rollouts_per_leaf=args["rollouts_per_leaf"], tree_parallel=args["tree_parallel"], virtual_value=args["virtual_value"],
                  seed=seed, depth_mode=args["depth_mode"],
                  precision=args["precision"])
    algo = args["algorithm"]
    if algo == "vanilla":
        cls = Mcts if isinstance(observation, (int, np.integer)) else MctsHash
        return cls(MctsParameters(**common))
    if algo == "apw":
        return MctsApw(PwParameters(**common, alpha=args["alpha1"], k=args["k1"], action_expansion_function=get_function(args["ae"], args)))
    if algo == "spw":
        return MctsSpw(PwParameters(**common, alpha=args["alpha1"], k=args["k1"], action_expansion_function=None))
    if algo == "dpw":
        return MctsDpw(DpwParameters(**common, alphaApw=args["alpha1"], kApw=args["k1"], alphaSpw=args["alpha2"], kSpw=args["k2"]))
    raise NotImplementedError(algo)


def run_lockstep(args):
    """All episodes at once: one search launch per real step for the whole experiment, no host round trip in the loop."""
    from islamcts_b200 import _lib
    from islamcts_b200 import engine as E
    from islamcts_b200.episodes import EpisodeBatch
    name, algo = args["environment"], args["algorithm"]
    env_id = E.ENV_NAMES[name]
    n = args["n_episodes"]
    kw = dict(max_depth=args["max_depth"], C_=args["c"], gamma=args["gamma"], rollouts_per_leaf=args["rollouts_per_leaf"],
              precision=_lib.PRECISION_F64 if args["precision"] == "f64" else _lib.PRECISION_F32)
    grid = None
    if algo == "vanilla":
        agent = _lib.AGENT_VANILLA
        if name == "DiscreteCurveEnv":
            grid = (args["n_accelerations"], args["n_angles"])
        # MctsHash budgets its rollouts; Mcts (integer observations) plays zero-length rollouts at HEAD
        if env_id == _lib.ENV_FROZENLAKE and args["depth_mode"] == "reference_head":
            kw["budget_mode"] = _lib.BUDGET_ZERO
    elif algo == "apw":
        agent = _lib.AGENT_APW
        kw.update(alpha1=args["alpha1"], k1=args["k1"], epsilon=args["epsilon"], voo_n_samples=args["n_sample"],
                  expansion={"random": 0, "voo": 1, "genetic2": 2, "genetic": 3}[args["ae"]])
        if args["depth_mode"] == "reference_head":
            kw["budget_mode"] = _lib.BUDGET_ZERO
    else:
        raise NotImplementedError("--lockstep: vanilla and apw")
    seed = 0 if args["seed"] is None else args["seed"]
    if env_id == _lib.ENV_CURVE:
        states = np.tile(np.array([-11.0, 21.0, 10.0, 90.0, 0.0]), (n, 1))             # CurveEnv.reset(), curve_env.py:181
    elif env_id == _lib.ENV_FROZENLAKE:
        states = np.zeros((n, 1))
    else:
        from islamcts_b200 import envs
        rows = []
        for i in range(n):
            e = envs.make(name, api=4, seed=seed + i)
            e.reset()
            rows.append(np.asarray(e.state, dtype=np.float64))
        states = np.stack(rows)
    t0 = time.time()
    out = EpisodeBatch(env_id, agent=agent, n_episodes=n, n_sim=args["nsim"], max_steps=args["max_steps"], seed=seed,
                       action_grid=grid, timeout_reward=args["timeout_reward"],
                       reuse_tree=args["reuse_tree"] and algo == "vanilla" and env_id in (_lib.ENV_CARTPOLE, _lib.ENV_FROZENLAKE),
                       **kw).run(states)
    dt = time.time() - t0
    print(json.dumps({"episodes": n, "real_steps": out["real_steps"], "seconds": dt, "real_steps_per_s": out["real_steps"] / dt,
                      "mean_reward": float(out["total_reward"].mean()), "max_reward": float(out["total_reward"].max()),
                      "min_reward": float(out["total_reward"].min()), "mean_steps": float(out["n_steps"].mean())}), flush=True)


def main(argv=None):
    args = argument_parser().parse_args(argv).__dict__
    if args["environment"] is None:
        args["environment"] = "DiscreteCurveEnv" if args["algorithm"] in ("vanilla", "spw") else "CurveEnv"
    if args["lockstep"]:
        return run_lockstep(args)
    from islamcts_b200 import envs
    rewards = []
    for ep in range(args["n_episodes"]):
        ep_seed = None if args["seed"] is None else args["seed"] + ep
        if args["environment"] == "CurveEnv":
            from islamcts_b200.environments import CurveEnv
            env = CurveEnv(seed=ep_seed)                                                    # sweep.py:175
        elif args["environment"] == "DiscreteCurveEnv":
            from islamcts_b200.environments import DiscreteCurveEnv
            env = DiscreteCurveEnv([args["n_accelerations"], args["n_angles"]], seed=ep_seed)   # sweep_vanilla.py:177
        else:
            env = envs.make(args["environment"], api=4, seed=ep_seed)
        observation = env.reset()
        done, total_reward, n_steps, t0 = False, 0.0, 0, time.time()
        while not done and n_steps < args["max_steps"]:
            step_seed = None if args["seed"] is None else args["seed"] + 100003 * ep + n_steps
            agent = create_agent(args, env.unwrapped, observation, step_seed)   # a new agent (tree) every real step, sweep.py:186
            agent.root.data = observation
            action = agent.fit()
            observation, reward, done, _ = env.step(action)
            n_steps += 1
            if n_steps >= args["max_steps"]:
                reward = args["timeout_reward"]          # sweep.py:206-209: the step that reaches the limit books -1000
            total_reward += reward
        rewards.append(total_reward)
        print(json.dumps({"episode": ep, "total_reward": total_reward, "n_steps": n_steps, "seconds": time.time() - t0,
                          "n_actions": len(agent.q_values)}), flush=True)
    print(json.dumps({"mean_reward": float(np.mean(rewards)), "max_reward": float(np.max(rewards)),
                      "min_reward": float(np.min(rewards))}), flush=True)


if __name__ == '__main__':
    main()
