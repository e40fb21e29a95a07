"""TEST INFRASTRUCTURE ONLY -- pure-Python port of the reference's vanilla simulate loop.

Purpose: the CPU baseline ("kind": "port") that ``bench.py`` times on the GPU box, where
/root/reference does not exist.  It performs the same Python-level work per node as the reference
(object per node, dict children keyed by ``obs.tobytes()``, numpy UCB over tiny arrays,
``np.random.choice(np.flatnonzero(...))`` tie-breaks, one pickle clone of the env per simulation) so
that its speed represents the reference's own CPU path.  tests/test_py_port.py checks that, fed
the same random stream, it makes exactly the reference's draws and builds exactly the reference's
tree (tests/golden/hash_*.npz), and DESIGN.md records its measured speed next to the real reference's.

Follows: MctsHash.fit (islaMcts/agents/mcts_hash.py:22-50), StateNodeHash.build_tree_state (:59-83),
ActionNodeHash.build_tree_action (:87-131), AbstractStateNode.rollout (agents/abstract_mcts.py:72-91),
ucb1 (utils/action_selection_functions.py:9-25), my_deepcopy (utils/mcts_utils.py:8-16).
"""
from __future__ import annotations

import pickle

import numpy as np


class _Cfg:
    __slots__ = ("env", "n_sim", "C", "gamma", "max_depth", "n_actions", "zero_budget", "hash_keys")


class _SNode:
    __slots__ = ("data", "cfg", "terminal", "total", "ns", "actions", "visit_actions")

    def __init__(self, data, cfg):
        self.data = data
        self.cfg = cfg
        self.terminal = False
        self.total = 0
        self.ns = 0
        self.actions = {}
        self.visit_actions = np.zeros(cfg.n_actions)

    def ucb(self):
        kids = list(self.actions.values())
        visits, values = [], []
        for k in kids:
            visits.append(k.na)
            values.append(k.total / k.na)
        score = np.array(values) + self.cfg.C * np.sqrt(np.log(self.ns) / np.array(visits))
        pick = np.random.choice(np.flatnonzero(score == score.max()))
        return kids[pick].data

    def rollout(self, depth):
        env = self.cfg.env
        finished, ret, t = False, 0, 0
        while not finished and depth + t < self.cfg.max_depth:
            out = env.step(env.action_space.sample())
            ret += out[1] * pow(self.cfg.gamma, t)
            finished = out[2]
            t += 1
        return ret

    def visit(self, depth):
        if 0 in self.visit_actions:
            a = np.random.choice(np.flatnonzero(self.visit_actions == 0))
            kid = _ANode(a, self.cfg)
            self.actions[a] = kid
        else:
            a = self.ucb()
            kid = self.actions.get(a)
        g = kid.visit(depth)
        self.ns += 1
        self.visit_actions[a] += 1
        self.total += g
        return g


class _ANode:
    __slots__ = ("data", "cfg", "total", "na", "children")

    def __init__(self, data, cfg):
        self.data = data
        self.cfg = cfg
        self.total = 0
        self.na = 0
        self.children = {}

    @property
    def q_value(self):
        return self.total / self.na

    def visit(self, depth):
        cfg = self.cfg
        out = cfg.env.step(self.data)
        obs, r, term = out[0], out[1], out[2]
        key = obs.tobytes() if cfg.hash_keys else obs
        nxt = self.children.get(key, None)
        nd = depth + 1 if cfg.hash_keys else depth      # MctsHash counts levels; Mcts passes max_depth through
        if term:
            if nxt is None:
                nxt = _SNode(obs, cfg)
                nxt.terminal = True
                self.children[key] = nxt
            self.total += r
            self.na += 1
            nxt.ns += 1
            return r
        if nxt is None:
            nxt = _SNode(obs, cfg)
            self.children[key] = nxt
            g = r + cfg.gamma * nxt.rollout(nd)
            self.na += 1
            nxt.ns += 1
            self.total += g
            nxt.total += g
            return g
        g = r + cfg.gamma * nxt.visit(nd)
        self.total += g
        self.na += 1
        return g


class VanillaPort:
    """``VanillaPort(env, root_obs, ...).fit()`` == MctsHash(param).fit() (hash_keys) / Mcts(param).fit()."""

    def __init__(self, env, root_obs, n_sim, C, gamma, max_depth, n_actions, hash_keys=True):
        cfg = _Cfg()
        cfg.env, cfg.n_sim, cfg.C, cfg.gamma, cfg.max_depth, cfg.n_actions = env, n_sim, C, gamma, max_depth, n_actions
        cfg.hash_keys = hash_keys
        self.cfg = cfg
        self.root = _SNode(root_obs, cfg)
        self.q_values = None

    def _clone(self, src):
        env = pickle.loads(pickle.dumps(src))
        env.action_space = self.cfg.env.action_space   # the live space keeps its RNG stream (mcts_utils.py:15)
        return env

    def fit(self):
        cfg = self.cfg
        start = self._clone(cfg.env)
        for _ in range(cfg.n_sim):
            cfg.env = self._clone(start)
            self.root.visit(0 if cfg.hash_keys else cfg.max_depth)
        kids = list(self.root.actions.values())
        self.q_values = np.array([k.q_value for k in kids])
        top = self.q_values.max()
        return np.random.choice([k for k in kids if k.q_value == top]).data


def time_vanilla_fit(env_name="CartPole-v1", n_sim=1000, C=1.0, gamma=0.99, max_depth=500, seed=0):
    """One fit() on a fresh env; returns (seconds, sims, env_steps).  Used by bench.py's cpu_baseline."""
    import time

    from oracle import pyenvs

    env = pyenvs.make(env_name, api=4, seed=seed)
    obs = env.reset()
    np.random.seed(seed)
    n_actions = env.action_space.n
    port = VanillaPort(env, obs, n_sim, C, gamma, max_depth, n_actions, hash_keys=True)
    steps0 = pyenvs.STEP_COUNTER[0]
    t0 = time.perf_counter()
    port.fit()
    dt = time.perf_counter() - t0
    return dt, n_sim, pyenvs.STEP_COUNTER[0] - steps0
