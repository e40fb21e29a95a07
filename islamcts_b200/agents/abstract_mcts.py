"""Host-side mirror of islaMcts/agents/abstract_mcts.py: the agent API in front of the device engine.

``Agent(param).fit()`` keeps the reference's contract (agents/abstract_mcts.py:12-49, agents/mcts_hash.py:22-50):
synchronous, returns the action of the arg-max-q root child, ``fit()`` again continues the same tree, and
afterwards ``agent.root`` / ``agent.q_values`` can be inspected.  The simulate loop itself runs inside
libisla_b200.so; the node objects the reference builds while searching are replaced by lazy proxies decoded
from the device tree on first access.
"""
from __future__ import annotations

from typing import Any

import numpy as np
import torch

from .. import _lib
from ..engine import ENV_NAMES, N_ACTIONS, STATE_DIM, Engine
from .parameters.mcts_parameters import MctsParameters

_LAKE_MAP = ("SFFF", "FHFH", "FFFH", "HFFG")


def identify_env(env) -> int:
    """gym-style env object -> engine env id.  Anything else: NotImplementedError (no CPU fallback)."""
    u = getattr(env, "unwrapped", env)
    name = getattr(getattr(u, "spec", None), "id", None) or getattr(u, "spec_id", None) or type(u).__name__
    low = str(name).lower()
    if "cartpole" in low:
        return _lib.ENV_CARTPOLE
    if "pendulum" in low:
        return _lib.ENV_PENDULUM
    if "mountaincar" in low and ("continuous" in low or hasattr(getattr(u, "action_space", None), "low")):
        return _lib.ENV_MOUNTAINCAR
    if "curve" in low and hasattr(u, "car"):
        return _lib.ENV_CURVE                          # the reference's own islaMcts/environments/curve_env.py
    if "frozenlake" in low:
        desc = getattr(u, "desc", None)
        if desc is not None:
            rows = tuple("".join(c.decode() if isinstance(c, bytes) else str(c) for c in r) for r in np.asarray(desc))
            if rows != _LAKE_MAP:
                raise NotImplementedError(f"FrozenLake map {rows} is not on the device (only the default 4x4 map)")
        return _lib.ENV_FROZENLAKE
    raise NotImplementedError(f"environment {name!r} has no device implementation (supported: {sorted(ENV_NAMES)})")


def lake_transition_noise(env, n: int):
    """FrozenLake is_slippery=True: what the next n env.step calls of a CLONE of the env would draw.  The reference
    steps a pickled clone (generator state included) in every simulation (utils/mcts_utils.py:8-16), so all simulations
    of one fit() see this same sequence.  None for the deterministic lake."""
    import copy
    u = getattr(env, "unwrapped", env)
    slippery = getattr(u, "is_slippery", None)
    if slippery is None:
        P = getattr(u, "P", None)
        slippery = P is not None and len(P[0][0]) != 1
    if not slippery:
        return None
    rng = getattr(u, "np_random", None)
    if rng is None or not hasattr(rng, "random"):
        raise NotImplementedError("slippery FrozenLake without a numpy Generator in env.np_random")
    return np.asarray(copy.deepcopy(rng).random(int(n)), dtype=np.float64)


def observation_of(env_id: int, state) -> Any:
    """The observation the reference's env returns for the engine's env state (what node.data holds and what the
    children dicts are keyed with): CartPole / MountainCar float32 state, Pendulum (cos, sin, thetadot) float32,
    FrozenLake the cell as an int, CurveEnv the car's float64 (x, y, velocity, angle)."""
    s = np.asarray(state, dtype=np.float64).reshape(-1)
    if env_id == _lib.ENV_FROZENLAKE:
        return int(s[0])
    if env_id == _lib.ENV_PENDULUM:
        return np.array([np.cos(s[0]), np.sin(s[0]), s[1]], dtype=np.float32)
    if env_id == _lib.ENV_CURVE:
        return np.array(s[:4], dtype=np.float64)
    return np.asarray(s[: STATE_DIM[env_id]], dtype=np.float32)


def env_action_info(env, env_id: int):
    """(number of discrete actions or 0, DiscreteCurveEnv grid or None) of the env object."""
    if env_id != _lib.ENV_CURVE:
        return N_ACTIONS[env_id], None
    u = getattr(env, "unwrapped", env)
    n = getattr(getattr(u, "action_space", None), "n", None)
    if n is None:
        return 0, None
    table = np.asarray(getattr(u, "actions", ()), dtype=np.float64).reshape(-1, 2)
    if len(table) != int(n) or n < 1:
        raise NotImplementedError("discrete CurveEnv without the env.actions table of DiscreteCurveEnv")
    g1 = int(np.sum(table[:, 0] == table[0, 0])) if len(np.unique(table[:, 0])) > 1 else len(table)
    g0 = len(table) // g1
    want = np.array([(a, b) for a in np.linspace(-5.0, 5.0, g0) for b in np.linspace(-30.0, 30.0, g1)]).reshape(-1, 2)
    if g0 * g1 != len(table) or not np.array_equal(table, want):
        raise NotImplementedError("env.actions is not linspace(-5, 5, g0) x linspace(-30, 30, g1) (curve_env.py:190-198)")
    if len(table) > 128:
        raise NotImplementedError(f"DiscreteCurveEnv with {len(table)} actions: the device engine holds at most 128")
    return len(table), (g0, g1)


def env_root_state(env, env_id: int) -> np.ndarray:
    u = getattr(env, "unwrapped", env)
    if env_id == _lib.ENV_FROZENLAKE:
        return np.array([float(int(u.s))], dtype=np.float64)
    if env_id == _lib.ENV_CURVE:
        return np.array([*np.asarray(u.car.state, dtype=np.float64).reshape(-1)[:4], float(u.n_step)], dtype=np.float64)
    st = getattr(u, "state", None)
    if st is None:
        raise ValueError("env.unwrapped.state is None: call env.reset() before fit()")
    return np.asarray(st, dtype=np.float64).reshape(-1)[: STATE_DIM[env_id]].copy()


def _kind(fn, what):
    k = getattr(fn, "isla_kind", None)
    if k is None and getattr(fn, "__name__", "") in ("ucb1", "continuous_default_policy"):
        k = {"ucb1": "ucb1", "continuous_default_policy": "random"}[fn.__name__]   # the reference's own functions
    if k is None:
        raise NotImplementedError(f"{what}={fn!r} cannot run on the device: use the callables of "
                                  f"islamcts_b200.utils.action_selection_functions (no CPU fallback by design)")
    return k


# --------------------------------------------------------------------------- lazy node proxies
class AbstractStateNode:
    __slots__ = ("data", "param", "terminal", "total", "ns", "actions", "terminal_reward", "visit_actions")

    def __init__(self, data: Any, param):
        self.data, self.param = data, param
        self.terminal, self.total, self.ns = False, 0, 0
        self.actions: dict = {}
        self.terminal_reward = None
        self.visit_actions = None

    def get_depth_max(self, depth: int = 0):
        # agents/abstract_mcts.py:120-130
        depth += 1
        depths = [a.get_depth_max(depth) for a in self.actions.values()]
        return max(depths) if depths else depth

    def get_depth_mean(self, depth: int = 0):
        # agents/abstract_mcts.py:132-142
        depth += 1
        depths = []
        for a in self.actions.values():
            depths.extend(a.get_depth_mean(depth))
        return depths if depths else [depth]


class AbstractActionNode:
    __slots__ = ("data", "total", "na", "children", "param")

    def __init__(self, data: Any, param):
        self.data, self.param = data, param
        self.total, self.na = 0, 0
        self.children: dict = {}

    @property
    def q_value(self):
        return self.total / self.na   # ZeroDivisionError when na == 0, like the reference (:154-156)

    def get_depth_max(self, depth):
        depths = [s.get_depth_max(depth) for s in self.children.values()]
        return max(depths) if depths else depth

    def get_depth_mean(self, depth=0, root=False):
        depths = []
        for s in self.children.values():
            depths.extend(s.get_depth_mean(depth))
        if root:
            return float(np.mean(depths)) if depths else depth
        return depths


class _RootActionProxy(AbstractActionNode):
    """Root action node served from two small device arrays (root_children + depth_stats): what the sweep loop reads
    every real step (sweep.py:191-200) -- na, total, q_value, get_depth_max(0), get_depth_mean(0, True) -- needs no
    copy of the whole tree.  ``children`` (and the list form of get_depth_mean) decode the full tree on first use."""
    __slots__ = ("_agent", "_key", "_dmax", "_dmean")

    def __init__(self, data, param, agent, key, na, total, dmax, dmean):
        self.data, self.param = data, param
        self.total, self.na = float(total), int(na)
        self._agent, self._key, self._dmax, self._dmean = agent, key, int(dmax), float(dmean)

    @property
    def children(self):
        return self._agent._full_root().actions[self._key].children

    def get_depth_max(self, depth):
        return depth + self._dmax                      # agents/abstract_mcts.py:190-200: levels below are relative

    def get_depth_mean(self, depth=0, root=False):
        if root:
            return depth + self._dmean if self._dmax > 0 else depth     # :202-209
        return self._agent._full_root().actions[self._key].get_depth_mean(depth)


class _RootProxy(AbstractStateNode):
    """Root state node: ``actions`` from the device's root arrays; ns / total decode the tree on first use."""
    __slots__ = ("_agent",)

    def __init__(self, data, param, agent, actions):
        self.data, self.param, self._agent = data, param, agent
        self.terminal, self.actions = False, actions
        self.terminal_reward, self.visit_actions = None, None

    @property
    def ns(self):
        return self._agent._full_root().ns

    @property
    def total(self):
        return self._agent._full_root().total


_DEPTH_MODE_WARNED = False


def _warn_depth_mode_once(cls_name):
    """Mcts / APW / SPW / DPW of the reference HEAD pass max_depth as the current depth and therefore play zero-length
    rollouts; this package plays them to max_depth unless told otherwise."""
    global _DEPTH_MODE_WARNED
    if not _DEPTH_MODE_WARNED:
        _DEPTH_MODE_WARNED = True
        import warnings
        warnings.warn(f"{cls_name}: depth_mode not given -> 'budgeted' (rollouts to max_depth).  The reference at HEAD plays "
                      f"zero-length rollouts in this class; pass depth_mode='reference_head' (the default of the islaMcts "
                      f"compat package) for HEAD-identical statistics.", stacklevel=4)


# Engines are POOLED by configuration.  The reference's episode loop builds a new agent for every real step
# (sweep.py:186, README.md:99-121); creating a device engine for each of them (workspace, host libm tables, the return
# table, their uploads) cost more than the search itself (round 1: 27 ms per fit() for a 10 ms kernel).  An agent takes
# a free engine of its configuration from the pool, re-keys it (new Philox seed) and resets the tree for its root; the
# engine goes back to the pool when the agent is dropped.
_ENGINE_POOL: dict = {}
_ENGINE_POOL_MAX = 4   # free engines kept per configuration


def _acquire_engine(env_id, capacity, kw):
    key = (env_id, capacity, tuple(sorted((k, str(v)) for k, v in kw.items())))
    free = _ENGINE_POOL.get(key)
    if free:
        return key, free.pop(), True
    return key, None, False


def _release_engine(key, eng):
    if eng is None or key is None:
        return
    free = _ENGINE_POOL.setdefault(key, [])
    if len(free) < _ENGINE_POOL_MAX:
        free.append(eng)
    else:
        eng.close()


def clear_engine_pool():
    """Frees the pooled device engines (workspaces go back to the CUDA allocator)."""
    for free in _ENGINE_POOL.values():
        for e in free:
            e.close()
    _ENGINE_POOL.clear()


class AbstractMcts:
    """Common driver.  Subclasses set AGENT (engine enum), HEAD_BUDGET (what the class does at reference HEAD)."""

    AGENT = _lib.AGENT_VANILLA
    HEAD_ZERO_ROLLOUTS = True      # Mcts/APW/SPW/DPW start the recursion at max_depth
    DISCRETE = True
    HASH_KEYS = True               # children keyed by observation.tobytes() (mcts_hash.py:97); Mcts keys by the observation

    def __init__(self, param: MctsParameters):
        self.param = param
        self.data: Any = None
        self.q_values = None
        self.trajectories_x: list = []
        self.trajectories_y: list = []
        self.param.depths = []
        self._engine: Engine | None = None
        self._engine_key = None
        self._pool_key = None
        self._root_proxy = None
        self._full_proxy = None
        self._root_stub = AbstractStateNode(param.root_data, param)
        self.root_statistics = None

    # ``agent.root.data = observation`` is assigned by callers before fit() (sweep.py:189)
    @property
    def root(self):
        if self._engine is None:
            return self._root_stub
        if self._root_proxy is None:
            self._root_proxy = self._build_root_proxy()
        return self._root_proxy

    @root.setter
    def root(self, value):
        self._root_stub = value

    # ------------------------------------------------------------------ engine plumbing
    def _engine_kwargs(self, env_id):
        p = self.param
        if _kind(p.action_selection_fn, "action_selection_fn") != "ucb1":
            raise NotImplementedError("action_selection_fn must be ucb1")
        rk = _kind(p.rollout_selection_fn, "rollout_selection_fn")
        if rk == "grid" and env_id != _lib.ENV_FROZENLAKE:
            raise NotImplementedError("grid_policy reads env.s: FrozenLake only (utils/action_selection_functions.py:271-272)")
        if rk not in ("random", "grid"):
            raise NotImplementedError("rollout_selection_fn must be continuous_default_policy (env.action_space.sample()) "
                                      "or grid_policy(...)")
        if p.depth_mode not in (None, "budgeted", "reference_head"):
            raise ValueError("depth_mode must be 'budgeted' or 'reference_head'")
        if p.depth_mode is None and self.HEAD_ZERO_ROLLOUTS:
            _warn_depth_mode_once(type(self).__name__)
        zero = p.depth_mode == "reference_head" and self.HEAD_ZERO_ROLLOUTS
        kw = dict(agent=self.AGENT, n_trees=int(p.n_trees), max_depth=int(p.max_depth), C_=float(p.C), gamma=float(p.gamma),
                  budget_mode=_lib.BUDGET_ZERO if zero else _lib.BUDGET_TO_MAX_DEPTH,
                  rollouts_per_leaf=int(p.rollouts_per_leaf), tree_parallel=int(p.tree_parallel), virtual_value=float(p.virtual_value),
                  precision=_lib.PRECISION_F64 if p.precision == "f64" else _lib.PRECISION_F32, device=p.device)
        n, grid = env_action_info(p.env, env_id)
        if grid:
            kw["action_grid"] = grid
        if self.DISCRETE:
            if n == 0:
                raise NotImplementedError(f"{type(self).__name__} needs a discrete action space")
            if p.n_actions is not None and int(p.n_actions) != n:
                raise ValueError(f"n_actions={p.n_actions} but the environment has {n} actions")
            kw["n_actions"] = n
        elif n != 0:
            raise NotImplementedError(f"{type(self).__name__} samples a continuous Box action space")
        self._agent_kwargs(kw)
        return kw

    def _agent_kwargs(self, kw):
        pass

    def _ensure_engine(self, env_id, root_state, noise=None):
        p = self.param
        kw = self._engine_kwargs(env_id)
        # a tree is continued by the next fit() only from the same root under the same transition noise (slippery
        # FrozenLake: the stored children were produced by the clone's draws as of the fit() that created them)
        key = (env_id, tuple(sorted((k, v) for k, v in kw.items())), tuple(root_state.tolist()),
               None if noise is None else noise.tobytes())
        if self._engine is not None and key == self._engine_key and self._sims_done + p.n_sim <= self._capacity:
            return                                     # fit() again continues the same tree
        if self._engine is not None and key == self._engine_key:
            raise _lib.IslaError(_lib.ERR_CAPACITY, "tree capacity exhausted; create a new agent")
        seed = p.seed if p.seed is not None else int(np.random.randint(0, 2**31 - 1))
        self._capacity = max(4 * int(p.n_sim), 64)
        self._drop_engine()
        self._pool_key, eng, pooled = _acquire_engine(env_id, self._capacity, kw)
        if pooled:
            eng.reseed(seed, 0)
        else:
            eng = Engine(env_id, sim_capacity=self._capacity, seed=seed, **kw)
        self._engine = eng
        if _kind(p.rollout_selection_fn, "rollout_selection_fn") == "grid":
            eng.set_rollout_heuristic(p.rollout_selection_fn.table)
        elif pooled and env_id == _lib.ENV_FROZENLAKE:
            eng.set_rollout_heuristic(None)
        eng.set_roots(np.tile(root_state[None, :], (int(p.n_trees), 1)))
        self._engine_key = key
        self._sims_done = 0

    def _drop_engine(self):
        eng, self._engine = self._engine, None
        if eng is not None:
            _release_engine(self._pool_key, eng)
        self._pool_key = None

    def __del__(self):
        try:
            self._drop_engine()
        except Exception:
            pass

    # ------------------------------------------------------------------ fit
    def fit(self):
        """Builds the tree(s) on the device and returns the best action (reference: mcts_hash.py:22-50)."""
        p = self.param
        if p.env is None:
            raise ValueError("param.env is None")
        env_id = identify_env(p.env)
        root_state = env_root_state(p.env, env_id)
        noise = None
        live_env = self.AGENT in (_lib.AGENT_SPW, _lib.AGENT_DPW)
        if env_id == _lib.ENV_FROZENLAKE:
            # slippery lake.  Mcts / MctsHash / apw step a pickled clone per simulation: the clone's per-step draws, as of
            # this fit().  spw / dpw step the LIVE env: the draws its generator will produce run on through the whole
            # fit() (one per env.step of the search), and the env's generator is advanced by what the search consumed.
            n_draws = max(int(p.max_depth), 4 * int(p.n_sim), 64) + 4
            if live_env:
                zero = p.depth_mode == "reference_head" and self.HEAD_ZERO_ROLLOUTS
                n_draws = min(int(p.n_sim) * ((0 if zero else int(p.max_depth)) + 64) + 64, 1 << 24)
            noise = lake_transition_noise(p.env, n_draws)
        self._ensure_engine(env_id, root_state, noise)
        eng = self._engine
        if env_id == _lib.ENV_FROZENLAKE:
            eng.set_transition_noise(noise)
        eng.run(int(p.n_sim))
        self._sims_done += int(p.n_sim)
        self._root_proxy = None
        self._full_proxy = None
        cnt = eng.counters()
        if cnt["faults"]:
            raise _lib.IslaError(_lib.ERR_CAPACITY, f"device search reported faults: {cnt}")
        if noise is not None and live_env:
            rng = getattr(getattr(p.env, "unwrapped", p.env), "np_random", None)
            if rng is not None and hasattr(rng, "random"):
                rng.random(int(cnt["env_steps"]))       # the search stepped the live env this many times
        return self._finish(eng, env_id)

    def _finish(self, eng, env_id):
        rank, act = eng.best()
        if int(self.param.n_trees) > 1:
            return self._finish_ensemble(eng)
        rank0, act0 = int(rank[0].item()), act[0].cpu().numpy().reshape(-1)
        self._fill_q_values(eng)
        return self._format_action(rank0, act0)

    def _fill_q_values(self, eng):
        if eng.layout.kernel == _lib.KERNEL_CTA_PER_TREE:
            a, na, tot = eng.root_children(0)
        else:
            a, na, tot = eng.root_edges(0)
        self.root_statistics = dict(action=a, na=na, total=tot)
        self.q_values = tot / na

    def _finish_ensemble(self, eng):
        """Root-parallel ensemble (n_trees > 1): root statistics summed over the trees (discrete agents), or the union of
        the trees' root children (continuous agents: every tree sampled its own actions)."""
        if not self.DISCRETE:
            from .. import sharding
            rows = [eng.root_children(i) for i in range(eng.n_trees)]
            a = torch.as_tensor(np.concatenate([np.asarray(r[0]).reshape(len(r[1]), -1) for r in rows]))
            na = torch.as_tensor(np.concatenate([r[1] for r in rows]))
            tot = torch.as_tensor(np.concatenate([r[2] for r in rows]))
            ma, mna, mtot = sharding.merge_root_children(a, na, tot)
            self.root_statistics = dict(action=ma.numpy(), na=mna.numpy(), total=mtot.numpy())
            self.q_values = self.root_statistics["total"] / self.root_statistics["na"]
            i, row = sharding.pick_continuous(ma, mna, mtot)
            return self._format_action(i, row.numpy())
        from ..engine import merge_root_stats
        na, tot = eng.root_stats()
        group = torch.zeros(eng.n_trees, dtype=torch.int32, device=na.device)   # every tree shares the one root
        mna, mtot = merge_root_stats(na, tot, group, 1)
        na0, tot0 = mna[0].cpu().numpy(), mtot[0].cpu().numpy()
        seen = na0 > 0
        q = np.where(seen, tot0 / np.maximum(na0, 1), -np.inf)
        self.root_statistics = dict(action=np.flatnonzero(seen).astype(np.float64), na=na0[seen], total=tot0[seen])
        self.q_values = q[seen]
        best = np.flatnonzero(q == q.max())
        return np.int64(np.random.choice(best))

    def _format_action(self, rank, action_value):
        return np.int64(int(action_value[0]))

    # ------------------------------------------------------------------ proxies
    def _make_state(self, data):
        return AbstractStateNode(data, self.param)

    def _action_key(self, a_row):
        """(data, dict key) of a root action as the reference stores it (int for discrete agents, ndarray.tobytes())."""
        a_row = np.atleast_1d(a_row)
        if self.DISCRETE:
            return int(a_row[0]), int(a_row[0])
        if a_row.size == 2:                            # CurveEnv: float64 (acceleration, steering)
            data = np.array(a_row, dtype=np.float64)
            return data, data.tobytes()
        space = getattr(self.param.env, "action_space", None)
        dtype = getattr(space, "dtype", np.float32)
        data = np.array([a_row[0]], dtype=np.float64 if a_row[0] != np.float32(a_row[0]) else dtype)
        return data, data.tobytes()

    def _build_root_proxy(self):
        eng = self._engine
        if eng.layout.kernel == _lib.KERNEL_CTA_PER_TREE:
            a, na, tot = eng.root_children(0)
        else:
            a, na, tot = eng.root_edges(0)
        dmax, dmean = eng.depth_stats(0)
        actions = {}
        for i in range(len(na)):
            data, key = self._action_key(a[i])
            actions[key] = _RootActionProxy(data, self.param, self, key, na[i], tot[i], dmax[i], dmean[i])
        return _RootProxy(self._root_stub.data, self.param, self, actions)

    def _full_root(self):
        if self._full_proxy is None:
            self._full_proxy = self._build_proxy()
        return self._full_proxy

    def _build_proxy(self):
        ser = self._engine.export_tree(0)
        kind, depth, count, total = ser["kind"], ser["depth"], ser["count"], ser["total"]
        term, action = ser["terminal"], ser["action"]
        action2 = ser.get("action2") if getattr(self._engine, "action_dim", 1) == 2 else None
        states = ser.get("state")
        env_id = self._engine.env_id
        root = self._make_state(self._root_stub.data)
        root.ns, root.total = int(count[0]), float(total[0])
        stack = [(root, 0)]              # (node object, depth) of the open state nodes
        action_at = {}                   # state depth -> the action node last emitted below a state of that depth
        space = getattr(self.param.env, "action_space", None)
        dtype = getattr(space, "dtype", np.float32)
        for i in range(1, len(kind)):
            if kind[i] == 1:             # action node: child of the open state node at this depth
                while stack[-1][1] > depth[i]:
                    stack.pop()
                parent = stack[-1][0]
                if self.DISCRETE:
                    data = int(action[i]); key = data
                elif action2 is not None:          # CurveEnv: float64 (acceleration, steering) keyed like action.tobytes()
                    data = np.array([action[i], action2[i]], dtype=np.float64)
                    key = data.tobytes()
                else:
                    data = np.array([action[i]], dtype=np.float64 if action[i] != np.float32(action[i]) else dtype)
                    key = data.tobytes()
                node = AbstractActionNode(data, self.param)
                node.na, node.total = int(count[i]), float(total[i])
                parent.actions[key] = node
                action_at[int(depth[i])] = node
            else:                        # state node: a child of the action node last emitted one level up (an action node
                                         # has several on a stochastic env), keyed like the reference keys its children
                                         # dict: by the observation (Mcts) or observation.tobytes() (the *Hash agents)
                obs = observation_of(env_id, states[i]) if states is not None else None
                st = self._make_state(obs)
                st.ns, st.total, st.terminal = int(count[i]), float(total[i]), bool(term[i])
                owner = action_at[int(depth[i]) - 1]
                if obs is None:
                    key = len(owner.children)
                elif self.HASH_KEYS and hasattr(obs, "tobytes"):
                    key = obs.tobytes()
                else:
                    key = obs if not isinstance(obs, np.ndarray) else obs.tobytes()
                owner.children[key] = st
                while stack[-1][1] >= depth[i]:
                    stack.pop()
                stack.append((st, int(depth[i])))
        return root

    def visualize(self, extension: str = "0") -> None:
        raise NotImplementedError("graphviz drawing is out of scope (SURVEY section 2, row 1)")
