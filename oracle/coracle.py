"""TEST INFRASTRUCTURE ONLY -- ctypes binding of oracle/c/libisla_oracle.so."""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "c", "libisla_oracle.so")

ENV_CARTPOLE, ENV_PENDULUM, ENV_MOUNTAINCAR, ENV_FROZENLAKE, ENV_CURVE = 0, 1, 2, 3, 4
AGENT_VANILLA, AGENT_APW, AGENT_SPW, AGENT_DPW = 0, 1, 2, 3
BUDGET_TO_MAX_DEPTH, BUDGET_ZERO = 0, 1
EXPAND_RANDOM, EXPAND_VOO, EXPAND_GENETIC2, EXPAND_GENETIC = 0, 1, 2, 3


class Config(C.Structure):
    _fields_ = [
        ("env_id", C.c_int32), ("agent", C.c_int32), ("n_actions", C.c_int32), ("max_depth", C.c_int32),
        ("budget_mode", C.c_int32), ("rollouts_per_leaf", C.c_int32), ("expansion", C.c_int32),
        ("voo_n_samples", C.c_int32), ("voo_max_try", C.c_int32), ("grid0", C.c_int32), ("grid1", C.c_int32),
        ("reserved", C.c_int32),
        ("C", C.c_double), ("gamma", C.c_double), ("alpha1", C.c_double), ("k1", C.c_double),
        ("alpha2", C.c_double), ("k2", C.c_double), ("epsilon", C.c_double),
        ("seed", C.c_uint64), ("tree_id", C.c_uint64),
    ]


def build(force: bool = False) -> str:
    src = os.path.join(_HERE, "c", "isla_oracle.c")
    hdr = os.path.join(_HERE, "c", "isla_oracle.h")
    stale = (not os.path.exists(_SO)) or any(
        os.path.exists(p) and os.path.getmtime(p) > os.path.getmtime(_SO) for p in (src, hdr))
    if force or stale:
        subprocess.run(["make", "-C", os.path.join(_HERE, "c"), "-B"], check=True, capture_output=True)
    return _SO


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(_SO)
        L.orc_tree_create.restype = C.c_void_p
        L.orc_tree_create.argtypes = [C.POINTER(Config), C.c_void_p]
        L.orc_tree_destroy.argtypes = [C.c_void_p]
        L.orc_tree_set_heuristic.restype = C.c_int
        L.orc_tree_set_heuristic.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int]
        L.orc_tree_set_noise.restype = C.c_int
        L.orc_tree_set_noise.argtypes = [C.c_void_p, C.c_void_p, C.c_int]
        L.orc_tree_run.restype = C.c_int
        L.orc_tree_run.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_int64]
        L.orc_tree_trace_pos.restype = C.c_int64
        L.orc_tree_trace_pos.argtypes = [C.c_void_p]
        L.orc_tree_best.restype = C.c_int
        L.orc_tree_best.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
        L.orc_tree_root_children.restype = C.c_int
        L.orc_tree_root_children.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
        L.orc_tree_depth_stats.restype = C.c_int
        L.orc_tree_depth_stats.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p]
        L.orc_tree_serialise.restype = C.c_int64
        L.orc_tree_serialise.argtypes = [C.c_void_p, C.c_int64] + [C.c_void_p] * 8
        L.orc_tree_counters.argtypes = [C.c_void_p, C.c_void_p]
        L.orc_run_many.restype = C.c_int
        L.orc_run_many.argtypes = [C.POINTER(Config), C.c_void_p, C.c_int64, C.c_int, C.c_int,
                                   C.c_void_p, C.c_void_p, C.c_void_p]
        L.orc_state_dim.restype = C.c_int
        L.orc_state_dim.argtypes = [C.c_int]
        L.orc_env_step.restype = C.c_int
        L.orc_env_step.argtypes = [C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
        L.orc_action_dim.restype = C.c_int
        L.orc_action_dim.argtypes = [C.c_int]
        L.orc_env_reset.argtypes = [C.c_int, C.c_uint64, C.c_uint64, C.c_void_p]
        L.orc_rollout.restype = C.c_int
        L.orc_rollout.argtypes = [C.c_int, C.c_void_p, C.c_int, C.c_double, C.c_uint64, C.c_uint64,
                                  C.c_uint32, C.c_uint32, C.c_void_p]
        L.orc_rollout_grid.restype = C.c_int
        L.orc_rollout_grid.argtypes = [C.c_int, C.c_void_p, C.c_int, C.c_double, C.c_uint64, C.c_uint64,
                                       C.c_uint32, C.c_uint32, C.c_int, C.c_int, C.c_void_p]
        L.orc_philox4x32.argtypes = [C.c_uint32] * 6 + [C.c_void_p]
        _lib = L
    return _lib


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


def make_config(env_id, agent=AGENT_VANILLA, n_actions=0, max_depth=500, budget_mode=BUDGET_TO_MAX_DEPTH,
                rollouts_per_leaf=1, expansion=EXPAND_RANDOM, voo_n_samples=5, voo_max_try=1000, C_=1.0,
                gamma=0.99, alpha1=0.0, k1=1.0, alpha2=0.0, k2=1.0, epsilon=0.7, seed=0, tree_id=0, grid0=0, grid1=0) -> Config:
    return Config(env_id, agent, n_actions, max_depth, budget_mode, rollouts_per_leaf, expansion,
                  voo_n_samples, voo_max_try, grid0, grid1, 0, C_, gamma, alpha1, k1, alpha2, k2, epsilon, seed, tree_id)


class Tree:
    def __init__(self, cfg: Config, root_state):
        self.cfg = cfg
        self._root = np.zeros(8, dtype=np.float64)
        rs = np.asarray(root_state, dtype=np.float64).reshape(-1)
        self._root[: rs.size] = rs
        self._h = lib().orc_tree_create(C.byref(cfg), _p(self._root))

    def close(self):
        if self._h:
            lib().orc_tree_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def set_heuristic(self, table):
        """grid_policy prior knowledge [16, 4] (lower = better) for FrozenLake rollouts; None = random rollouts."""
        if table is None:
            return lib().orc_tree_set_heuristic(self._h, None, 0, 0)
        tb = np.ascontiguousarray(table, dtype=np.float64)
        return lib().orc_tree_set_heuristic(self._h, _p(tb), tb.shape[0], tb.shape[1])

    def set_noise(self, u):
        """Per-step transition noise of a slippery FrozenLake fit (same for every simulation); None = deterministic."""
        if u is None:
            return lib().orc_tree_set_noise(self._h, None, 0)
        u = np.ascontiguousarray(u, dtype=np.float64)
        return lib().orc_tree_set_noise(self._h, _p(u), len(u))

    def run(self, n_sim, trace=None) -> int:
        if trace is None:
            return lib().orc_tree_run(self._h, int(n_sim), None, None, -1)
        kinds, vals = trace
        kinds = np.ascontiguousarray(kinds, dtype=np.uint8)
        vals = np.ascontiguousarray(vals, dtype=np.float64)
        self._keep = (kinds, vals)
        return lib().orc_tree_run(self._h, int(n_sim), _p(kinds), _p(vals), len(kinds))

    def trace_pos(self) -> int:
        return int(lib().orc_tree_trace_pos(self._h))

    def best(self):
        rank = C.c_int32(0)
        act = np.zeros(2, np.float64)
        rc = lib().orc_tree_best(self._h, C.byref(rank), _p(act))
        self.best_action2 = float(act[1])
        return rc, rank.value, float(act[0])

    def root_children(self):
        n = lib().orc_tree_root_children(self._h, 0, None, None, None, None)
        a = np.zeros(n, np.float64)
        a2 = np.zeros(n, np.float64)
        na = np.zeros(n, np.int64)
        tot = np.zeros(n, np.float64)
        lib().orc_tree_root_children(self._h, n, _p(a), _p(na), _p(tot), _p(a2))
        self.root_action2 = a2
        return a, na, tot

    def depth_stats(self):
        """(get_depth_max(0), get_depth_mean(0, True)) of every root action node, root.actions order."""
        n = lib().orc_tree_root_children(self._h, 0, None, None, None, None)
        dmax = np.zeros(n, np.int32)
        dmean = np.zeros(n, np.float64)
        lib().orc_tree_depth_stats(self._h, n, _p(dmax), _p(dmean))
        return dmax, dmean

    def counters(self):
        out = np.zeros(8, np.int64)
        lib().orc_tree_counters(self._h, _p(out))
        return dict(env_steps=int(out[0]), rollout_steps=int(out[1]), levels=int(out[2]), sims=int(out[3]),
                    state_nodes=int(out[4]), action_nodes=int(out[5]), expand_steps=int(out[6]))

    def serialise(self):
        c = self.counters()
        cap = c["state_nodes"] + c["action_nodes"] + 4
        kind = np.zeros(cap, np.uint8)
        depth = np.zeros(cap, np.int32)
        count = np.zeros(cap, np.int64)
        total = np.zeros(cap, np.float64)
        term = np.zeros(cap, np.uint8)
        fan = np.zeros(cap, np.int32)
        act = np.zeros(cap, np.float64)
        act2 = np.zeros(cap, np.float64)
        n = lib().orc_tree_serialise(self._h, cap, _p(kind), _p(depth), _p(count), _p(total), _p(term), _p(fan), _p(act), _p(act2))
        assert n <= cap
        return dict(kind=kind[:n], depth=depth[:n], count=count[:n], total=total[:n], terminal=term[:n],
                    fanout=fan[:n], action=act[:n], action2=act2[:n])


def run_many(cfg: Config, roots, n_sim, n_threads=1):
    roots = np.ascontiguousarray(roots, dtype=np.float64)
    n = roots.shape[0]
    A = cfg.n_actions
    na = np.zeros((n, A), np.int64)
    tot = np.zeros((n, A), np.float64)
    cnt = np.zeros(8, np.int64)
    lib().orc_run_many(C.byref(cfg), _p(roots), n, int(n_sim), int(n_threads), _p(na), _p(tot), _p(cnt))
    counters = dict(env_steps=int(cnt[0]), rollout_steps=int(cnt[1]), levels=int(cnt[2]), sims=int(cnt[3]),
                    state_nodes=int(cnt[4]), action_nodes=int(cnt[5]), expand_steps=int(cnt[6]))
    return na, tot, counters


def state_dim(env_id):
    return lib().orc_state_dim(env_id)


def env_step(env_id, state, action):
    s = np.zeros(8, np.float64)
    st = np.asarray(state, np.float64).reshape(-1)
    s[: st.size] = st
    a = np.zeros(2, np.float64)
    av = np.asarray(action, np.float64).reshape(-1)
    a[: av.size] = av
    ns = np.zeros(8, np.float64)
    r = C.c_double(0)
    t = C.c_int32(0)
    obs = np.zeros(8, np.float64)
    lib().orc_env_step(env_id, _p(s), _p(a), _p(ns), C.byref(r), C.byref(t), _p(obs))
    d = state_dim(env_id)
    return ns[:d].copy(), r.value, bool(t.value), obs


def env_reset(env_id, seed, index):
    s = np.zeros(8, np.float64)
    lib().orc_env_reset(env_id, seed, index, _p(s))
    return s[: state_dim(env_id)].copy()


def rollout(env_id, state, budget, gamma, seed, tree, sim, j, grid=(0, 0)):
    s = np.zeros(8, np.float64)
    st = np.asarray(state, np.float64).reshape(-1)
    s[: st.size] = st
    ret = C.c_double(0)
    steps = lib().orc_rollout_grid(env_id, _p(s), int(budget), float(gamma), seed, tree, sim, j, int(grid[0]), int(grid[1]),
                                   C.byref(ret))
    return ret.value, steps


def philox(c0, c1, c2, c3, k0, k1):
    out = np.zeros(4, np.uint32)
    lib().orc_philox4x32(c0, c1, c2, c3, k0, k1, _p(out))
    return out
