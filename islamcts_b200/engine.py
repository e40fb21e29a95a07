"""Host-side handle of the device MCTS engine (thin wrapper over the C ABI in include/isla_b200.h).

PyTorch is plumbing here: it owns the device workspace / IO tensors and the CUDA stream; every
computation is a hand-written sm_100a kernel inside libisla_b200.so.  No CPU fallback exists.
"""
from __future__ import annotations

import ctypes as C

import numpy as np
import torch

from . import _lib
from ._lib import IslaConfig, IslaLayout, check

ENV_NAMES = {
    "CartPole-v1": _lib.ENV_CARTPOLE, "CartPole-v0": _lib.ENV_CARTPOLE,
    "Pendulum-v1": _lib.ENV_PENDULUM,
    "MountainCarContinuous-v0": _lib.ENV_MOUNTAINCAR,
    "FrozenLake-v1": _lib.ENV_FROZENLAKE,
    "CurveEnv": _lib.ENV_CURVE, "DiscreteCurveEnv": _lib.ENV_CURVE,
}
STATE_DIM = {_lib.ENV_CARTPOLE: 4, _lib.ENV_PENDULUM: 2, _lib.ENV_MOUNTAINCAR: 2, _lib.ENV_FROZENLAKE: 1, _lib.ENV_CURVE: 5}
N_ACTIONS = {_lib.ENV_CARTPOLE: 2, _lib.ENV_PENDULUM: 0, _lib.ENV_MOUNTAINCAR: 0, _lib.ENV_FROZENLAKE: 4, _lib.ENV_CURVE: 0}
ACTION_DIM = {_lib.ENV_CURVE: 2}   # every other env: 1
ACTION_BOUNDS = {_lib.ENV_PENDULUM: (-2.0, 2.0), _lib.ENV_MOUNTAINCAR: (-1.0, 1.0)}

LINK_BLOCK_MASK, LINK_TERMINAL, LINK_REWARD_ONE, LINK_RANK_SHIFT = 0x0FFFFFFF, 1 << 28, 1 << 29, 30


def _block_dtype(A):
    # node block of the thread-per-tree engine (isla_internal.h): {double total[A]; int32 na[A]; uint32 link[A]}
    return np.dtype([("total", "<f8", (A,)), ("na", "<i4", (A,)), ("link", "<u4", (A,))])


def _require_cuda(device) -> torch.device:
    dev = torch.device(device)
    if dev.type != "cuda" or not torch.cuda.is_available():
        raise _lib.IslaError(_lib.ERR_CUDA, "a CUDA device is required: islamcts_b200 has no CPU fallback")
    return dev


def _stream_ptr(dev) -> int:
    return int(torch.cuda.current_stream(dev).cuda_stream)


class Engine:
    """A batch of independent search trees living in HBM.

    Parameters mirror ``isla_config``; see include/isla_b200.h for the reference lines each replaces.
    """

    def __init__(self, env_id: int, agent: int = _lib.AGENT_VANILLA, n_trees: int = 1, sim_capacity: int = 1000,
                 max_depth: int = 500, C_: float = 1.0, gamma: float = 0.99, n_actions: int | None = None,
                 budget_mode: int = _lib.BUDGET_TO_MAX_DEPTH, rollouts_per_leaf: int = 1,
                 expansion: int = _lib.EXPAND_RANDOM, epsilon: float = 0.7, voo_n_samples: int = 5,
                 voo_max_try: int = 1000, alpha1: float = 0.0, k1: float = 1.0, alpha2: float = 0.0, k2: float = 1.0,
                 precision: int = _lib.PRECISION_F32, kernel: int = _lib.KERNEL_AUTO, block_threads: int = 0,
                 seed: int = 0, tree_id_base: int = 0, device="cuda:0", tuning: int | None = None,
                 action_grid=None, cluster_ctas: int = 0, tree_parallel: int = 0, virtual_value: float = 0.0):
        import os
        if tuning is None:
            tuning = int(os.environ.get("ISLA_TUNING", "0"))
        self.device = _require_cuda(device)
        self.lib = _lib.load()
        g0, g1 = (int(action_grid[0]), int(action_grid[1])) if action_grid else (0, 0)   # DiscreteCurveEnv([g0, g1])
        if n_actions is None:
            n_actions = g0 * g1 if env_id == _lib.ENV_CURVE else N_ACTIONS[env_id]
        self.cfg = IslaConfig(env_id, agent, int(n_actions or 0), int(max_depth), budget_mode, int(rollouts_per_leaf),
                              expansion, int(voo_n_samples), int(voo_max_try), precision, kernel, int(block_threads),
                              int(sim_capacity), int(tuning), g0, g1, int(cluster_ctas), int(tree_parallel), float(C_), float(gamma), float(alpha1), float(k1), float(alpha2),
                              float(k2), float(epsilon), int(seed), int(tree_id_base), int(n_trees), float(virtual_value))
        nbytes = check(self.lib.isla_engine_workspace_bytes(C.byref(self.cfg)))
        with torch.cuda.device(self.device):
            self.workspace = torch.empty(nbytes + 256, dtype=torch.uint8, device=self.device)
            base = self.workspace.data_ptr()
            self._ws_off = (-base) % 256
            handle = C.c_void_p()
            check(self.lib.isla_engine_create(C.byref(self.cfg), C.c_void_p(base + self._ws_off), nbytes, C.byref(handle)))
        self._h = handle
        self.layout = IslaLayout()
        check(self.lib.isla_engine_layout(self._h, C.byref(self.layout)))
        self.n_trees = int(n_trees)
        self.state_dim = STATE_DIM[env_id]
        self.n_actions = int(n_actions or 0)
        self.env_id = env_id
        self.action_dim = int(self.layout.action_dim)
        self._keep = []

    # ------------------------------------------------------------------ lifecycle
    def close(self):
        if getattr(self, "_h", None):
            self.lib.isla_engine_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ------------------------------------------------------------------ the fit() loop
    def set_roots(self, roots):
        """roots: [n_trees, state_dim] float64 (numpy / torch, host or device)."""
        t = torch.as_tensor(roots, dtype=torch.float64).reshape(self.n_trees, self.state_dim)
        t = t.to(self.device, non_blocking=True).contiguous()
        with torch.cuda.device(self.device):
            check(self.lib.isla_engine_set_roots(self._h, C.c_void_p(t.data_ptr()), C.c_void_p(_stream_ptr(self.device))))
        self._roots = t
        return self

    def set_rollout_heuristic(self, table):
        """grid_policy prior knowledge [16, 4] (lower = better) for FrozenLake rollouts; None = random rollouts."""
        if table is None:
            check(self.lib.isla_engine_set_rollout_heuristic(self._h, None, 0, 0))
            return self
        tb = np.ascontiguousarray(table, dtype=np.float64)
        with torch.cuda.device(self.device):
            check(self.lib.isla_engine_set_rollout_heuristic(self._h, tb.ctypes.data_as(C.c_void_p), tb.shape[0], tb.shape[1]))
        return self

    def set_transition_noise(self, u):
        """Slippery FrozenLake: u[k] = what the k-th env.step of every simulation of this fit draws (the reference clones
        the env's generator per simulation, utils/mcts_utils.py:8-16); None = deterministic env."""
        if u is None:
            check(self.lib.isla_engine_set_transition_noise(self._h, None, 0))
            return self
        u = np.ascontiguousarray(u, dtype=np.float64)
        with torch.cuda.device(self.device):
            check(self.lib.isla_engine_set_transition_noise(self._h, u.ctypes.data_as(C.c_void_p), len(u)))
        return self

    def reseed(self, seed: int, tree_id_base: int = 0):
        """New Philox key for the searches after the next set_roots() (a fresh tree per real step, sweep.py:186)."""
        check(self.lib.isla_engine_reseed(self._h, int(seed), int(tree_id_base)))
        return self

    def run(self, n_sim: int):
        with torch.cuda.device(self.device):
            check(self.lib.isla_engine_run(self._h, int(n_sim), C.c_void_p(_stream_ptr(self.device))))
        return self

    def run_scripted(self, tree: int, n_sim: int, kinds, values):
        k = torch.as_tensor(np.ascontiguousarray(kinds, dtype=np.uint8)).to(self.device)
        v = torch.as_tensor(np.ascontiguousarray(values, dtype=np.float64)).to(self.device)
        self._keep = [k, v]
        with torch.cuda.device(self.device):
            check(self.lib.isla_engine_run_scripted(self._h, int(tree), int(n_sim), C.c_void_p(k.data_ptr()),
                                                    C.c_void_p(v.data_ptr()), int(k.numel()),
                                                    C.c_void_p(_stream_ptr(self.device))))
        return self

    def root_stats(self):
        """(na int64 [n_trees, A], total float64 [n_trees, A]) device tensors, by ACTION index."""
        na = torch.empty((self.n_trees, self.n_actions), dtype=torch.int64, device=self.device)
        tot = torch.empty((self.n_trees, self.n_actions), dtype=torch.float64, device=self.device)
        with torch.cuda.device(self.device):
            check(self.lib.isla_engine_root_stats(self._h, C.c_void_p(na.data_ptr()), C.c_void_p(tot.data_ptr()),
                                                  C.c_void_p(_stream_ptr(self.device))))
        return na, tot

    def root_merge(self, n_roots: int, out: torch.Tensor | None = None):
        """root_stats + per-root sum in one launch (thread-per-tree engine; tree i belongs to root i // (n_trees // n_roots)):
        packed float64 [n_roots, A, 2] = (na, total), ready for ONE all-reduce."""
        if out is None:
            out = torch.empty((int(n_roots), self.n_actions, 2), dtype=torch.float64, device=self.device)
        with torch.cuda.device(self.device):
            check(self.lib.isla_engine_root_merge(self._h, int(n_roots), C.c_void_p(out.data_ptr()),
                                                  C.c_void_p(_stream_ptr(self.device))))
        return out

    def reroot(self, actions, n_sim_next: int, alive=None):
        """Tree reuse (build extension): the subtree below root action actions[tree] becomes the tree (thread-per-tree
        engine); the next run(n_sim_next) continues the search from the chosen child; a subtree too large for the workspace
        is dropped (fresh root).  actions: int32 [n_trees] device tensor."""
        a = torch.as_tensor(actions).to(self.device, torch.int32).contiguous()
        al = None if alive is None else torch.as_tensor(alive).to(self.device, torch.uint8).contiguous()
        self._keep = [a, al]
        with torch.cuda.device(self.device):
            check(self.lib.isla_engine_reroot(self._h, C.c_void_p(a.data_ptr()), C.c_void_p(al.data_ptr()) if al is not None else None,
                                              int(n_sim_next), C.c_void_p(_stream_ptr(self.device))))
        return self

    def best(self, scripted_pos: int = -1):
        """End of fit(): (rank in root.actions, action value) per tree -- device tensors (int32, float64;
        the action is [n_trees, 2] for the continuous CurveEnv box)."""
        rank = torch.empty((self.n_trees,), dtype=torch.int32, device=self.device)
        shape = (self.n_trees,) if self.action_dim == 1 else (self.n_trees, self.action_dim)
        act = torch.empty(shape, dtype=torch.float64, device=self.device)
        with torch.cuda.device(self.device):
            check(self.lib.isla_engine_best(self._h, int(scripted_pos), C.c_void_p(rank.data_ptr()),
                                            C.c_void_p(act.data_ptr()), C.c_void_p(_stream_ptr(self.device))))
        return rank, act

    def root_children(self, tree: int = 0):
        """(action, na, total) numpy arrays of one tree's root children in root.actions order (CTA engine)."""
        cnt = torch.zeros(1, dtype=torch.int32, device=self.device)
        cap = max(16, int(self.layout.a_cap))
        a = torch.empty(cap if self.action_dim == 1 else (cap, self.action_dim), dtype=torch.float64, device=self.device)
        na = torch.empty(cap, dtype=torch.int64, device=self.device)
        tot = torch.empty(cap, dtype=torch.float64, device=self.device)
        with torch.cuda.device(self.device):
            check(self.lib.isla_engine_root_children(self._h, int(tree), cap, C.c_void_p(a.data_ptr()),
                                                     C.c_void_p(na.data_ptr()), C.c_void_p(tot.data_ptr()),
                                                     C.c_void_p(cnt.data_ptr()), C.c_void_p(_stream_ptr(self.device))))
        n = int(cnt.item())
        return a[:n].cpu().numpy(), na[:n].cpu().numpy(), tot[:n].cpu().numpy()

    def depth_stats(self, tree: int = 0):
        """(get_depth_max(0) int32, get_depth_mean(0, True) float64) of one tree's root action nodes, root.actions order
        (reference: agents/abstract_mcts.py:190-209; the metrics sweep.py:199-200 logs every real step)."""
        cnt = torch.zeros(1, dtype=torch.int32, device=self.device)
        cap = max(16, int(self.layout.a_cap), self.n_actions)
        dmax = torch.empty(cap, dtype=torch.int32, device=self.device)
        dmean = torch.empty(cap, dtype=torch.float64, device=self.device)
        with torch.cuda.device(self.device):
            check(self.lib.isla_engine_depth_stats(self._h, int(tree), cap, C.c_void_p(dmax.data_ptr()),
                                                   C.c_void_p(dmean.data_ptr()), C.c_void_p(cnt.data_ptr()),
                                                   C.c_void_p(_stream_ptr(self.device))))
        n = int(cnt.item())
        return dmax[:n].cpu().numpy(), dmean[:n].cpu().numpy()

    def counters(self) -> dict:
        out = (C.c_int64 * 8)()
        with torch.cuda.device(self.device):
            check(self.lib.isla_engine_counters(self._h, out, C.c_void_p(_stream_ptr(self.device))))
        return dict(env_steps=out[0], rollout_steps=out[1], levels=out[2], sims=out[3], nodes=out[4], faults=out[5])

    # ------------------------------------------------------------------ host views of the device tree
    def _ws_bytes(self, offset: int, nbytes: int) -> np.ndarray:
        lo = self._ws_off + int(offset)
        return self.workspace[lo: lo + int(nbytes)].cpu().numpy()

    def tree_meta(self, tree: int = 0) -> dict:
        if self.layout.kernel == _lib.KERNEL_CTA_PER_TREE:
            off = self.pool_offsets()
            m = self._ws_bytes(self.layout.pool_offset + tree * self.layout.pool_stride + off[18], 32).view(np.int32)
            return dict(state_nodes=int(m[0]), action_nodes=int(m[1]), sims=int(m[3]), rng_words=int(m[4]),
                        fault=int(m[5]), trace_pos=int(m[6]))
        m = self._ws_bytes(self.layout.meta_offset + tree * 32, 32).view(np.int32)
        return dict(alloc=int(m[0]), sims=int(m[1]), rng_words=int(m[2]), fault=int(m[3]), trace_pos=int(m[4]))

    def pool_offsets(self):
        out = (C.c_int64 * 24)()
        check(self.lib.isla_engine_pool_offsets(self._h, out))
        return [int(x) for x in out]

    def raw_blocks(self, tree: int = 0) -> np.ndarray:
        L = self.layout
        return self._ws_bytes(L.rec_offset + tree * L.rec_stride, L.rec_stride).view(_block_dtype(self.n_actions))

    def raw_gfirst(self, tree: int = 0) -> np.ndarray:
        L = self.layout
        return self._ws_bytes(L.gfirst_offset + tree * L.gfirst_stride, L.gfirst_stride).view(np.float64)

    def root_edges(self, tree: int = 0):
        """(action, na, total) numpy arrays of a thread-per-tree tree's visited root edges in root.actions (insertion)
        order: one small device->host copy of the root's node block."""
        L = self.layout
        A = self.n_actions
        b = self._ws_bytes(L.rec_offset + tree * L.rec_stride + A * 16, A * 16).view(_block_dtype(A))[0]
        order = sorted((int(b["link"][k] >> LINK_RANK_SHIFT), k) for k in range(A) if b["na"][k] > 0)
        ks = [k for _, k in order]
        return (np.array(ks, dtype=np.float64), np.array([int(b["na"][k]) for k in ks], dtype=np.int64),
                np.array([float(b["total"][k]) for k in ks], dtype=np.float64))

    def root_order(self, tree: int = 0):
        """Visited root actions of a thread-per-tree tree in insertion order (root.actions order)."""
        b = self.raw_blocks(tree)[1]
        order = sorted((int(b["link"][k] >> LINK_RANK_SHIFT), k) for k in range(self.n_actions) if b["na"][k] > 0)
        return [k for _, k in order]

    def raw_states(self, tree: int = 0) -> np.ndarray:
        L = self.layout
        dt = np.float32 if L.real_bytes == 4 else np.float64
        return self._ws_bytes(L.state_offset + tree * L.state_stride, L.state_stride).view(dt).reshape(-1, L.state_dim)

    def export_tree(self, tree: int = 0) -> dict:
        """Canonical DFS serialisation (same format as oracle/trace.py) decoded on the host."""
        if self.layout.kernel == _lib.KERNEL_THREAD_PER_TREE:
            return _export_tpt(self.raw_blocks(tree), self.raw_gfirst(tree), self.n_actions, self.raw_states(tree))
        from .tree_export import export_pool  # CTA-per-tree general pools
        return export_pool(self, tree)


def _export_tpt(blocks: np.ndarray, gfirst: np.ndarray, A: int, states: np.ndarray | None = None) -> dict:
    kinds, depths, counts, totals, terms, fans, acts, slots = [], [], [], [], [], [], [], []

    def children(block):
        """visited edges of a block in insertion (rank) order -> list of (action, na, total, link)"""
        if block == 0:
            return []
        b = blocks[block]
        out = [(int(b["link"][a] >> LINK_RANK_SHIFT), a) for a in range(A) if b["na"][a] > 0]
        out.sort()
        return [(a, int(b["na"][a]), float(b["total"][a]), int(b["link"][a]), block * A + a) for _, a in out]

    # stack entries: ("S", edge tuple or None for the root, depth) / ("A", edge tuple, depth)
    stack = [("S", None, 0)]
    while stack:
        kind, edge, d = stack.pop()
        if kind == "S":
            if edge is None:
                ch = children(1)
                ns, tot, term = sum(c[1] for c in ch), float(sum(c[2] for c in ch)), 0
            else:
                _, na, _, link, slot = edge
                term = 1 if link & LINK_TERMINAL else 0
                if term:
                    ch, ns, tot = [], na, 0.0
                else:
                    ch = children(link & LINK_BLOCK_MASK)
                    ns = 1 + sum(c[1] for c in ch)
                    tot = float(gfirst[slot]) + float(sum(c[2] for c in ch))
            kinds.append(0); depths.append(d); counts.append(ns); totals.append(tot); terms.append(term)
            fans.append(len(ch)); acts.append(0.0); slots.append(0 if edge is None else edge[4])
            for c in reversed(ch):
                stack.append(("A", c, d))
        else:
            a, na, total, link, slot = edge
            kinds.append(1); depths.append(d); counts.append(na); totals.append(total)
            terms.append(0); fans.append(1); acts.append(float(a)); slots.append(-1)
            stack.append(("S", edge, d + 1))
    out = dict(kind=np.asarray(kinds, np.uint8), depth=np.asarray(depths, np.int32), count=np.asarray(counts, np.int64),
               total=np.asarray(totals, np.float64), terminal=np.asarray(terms, np.uint8),
               fanout=np.asarray(fans, np.int32), action=np.asarray(acts, np.float64))
    if states is not None:
        # env state of every state record (the slot of the edge that leads to it; the root's is slot 0); action records: 0
        sl = np.asarray(slots, np.int64)
        st = np.zeros((len(sl), states.shape[1]), dtype=np.float64)
        st[sl >= 0] = states[sl[sl >= 0]]
        out["state"] = st
    return out


# ---------------------------------------------------------------------- stateless batched primitives
def _real(precision):
    return torch.float64 if precision == _lib.PRECISION_F64 else torch.float32


def env_step(env_id: int, states, actions, precision: int = _lib.PRECISION_F32, device="cuda:0"):
    """gym ``env.step`` over a batch on the device.  Returns (next_states, rewards, terminated)."""
    dev = _require_cuda(device)
    lib = _lib.load()
    dt = _real(precision)
    S = STATE_DIM[env_id]
    s = torch.as_tensor(states).to(dev, dt).reshape(-1, S).contiguous()
    a = torch.as_tensor(actions).to(dev, dt).reshape(-1).contiguous()
    n = s.shape[0]
    if a.numel() != n * ACTION_DIM.get(env_id, 1):
        raise _lib.IslaError(_lib.ERR_INVALID, f"env_step: {n} states need {n * ACTION_DIM.get(env_id, 1)} action reals, got {a.numel()}")
    ns = torch.empty_like(s)
    r = torch.empty(n, dtype=dt, device=dev)
    t = torch.empty(n, dtype=torch.uint8, device=dev)
    with torch.cuda.device(dev):
        check(lib.isla_env_step(env_id, precision, C.c_void_p(s.data_ptr()), C.c_void_p(a.data_ptr()), n,
                                C.c_void_p(ns.data_ptr()), C.c_void_p(r.data_ptr()), C.c_void_p(t.data_ptr()),
                                C.c_void_p(_stream_ptr(dev))))
    return ns, r, t


def rollout(env_id: int, states, n: int | None = None, budget: int = 500, gamma: float = 0.99, seed: int = 0,
            sim: int = 0, precision: int = _lib.PRECISION_F32, device="cuda:0", action_grid=None):
    """Random-policy rollouts (abstract_mcts.py:72-91).  states=None starts from the reset distribution.
    action_grid=[g0, g1] samples DiscreteCurveEnv's grid instead of CurveEnv's box."""
    g0, g1 = (int(action_grid[0]), int(action_grid[1])) if action_grid else (0, 0)
    dev = _require_cuda(device)
    lib = _lib.load()
    dt = _real(precision)
    if states is not None:
        s = torch.as_tensor(states).to(dev, dt).reshape(-1, STATE_DIM[env_id]).contiguous()
        n = s.shape[0]
        sp = C.c_void_p(s.data_ptr())
    else:
        sp = C.c_void_p(0)
    ret = torch.empty(n, dtype=torch.float64, device=dev)
    steps = torch.empty(n, dtype=torch.int32, device=dev)
    with torch.cuda.device(dev):
        check(lib.isla_rollout(env_id, precision, sp, int(n), int(budget), float(gamma), int(seed), int(sim), g0, g1,
                               C.c_void_p(ret.data_ptr()), C.c_void_p(steps.data_ptr()), C.c_void_p(_stream_ptr(dev))))
    return ret, steps


def kernel_launches() -> int:
    """Kernels libisla_b200 has launched in this process."""
    return int(_lib.load().isla_kernel_launches())


def merge_root_stats(na, total, group, n_roots: int):
    """Per-root sums over trees sharing a root (the local half of the root-parallel merge), in a fixed order.
    `group` must be ascending (the trees of a root contiguous), as sharding.plan() lays them out."""
    lib = _lib.load()
    dev = na.device
    n_trees, A = na.shape
    na_out = torch.empty((n_roots, A), dtype=torch.int64, device=dev)
    tot_out = torch.empty((n_roots, A), dtype=torch.float64, device=dev)
    gp = C.c_void_p(group.data_ptr()) if group is not None else C.c_void_p(0)
    with torch.cuda.device(dev):
        check(lib.isla_merge_root_stats(C.c_void_p(na.data_ptr()), C.c_void_p(total.data_ptr()), gp, n_trees, A,
                                        int(n_roots), C.c_void_p(na_out.data_ptr()), C.c_void_p(tot_out.data_ptr()),
                                        C.c_void_p(_stream_ptr(dev))))
    return na_out, tot_out
