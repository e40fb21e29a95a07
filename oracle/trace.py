"""TEST INFRASTRUCTURE ONLY -- decision-trace recorder and tree serialiser.

A *decision trace* is the ordered list of every random draw one ``fit()`` of the
reference makes.  Feeding the same trace to the C oracle and to the CUDA engine
("scripted mode") must reproduce the reference's tree exactly: integer visit
counts bit-exact, float totals to a stated tolerance.

Random call sites of the reference that are recorded (all under /root/reference):
  K_CHOICE   np.random.choice(candidates)         islaMcts/agents/mcts.py:57 (unvisited pick),
                                                  utils/action_selection_functions.py:24 (UCB tie-break),
                                                  agents/mcts.py:37 (final arg-max tie-break)
             value = POSITION of the pick inside the candidate list.
  K_SAMPLE_D action_space.sample() on Discrete    utils/action_selection_functions.py:47 via abstract_mcts.py:84
             value = the action.
  K_SAMPLE_C action_space.sample() on Box         same; also mcts_double_...:57.  One entry per action dim.
  K_CHOICES  random.choices(pop, weights)         mcts_state_...:116, mcts_double_...:126.  value = index in pop.
  K_U01      np.random.random()                   action_selection_functions.py:199 (voo epsilon test)
  K_UNIFORM  np.random.uniform(low, high, size)   action_selection_functions.py:162 (voo sample), one entry per dim.
"""
from __future__ import annotations

import contextlib
import random as _py_random

import numpy as np

K_CHOICE = 0
K_SAMPLE_D = 1
K_SAMPLE_C = 2
K_CHOICES = 3
K_U01 = 4
K_UNIFORM = 5


class Trace:
    def __init__(self):
        self.kinds: list[int] = []
        self.values: list[float] = []

    def add(self, kind: int, value: float):
        self.kinds.append(kind)
        self.values.append(float(value))

    def arrays(self):
        return np.asarray(self.kinds, dtype=np.uint8), np.asarray(self.values, dtype=np.float64)

    def __len__(self):
        return len(self.kinds)


class RecordingSpace:
    """Wraps an action space; ``sample()`` is logged.  Pickles WITHOUT the trace:
    my_deepcopy (/root/reference/islaMcts/utils/mcts_utils.py:15) re-attaches the
    live space after every clone, so the pickled copy is never sampled."""

    def __init__(self, space, trace: Trace):
        self._space = space
        self._trace = trace

    def __getattr__(self, name):
        if name.startswith("_"):
            raise AttributeError(name)
        return getattr(self._space, name)

    def __getstate__(self):
        return {"_space": None, "_trace": None}

    def __setstate__(self, st):
        self.__dict__.update(st)

    def sample(self):
        a = self._space.sample()
        if isinstance(a, np.ndarray):
            for v in a.reshape(-1):
                self._trace.add(K_SAMPLE_C, float(v))
        else:
            self._trace.add(K_SAMPLE_D, int(a))
        return a


@contextlib.contextmanager
def recording(trace: Trace, seed: int):
    """Replace numpy's / random's module-level draws by seeded, logged equivalents."""
    rng = np.random.default_rng(seed)
    orig_choice = np.random.choice
    orig_random = np.random.random
    orig_uniform = np.random.uniform
    orig_choices = _py_random.choices

    def choice(a, *args, **kwargs):
        assert not args and not kwargs, "reference only calls np.random.choice(a)"
        n = len(a)
        pos = int(rng.integers(n))
        trace.add(K_CHOICE, pos)
        return a[pos]

    def random_():
        v = float(rng.random())
        trace.add(K_U01, v)
        return v

    def uniform(low=0.0, high=1.0, size=None):
        # legacy np.random.uniform = low + (high - low) * random_sample(size); it accepts high < low (genetic_policy
        # passes the two lowest-q actions in either order), which Generator.uniform rejects
        lo, hi = np.asarray(low, dtype=np.float64), np.asarray(high, dtype=np.float64)
        v = lo + (hi - lo) * rng.random(size)
        for x in np.asarray(v, dtype=np.float64).reshape(-1):
            trace.add(K_UNIFORM, float(x))
        return v

    def choices(population, weights=None, **kwargs):
        pop = list(population)
        w = np.ones(len(pop)) if weights is None else np.asarray(weights, dtype=np.float64)
        cum = np.cumsum(w)
        u = float(rng.random()) * cum[-1]
        idx = int(np.searchsorted(cum, u, side="right"))
        idx = min(idx, len(pop) - 1)
        trace.add(K_CHOICES, idx)
        return [pop[idx]]

    np.random.choice = choice
    np.random.random = random_
    np.random.uniform = uniform
    _py_random.choices = choices
    try:
        yield trace
    finally:
        np.random.choice = orig_choice
        np.random.random = orig_random
        np.random.uniform = orig_uniform
        _py_random.choices = orig_choices


# ---------------------------------------------------------------------------------
# Canonical tree serialisation (shared by reference objects, C oracle and CUDA export)
# ---------------------------------------------------------------------------------
# DFS pre-order, children in INSERTION order (python dict order in the reference):
#   state record : (kind=0, depth, ns,  total, terminal, n_actions,  0.0)
#   action record: (kind=1, depth, na,  total, 0,        n_children, action_value[, action_value2])
# ``depth`` counts state levels (root = 0; an action node has its parent's depth).


def serialise_reference_tree(root, action_to_float=None):
    kinds, depths, counts, totals, flags, fanout, avals, avals2 = [], [], [], [], [], [], [], []
    if action_to_float is None:
        def action_to_float(a):
            return float(np.asarray(a, dtype=np.float64).reshape(-1)[0])

    stack = [(0, root, 0)]
    while stack:
        kind, node, depth = stack.pop()
        if kind == 0:
            kinds.append(0)
            depths.append(depth)
            counts.append(int(node.ns))
            totals.append(float(node.total))
            flags.append(int(bool(node.terminal)))
            fanout.append(len(node.actions))
            avals.append(0.0)
            avals2.append(0.0)
            for act in reversed(list(node.actions.values())):
                stack.append((1, act, depth))
        else:
            kinds.append(1)
            depths.append(depth)
            counts.append(int(node.na))
            totals.append(float(node.total))
            flags.append(0)
            fanout.append(len(node.children))
            avals.append(action_to_float(node.data))
            flat = np.asarray(node.data, dtype=np.float64).reshape(-1)
            avals2.append(float(flat[1]) if flat.size > 1 else 0.0)
            for st in reversed(list(node.children.values())):
                stack.append((0, st, depth + 1))
    return {
        "kind": np.asarray(kinds, dtype=np.uint8),
        "depth": np.asarray(depths, dtype=np.int32),
        "count": np.asarray(counts, dtype=np.int64),
        "total": np.asarray(totals, dtype=np.float64),
        "terminal": np.asarray(flags, dtype=np.uint8),
        "fanout": np.asarray(fanout, dtype=np.int32),
        "action": np.asarray(avals, dtype=np.float64),
        "action2": np.asarray(avals2, dtype=np.float64),   # second action dimension (CurveEnv), else 0
    }
