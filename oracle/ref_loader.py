"""TEST INFRASTRUCTURE ONLY -- import the reference's own agents as a live oracle.

Works only where ``/root/reference`` exists (the build container; NOT the GPU
box).  The reference imports ``graphviz``, ``gym`` and ``gymnasium`` at module
scope (/root/reference/islaMcts/agents/abstract_mcts.py:5,
agents/parameters/mcts_parameters.py:5, utils/mcts_utils.py:5) although the
search arithmetic needs none of them; stub modules are injected so the agent
modules import and run UNMODIFIED (recipe: SURVEY.md appendix A).

Shims that live outside reference code:
  * ``SpwParams`` / ``DpwParams``: params subclasses that add the
    ``state_variable`` slot the SPW/DPW agents read
    (mcts_state_progressive_widening_hash.py:24, mcts_double_...:25) but the
    dataclass no longer declares (mcts_parameters.py:23-24).
  * ``ucb1_index``: index-returning UCB for DPW (mcts_double_...:63-64 indexes
    a list with the result, while ucb1 returns the action data).
"""
from __future__ import annotations

import os
import sys
import types

# /root/reference in the build container; on the GPU box the copy staged by oracle/ref_bench.py (baseline/_ref, git-ignored)
_STAGED = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "baseline", "_ref")
REFERENCE_ROOT = os.environ.get("ISLA_REFERENCE_ROOT") or ("/root/reference" if os.path.isdir("/root/reference/islaMcts") else _STAGED)


def available() -> bool:
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "islaMcts", "agents"))


def _mod(name, **attrs):
    m = types.ModuleType(name)
    m.__dict__.update(attrs)
    sys.modules[name] = m
    return m


_loaded = None


def load():
    """Returns a namespace with the reference classes/functions (cached)."""
    global _loaded
    if _loaded is not None:
        return _loaded
    if not available():
        raise RuntimeError(f"reference tree not present at {REFERENCE_ROOT}")
    # The product ships a compat package also called ``islaMcts``; make sure the
    # REAL reference wins inside this process.
    for k in [k for k in sys.modules if k == "islaMcts" or k.startswith("islaMcts.")]:
        del sys.modules[k]
    if "graphviz" not in sys.modules:
        _mod("graphviz", Digraph=type("Digraph", (), {"__init__": lambda self, *a, **k: None}))
    if "gym" not in sys.modules:
        g = _mod("gym", Env=type("Env", (), {}))
        g.spaces = _mod("gym.spaces", Discrete=type("Discrete", (), {}), Box=type("Box", (), {}))
    if "gymnasium" not in sys.modules:
        from oracle import pyenvs
        gy = _mod("gymnasium", Env=type("Env", (), {"unwrapped": property(lambda self: self)}))
        gy.spaces = _mod("gymnasium.spaces", Discrete=pyenvs.Discrete, Box=pyenvs.Box)
        gy.envs = _mod("gymnasium.envs")
        gy.envs.registration = _mod("gymnasium.envs.registration", register=lambda **kw: None)
    sys.path.insert(0, REFERENCE_ROOT)
    try:
        from islaMcts.agents.mcts import Mcts
        from islaMcts.agents.mcts_hash import MctsHash
        from islaMcts.agents.mcts_action_progressive_widening_hash import MctsActionProgressiveWideningHash
        from islaMcts.agents.mcts_state_progressive_widening_hash import MctsStateProgressiveWideningHash
        from islaMcts.agents.mcts_double_progressive_widening_hash import MctsDoubleProgressiveWideningHash
        from islaMcts.agents.parameters.mcts_parameters import MctsParameters
        from islaMcts.agents.parameters.pw_parameters import PwParameters
        from islaMcts.agents.parameters.dpw_parameters import DpwParameters
        import islaMcts.utils.action_selection_functions as asf
        import islaMcts.agents.abstract_mcts as abstract_mcts
        from islaMcts.environments.curve_env import CurveEnv, DiscreteCurveEnv
    finally:
        sys.path.remove(REFERENCE_ROOT)
    import islaMcts as _pkg

    assert os.path.realpath(_pkg.__file__).startswith(os.path.realpath(REFERENCE_ROOT)), _pkg.__file__

    import numpy as np
    from dataclasses import dataclass

    @dataclass(slots=True)
    class SpwParams(PwParameters):
        state_variable: str = "state"

    @dataclass(slots=True)
    class DpwParams(DpwParameters):
        state_variable: str = "state"

    def ucb1_index(node):
        # same arithmetic as asf.ucb1 (action_selection_functions.py:9-25) but returns
        # the child INDEX, which is what mcts_double_...:63-64 expects.
        children = list(node.actions.values())
        na = np.array([c.na for c in children])
        q = np.array([c.total / c.na for c in children])
        score = q + node.param.C * np.sqrt(np.log(node.ns) / na)
        return np.random.choice(np.flatnonzero(score == score.max()))

    ns = types.SimpleNamespace(
        Mcts=Mcts,
        MctsHash=MctsHash,
        MctsApw=MctsActionProgressiveWideningHash,
        MctsSpw=MctsStateProgressiveWideningHash,
        MctsDpw=MctsDoubleProgressiveWideningHash,
        MctsParameters=MctsParameters,
        PwParameters=PwParameters,
        DpwParameters=DpwParameters,
        SpwParams=SpwParams,
        DpwParams=DpwParams,
        ucb1=asf.ucb1,
        ucb1_index=ucb1_index,
        continuous_default_policy=asf.continuous_default_policy,
        voo=asf.voo,
        asf=asf,
        abstract_mcts=abstract_mcts,
        CurveEnv=CurveEnv,
        DiscreteCurveEnv=DiscreteCurveEnv,
    )
    _loaded = ns
    return ns


def unload():
    """Drop the reference modules so the product's compat ``islaMcts`` can be imported."""
    global _loaded
    _loaded = None
    for k in [k for k in sys.modules if k == "islaMcts" or k.startswith("islaMcts.")]:
        del sys.modules[k]
