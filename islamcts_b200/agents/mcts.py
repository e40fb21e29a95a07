"""Drop-in for islaMcts/agents/mcts.py: vanilla MCTS, discrete actions, hashable observations (FrozenLake)."""
from .abstract_mcts import AbstractActionNode as ActionNode  # noqa: F401  (names the reference exports)
from .abstract_mcts import AbstractMcts
from .abstract_mcts import AbstractStateNode as StateNode  # noqa: F401


class Mcts(AbstractMcts):
    """``Mcts(param).fit()`` (reference agents/mcts.py:18-38).  At reference HEAD this class starts the recursion at
    ``max_depth`` so its rollouts have zero length (``depth_mode="reference_head"``); the default
    ``depth_mode="budgeted"`` plays them to ``max_depth`` as north_star asks."""
    HEAD_ZERO_ROLLOUTS = True
    HASH_KEYS = False              # children[observation] (agents/mcts.py:83-84), not observation.tobytes()
