"""Drop-in for islaMcts/agents/mcts_hash.py: vanilla MCTS with ndarray observations (CartPole)."""
from .abstract_mcts import AbstractActionNode as ActionNodeHash  # noqa: F401
from .abstract_mcts import AbstractMcts
from .abstract_mcts import AbstractStateNode as StateNodeHash  # noqa: F401


class MctsHash(AbstractMcts):
    """``MctsHash(param).fit()`` (reference agents/mcts_hash.py:22-50): the only agent whose rollouts are
    depth-budgeted at HEAD, so both depth modes behave the same."""
    HEAD_ZERO_ROLLOUTS = False
