"""Drop-in for islaMcts/utils/agent_factory.py:16-46 (same strings, same hashable-observation dispatch)."""
from collections.abc import Hashable

from ..agents.mcts import Mcts
from ..agents.mcts_action_progressive_widening_hash import MctsActionProgressiveWideningHash
from ..agents.mcts_double_progressive_widening_hash import MctsDoubleProgressiveWideningHash
from ..agents.mcts_hash import MctsHash
from ..agents.mcts_state_progressive_widening_hash import MctsStateProgressiveWideningHash


def get_agent(agent_type: str, params):
    hashable = isinstance(params.root_data, Hashable)
    if agent_type == "vanilla":
        return Mcts(params) if hashable else MctsHash(params)
    if agent_type == "spw" and not hashable:
        return MctsStateProgressiveWideningHash(params)
    if agent_type == "apw" and not hashable:
        return MctsActionProgressiveWideningHash(params)
    if agent_type == "dpw" and not hashable:
        return MctsDoubleProgressiveWideningHash(params)
    raise NotImplementedError   # "random", "optimal_goddard" and unknown names: not searches
