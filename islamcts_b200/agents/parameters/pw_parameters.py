"""Drop-in for islaMcts/agents/parameters/pw_parameters.py:7-11."""
from dataclasses import dataclass
from typing import Callable

from .mcts_parameters import MctsParameters


@dataclass(slots=True)
class PwParameters(MctsParameters):
    alpha: float
    k: int
    action_expansion_function: Callable
