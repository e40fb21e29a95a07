"""Drop-in for islaMcts/agents/parameters/dpw_parameters.py:6-13."""
from dataclasses import dataclass

from .mcts_parameters import MctsParameters


@dataclass(slots=True)
class DpwParameters(MctsParameters):
    # Action Progressive Widening Parameters
    alphaApw: float
    kApw: int
    # State Progressive Widening Parameters
    alphaSpw: float
    kSpw: int
