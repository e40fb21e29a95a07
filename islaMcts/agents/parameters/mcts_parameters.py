"""Compat path of the reference's islaMcts/agents/parameters/mcts_parameters.py.  Code imported unchanged from the reference
gets the reference's HEAD semantics by default: depth_mode="reference_head" (Mcts / APW / SPW / DPW play zero-length
rollouts); islamcts_b200's own classes default to budgeted rollouts."""
from dataclasses import dataclass, field

from islamcts_b200.agents.parameters import mcts_parameters as _m


@dataclass(slots=True)
class MctsParameters(_m.MctsParameters):
    depth_mode: str | None = field(default="reference_head", kw_only=True)
