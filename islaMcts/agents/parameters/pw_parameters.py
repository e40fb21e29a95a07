"""Compat path of islaMcts/agents/parameters/pw_parameters.py (depth_mode defaults to "reference_head")."""
from dataclasses import dataclass, field

from islamcts_b200.agents.parameters import pw_parameters as _m


@dataclass(slots=True)
class PwParameters(_m.PwParameters):
    depth_mode: str | None = field(default="reference_head", kw_only=True)
