"""Compat path of islaMcts/agents/parameters/dpw_parameters.py (depth_mode defaults to "reference_head")."""
from dataclasses import dataclass, field

from islamcts_b200.agents.parameters import dpw_parameters as _m


@dataclass(slots=True)
class DpwParameters(_m.DpwParameters):
    depth_mode: str | None = field(default="reference_head", kw_only=True)
