"""Drop-in for islaMcts/agents/parameters/mcts_parameters.py:8-31 (same field names and order).

Differences, all additive: ``x_values / y_values / depths`` get defaults so that the README form
(README.md:67-77) constructs; an optional ``state_variable`` is accepted (the reference's SPW/DPW agents read
it although the dataclass no longer declares it, mcts_parameters.py:23-24); keyword-only build options select
the B200 engine's extensions.
"""
from __future__ import annotations

from collections.abc import Callable
from dataclasses import dataclass, field
from typing import Any, Union


@dataclass(slots=True)
class MctsParameters:
    # data of the root node
    root_data: Any
    # simulation environment (gym-style; identified by the engine, see islamcts_b200.agents.abstract_mcts)
    env: Any
    # number of simulations from root node
    n_sim: int
    # exploration-exploitation factor
    C: Union[float, int]
    # the function to select actions (must be islamcts_b200.utils.action_selection_functions.ucb1)
    action_selection_fn: Callable
    # discount factor
    gamma: float
    # rollout policy (must be continuous_default_policy = env.action_space.sample())
    rollout_selection_fn: Callable
    # max number of steps during rollout
    max_depth: int
    # number of available actions
    n_actions: Union[None, int]
    x_values: list = field(default_factory=list, kw_only=True)
    y_values: list = field(default_factory=list, kw_only=True)
    depths: list = field(default_factory=list, kw_only=True)
    state_variable: str | None = field(default=None, kw_only=True)
    # ---- build options of the B200 engine (not in the reference)
    # "budgeted": rollouts to max_depth with MctsHash's depth accounting (mcts_hash.py:34,116,126);
    # "reference_head": the zero-length rollouts Mcts/APW/SPW/DPW perform at HEAD.
    # None = "budgeted", with a one-time warning from the agents whose HEAD behaviour differs; the compat package
    # ``islaMcts`` (code imported unchanged from the reference) defaults to "reference_head".
    depth_mode: str | None = field(default=None, kw_only=True)
    rollouts_per_leaf: int = field(default=1, kw_only=True)   # leaf-parallel batch: mean of R rollouts, ONE visit
    n_trees: int = field(default=1, kw_only=True)             # root-parallel ensemble of independent trees
    # tree-parallel search of ONE tree: that many simulations in flight at once (virtual loss; each leaf is provisionally
    # credited with `virtual_value` until its rollouts arrive).  1 = the reference's sequential loop (exact statistics).
    tree_parallel: int = field(default=1, kw_only=True)
    virtual_value: float = field(default=0.0, kw_only=True)
    seed: int | None = field(default=None, kw_only=True)      # Philox key; None = drawn from numpy's global RNG
    precision: str = field(default="f32", kw_only=True)       # env arithmetic on the device: "f32" | "f64"
    device: str = field(default="cuda:0", kw_only=True)
