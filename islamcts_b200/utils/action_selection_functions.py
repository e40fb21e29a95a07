"""Drop-in for islaMcts/utils/action_selection_functions.py.

The reference stores Python callables in the params and calls them on every node visit
(agents/mcts.py:61, agents/abstract_mcts.py:84, mcts_action_...:63).  Device code cannot call Python, so the
objects exported here are *tags*: each carries ``isla_kind`` (and its parameters) and the agent maps it to an
engine enum.  They remain callable on the host proxies for compatibility, but the search itself never calls
them.  Unknown callables make ``fit()`` raise NotImplementedError -- there is no CPU fallback by design.
"""
from __future__ import annotations

import numpy as np


def ucb1(node):
    """UCB1 child selection (reference :9-25).  Callable on a host node proxy; tag for the device engine."""
    n_visits = node.ns
    children = list(node.actions.values())
    visit_child = np.array([c.na for c in children])
    node_values = np.array([c.total / c.na for c in children])
    score = node_values + node.param.C * np.sqrt(np.log(n_visits) / visit_child)
    index = np.random.choice(np.flatnonzero(score == score.max()))
    return children[index].data


ucb1.isla_kind = "ucb1"


def continuous_default_policy(node, *args, **kwargs):
    """``env.action_space.sample()`` (reference :46-47); used for discrete spaces too (README.md:75)."""
    return node.param.env.action_space.sample()


continuous_default_policy.isla_kind = "random"


class _Voo:
    """Product of ``voo(...)`` (reference :139-205): a tag carrying epsilon / n_samples / max_try."""

    isla_kind = "voo"

    def __init__(self, epsilon, default_policy, n_samples, max_try):
        if getattr(default_policy, "isla_kind", None) != "random":
            raise NotImplementedError("voo: only continuous_default_policy is supported as the default policy")
        self.epsilon, self.default_policy, self.n_samples, self.max_try = float(epsilon), default_policy, int(n_samples), int(max_try)

    def __call__(self, node):
        raise NotImplementedError("voo runs on the device inside fit(); it is not evaluated on the host")


def voo(epsilon, default_policy, n_samples, max_try=1000):
    return _Voo(epsilon, default_policy, n_samples, max_try)


class _Genetic2:
    """Product of ``genetic_policy2(epsilon, default_policy)`` (reference :104-136): a tag carrying epsilon."""

    isla_kind = "genetic2"

    def __init__(self, epsilon, default_policy):
        if getattr(default_policy, "isla_kind", None) != "random":
            raise NotImplementedError("genetic_policy2: only continuous_default_policy is supported as the default policy")
        self.epsilon, self.default_policy = float(epsilon), default_policy

    def __call__(self, node):
        raise NotImplementedError("genetic_policy2 runs on the device inside fit(); it is not evaluated on the host")


def genetic_policy2(epsilon, default_policy):
    return _Genetic2(epsilon, default_policy)


def _next_round(name):
    def factory(*args, **kwargs):
        raise NotImplementedError(f"{name}: not on the B200 hot path yet (SURVEY section 8f); no CPU fallback by design")
    factory.__name__ = name
    return factory


class _Genetic:
    """Product of ``genetic_policy(epsilon, default_policy, n_samples)`` (reference :50-102): a tag."""

    isla_kind = "genetic"

    def __init__(self, epsilon, default_policy, n_samples):
        if getattr(default_policy, "isla_kind", None) != "random":
            raise NotImplementedError("genetic_policy: only continuous_default_policy is supported as the default policy")
        self.epsilon, self.default_policy, self.n_samples = float(epsilon), default_policy, int(n_samples)

    def __call__(self, node):
        raise NotImplementedError("genetic_policy runs on the device inside fit(); it is not evaluated on the host")


def genetic_policy(epsilon, default_policy, n_samples):
    return _Genetic(epsilon, default_policy, n_samples)

class _GridPolicy:
    """Product of ``grid_policy(prior_knowledge, n_actions)`` (reference :258-282): a rollout policy for FrozenLake that
    takes an arg-min of ``prior_knowledge[state]`` (ties uniformly).  A tag carrying the table; it runs on the device."""

    isla_kind = "grid"

    def __init__(self, prior_knowledge, n_actions):
        import numpy as np
        self.n_actions = int(n_actions)
        states = sorted(prior_knowledge)
        if states != list(range(len(states))):
            raise ValueError("grid_policy: prior_knowledge must be keyed by the states 0 .. n-1")
        self.table = np.array([list(prior_knowledge[s]) for s in states], dtype=np.float64)
        if self.table.shape[1] != self.n_actions:
            raise ValueError("grid_policy: every state needs one value per action")

    def __call__(self, node):
        raise NotImplementedError("grid_policy runs on the device inside fit(); it is not evaluated on the host")


def grid_policy(prior_knowledge, n_actions):
    return _GridPolicy(prior_knowledge, n_actions)
discrete_default_policy = _next_round("discrete_default_policy")
