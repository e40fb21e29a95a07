"""Drop-in for islaMcts/agents/mcts_action_progressive_widening_hash.py (APW, continuous actions)."""
import numpy as np

from .. import _lib
from .abstract_mcts import AbstractMcts, _kind


class MctsActionProgressiveWideningHash(AbstractMcts):
    AGENT = _lib.AGENT_APW
    DISCRETE = False

    def _agent_kwargs(self, kw):
        p = self.param
        kw.update(alpha1=float(p.alpha), k1=float(p.k))
        fn = p.action_expansion_function
        kind = _kind(fn, "action_expansion_function")
        if kind == "random":
            kw.update(expansion=_lib.EXPAND_RANDOM)
        elif kind == "voo":
            kw.update(expansion=_lib.EXPAND_VOO, epsilon=fn.epsilon, voo_n_samples=fn.n_samples, voo_max_try=fn.max_try)
        elif kind == "genetic2":
            kw.update(expansion=_lib.EXPAND_GENETIC2, epsilon=fn.epsilon)
        elif kind == "genetic":
            kw.update(expansion=_lib.EXPAND_GENETIC, epsilon=fn.epsilon, voo_n_samples=fn.n_samples)
        else:
            raise NotImplementedError(f"action_expansion_function kind {kind!r} is not on the device yet")

    def _format_action(self, rank, action_value):
        space = getattr(self.param.env, "action_space", None)
        return np.array(action_value, dtype=getattr(space, "dtype", np.float32)).reshape(-1)   # policy.data (:44)
