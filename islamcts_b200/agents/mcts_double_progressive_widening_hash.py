"""Drop-in for islaMcts/agents/mcts_double_progressive_widening_hash.py (DPW, continuous actions)."""
import numpy as np

from .. import _lib
from .abstract_mcts import AbstractMcts


class MctsDoubleProgressiveWideningHash(AbstractMcts):
    AGENT = _lib.AGENT_DPW
    DISCRETE = False

    def _agent_kwargs(self, kw):
        p = self.param
        kw.update(alpha1=float(p.alphaApw), k1=float(p.kApw), alpha2=float(p.alphaSpw), k2=float(p.kSpw))

    def _format_action(self, rank, action_value):
        # the reference decodes the key with dtype=float (:41): a float64 array for float64 spaces; for float32
        # spaces that line raises at HEAD.  Here the action comes back in the space's own dtype, so that
        # ``agent.root.actions[action.tobytes()]`` (sweep.py:196) finds the node.
        space = getattr(self.param.env, "action_space", None)
        return np.array(action_value, dtype=getattr(space, "dtype", np.float64)).reshape(-1)
