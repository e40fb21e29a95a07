"""Drop-in for islaMcts/agents/mcts_state_progressive_widening_hash.py (SPW, discrete actions)."""
from .. import _lib
from .abstract_mcts import AbstractMcts


class MctsStateProgressiveWideningHash(AbstractMcts):
    AGENT = _lib.AGENT_SPW

    def _agent_kwargs(self, kw):
        kw.update(alpha1=float(self.param.alpha), k1=float(self.param.k))

    def _format_action(self, rank, action_value):
        return int(rank)   # the INDEX into the key-sorted root actions (:38)
