from .mcts import Mcts
from .mcts_action_progressive_widening_hash import MctsActionProgressiveWideningHash
from .mcts_double_progressive_widening_hash import MctsDoubleProgressiveWideningHash
from .mcts_hash import MctsHash
from .mcts_state_progressive_widening_hash import MctsStateProgressiveWideningHash

# short aliases used in BASELINE.json's north_star
MctsApw = MctsActionProgressiveWideningHash
MctsSpw = MctsStateProgressiveWideningHash
MctsDpw = MctsDoubleProgressiveWideningHash

__all__ = ["Mcts", "MctsHash", "MctsApw", "MctsSpw", "MctsDpw", "MctsActionProgressiveWideningHash",
           "MctsStateProgressiveWideningHash", "MctsDoubleProgressiveWideningHash"]
