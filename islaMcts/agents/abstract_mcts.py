from islamcts_b200.agents.abstract_mcts import *  # noqa: F401,F403
from islamcts_b200.agents import abstract_mcts as _m
globals().update({k: v for k, v in vars(_m).items() if not k.startswith('__')})
