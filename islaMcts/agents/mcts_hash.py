from islamcts_b200.agents.mcts_hash import *  # noqa: F401,F403
from islamcts_b200.agents import mcts_hash as _m
globals().update({k: v for k, v in vars(_m).items() if not k.startswith('__')})
