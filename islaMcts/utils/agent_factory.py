from islamcts_b200.utils.agent_factory import *  # noqa: F401,F403
from islamcts_b200.utils import agent_factory as _m
globals().update({k: v for k, v in vars(_m).items() if not k.startswith('__')})
