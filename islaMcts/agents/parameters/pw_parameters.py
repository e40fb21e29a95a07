from islamcts_b200.agents.parameters.pw_parameters import *  # noqa: F401,F403
from islamcts_b200.agents.parameters import pw_parameters as _m
globals().update({k: v for k, v in vars(_m).items() if not k.startswith('__')})
