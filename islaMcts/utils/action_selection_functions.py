from islamcts_b200.utils.action_selection_functions import *  # noqa: F401,F403
from islamcts_b200.utils import action_selection_functions as _m
globals().update({k: v for k, v in vars(_m).items() if not k.startswith('__')})
