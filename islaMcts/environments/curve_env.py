from islamcts_b200.environments.curve_env import Car, CurveEnv, DiscreteCurveEnv  # noqa: F401
