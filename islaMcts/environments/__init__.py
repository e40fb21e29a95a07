from islamcts_b200.environments import Car, CurveEnv, DiscreteCurveEnv  # noqa: F401
