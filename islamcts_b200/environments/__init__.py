"""The reference's own environments (islaMcts/environments/) with the step arithmetic on the device."""
from .curve_env import Car, CurveEnv, DiscreteCurveEnv  # noqa: F401
