"""gym-style environments whose ``step`` runs the SAME device dynamics the search uses (isla_env_step).

gym / gymnasium are optional: ``fit()`` accepts any gym-style object it can identify.  These classes exist
so that the episode loop of README.md:99-121 / sweep.py:185-227 runs on a box without gym -- the real step is
one launch of the env kernel on a batch of one (no CPU fallback).
"""
from __future__ import annotations

import numpy as np

from . import _lib
from . import engine as E


class Discrete:
    def __init__(self, n, seed=None):
        self.n, self.shape, self.dtype = int(n), (), np.int64
        self.np_random = np.random.default_rng(seed)

    def sample(self):
        return int(self.np_random.integers(self.n))


class Box:
    def __init__(self, low, high, shape=(1,), dtype=np.float32, seed=None):
        self.dtype, self.shape = np.dtype(dtype), tuple(shape)
        self.low, self.high = np.full(shape, low, dtype=dtype), np.full(shape, high, dtype=dtype)
        self.np_random = np.random.default_rng(seed)

    def sample(self):
        return self.np_random.uniform(self.low, self.high, self.shape).astype(self.dtype)


class DeviceEnv:
    """new_step_api: ``step`` returns (obs, reward, terminated, truncated, info); ``api=4`` gives the old 4-tuple."""

    def __init__(self, name: str, api: int = 5, seed=None, device="cuda:0", is_slippery: bool = False):
        self.spec_id = name
        self.is_slippery = bool(is_slippery)   # FrozenLake-v1 only: gym.make("FrozenLake-v1", is_slippery=True)
        self.env_id = E.ENV_NAMES[name]
        self.api, self.device = api, device
        self.np_random = np.random.default_rng(seed)
        n = E.N_ACTIONS[self.env_id]
        if n:
            self.action_space = Discrete(n, seed)
        else:
            lo, hi = E.ACTION_BOUNDS[self.env_id]
            self.action_space = Box(lo, hi, (1,), np.float32, seed)
        self.state = None
        self.s = 0
        if self.env_id == _lib.ENV_FROZENLAKE:
            self.desc = np.asarray([list(r) for r in ("SFFF", "FHFH", "FFFH", "HFFG")], dtype="U1")

    @property
    def unwrapped(self):
        return self

    def _obs(self):
        if self.env_id == _lib.ENV_FROZENLAKE:
            return int(self.s)
        if self.env_id == _lib.ENV_PENDULUM:
            th, thdot = self.state
            return np.array([np.cos(th), np.sin(th), thdot], dtype=np.float32)
        return np.asarray(self.state, dtype=np.float32)

    def reset(self, seed=None):
        if seed is not None:
            self.np_random = np.random.default_rng(seed)
        r = self.np_random
        if self.env_id == _lib.ENV_CARTPOLE:
            self.state = r.uniform(-0.05, 0.05, 4)
        elif self.env_id == _lib.ENV_PENDULUM:
            self.state = r.uniform([-np.pi, -1.0], [np.pi, 1.0])
        elif self.env_id == _lib.ENV_MOUNTAINCAR:
            self.state = np.array([r.uniform(-0.6, -0.4), 0.0], dtype=np.float32).astype(np.float64)
        else:
            self.s, self.state = 0, np.array([0.0])
        obs = self._obs()
        return obs if self.api == 4 else (obs, {})

    def step(self, action):
        a = float(np.asarray(action, dtype=np.float64).reshape(-1)[0])
        if self.env_id == _lib.ENV_FROZENLAKE and self.is_slippery:
            # gym's categorical_sample over the three outcomes a-1, a, a+1 (one draw of the env's generator per step)
            i = int(np.argmax(np.cumsum([1.0 / 3.0] * 3) > self.np_random.random()))
            a = float((int(a) - 1 + i) % 4)
        cur = np.array([float(self.s)]) if self.env_id == _lib.ENV_FROZENLAKE else np.asarray(self.state, np.float64)
        ns, r, t = E.env_step(self.env_id, cur[None, :], [a], precision=_lib.PRECISION_F64, device=self.device)
        self.state = ns[0].cpu().numpy()
        if self.env_id == _lib.ENV_FROZENLAKE:
            self.s = int(self.state[0])
        reward, term = float(r[0].item()), bool(t[0].item())
        if self.api == 4:
            return self._obs(), reward, term, {}
        return self._obs(), reward, term, False, {}


def make(name: str, api: int = 5, seed=None, device="cuda:0", is_slippery: bool = False) -> DeviceEnv:
    return DeviceEnv(name, api=api, seed=seed, device=device, is_slippery=is_slippery)
