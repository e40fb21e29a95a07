"""Drop-in for islaMcts/environments/curve_env.py: CurveEnv / DiscreteCurveEnv whose ``step`` runs the device
dynamics the search itself uses (isla_env_step, float64, csrc/isla_envs.cuh Env<ENV_CURVE>).

Interface kept from the reference (curve_env.py:22-199): ``env.car.state`` = [x, y, velocity, angle], the hidden
``env.n_step`` counter that scales the rewards and that ``reset()`` does NOT clear (:179-182), the 4-tuple
``(state, reward, terminal, None)`` returned by ``step`` (:177), ``action_space`` = Box([-5, -30], [5, 30], float64)
or Discrete(g0 * g1) with ``env.actions`` the (acceleration, angle) table (:184-199).  No CPU fallback: stepping
needs the CUDA library.
"""
from __future__ import annotations

import itertools

import numpy as np

from .. import _lib
from .. import engine as E
from ..envs import Box, Discrete

START_X, START_Y, START_VELOCITY, START_ANGLE = -11.0, 21.0, 10.0, 90.0
ACTION_LOW, ACTION_HIGH = (-5.0, -30.0), (5.0, 30.0)


class Car:
    """Pose record of the car; the arithmetic of the reference's Car.update lives in the device kernel."""

    def __init__(self, x, y, angle=START_ANGLE):
        self.position = np.array([x, y], dtype=float)
        self.velocity = START_VELOCITY
        self.angle = angle
        self.max_velocity, self.min_velocity = 20.0, 0.0
        self.state = np.array([*self.position, self.velocity, self.angle])

    def _load(self, vec):
        self.position = np.array(vec[:2], dtype=float)
        self.velocity, self.angle = float(vec[2]), float(vec[3])
        self.state = np.array([*self.position, self.velocity, self.angle])


class CurveEnv:
    def __init__(self, device="cuda:0", seed=None):
        self.device = device
        self.n_step = 0
        self.dt = 0.5
        self.target_x = 30
        self.car = None
        self.action_space = Box(np.array(ACTION_LOW), np.array(ACTION_HIGH), (2,), np.float64, seed)
        self.reset()

    @property
    def unwrapped(self):
        return self

    @staticmethod
    def higher_bound(x):
        return -(1 / 50) * (x ** 2) + 27.5

    @staticmethod
    def lower_bound(x):
        return -(1 / 50) * (3 * (x ** 2)) + 27.3

    def reset(self, *, seed=None, return_info=False, options=None):
        self.car = Car(START_X, START_Y)
        return self.car.state

    def step(self, action):
        a = np.asarray(action, dtype=np.float64).reshape(-1)
        if a.size != 2:
            raise ValueError("CurveEnv.step takes (acceleration, steering angle)")
        cur = np.array([*self.car.state, float(self.n_step)], dtype=np.float64)
        ns, r, t = E.env_step(_lib.ENV_CURVE, cur[None, :], a[None, :], precision=_lib.PRECISION_F64, device=self.device)
        vec = ns[0].cpu().numpy()
        self.car._load(vec)
        self.n_step = int(vec[4])
        return self.car.state, float(r[0].item()), bool(t[0].item()), None


class DiscreteCurveEnv(CurveEnv):
    def __init__(self, n_actions, device="cuda:0", seed=None):
        super().__init__(device=device, seed=seed)
        g0, g1 = int(n_actions[0]), int(n_actions[1])
        self.grid = (g0, g1)
        self.actions = list(itertools.product(np.linspace(ACTION_LOW[0], ACTION_HIGH[0], g0),
                                              np.linspace(ACTION_LOW[1], ACTION_HIGH[1], g1)))
        self.action_space = Discrete(len(self.actions), seed)

    def step(self, discrete_action):
        return super().step(np.array(self.actions[int(discrete_action)]))
