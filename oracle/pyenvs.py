"""TEST INFRASTRUCTURE ONLY -- float64 Python restatement of the Gymnasium envs.

The environment arithmetic of the IslaMcts hot path is NOT in the reference: it
lives in third-party ``gymnasium`` (unpinned, /root/reference/requirements.txt:14)
and legacy ``gym`` (/root/reference/islaMcts/agents/parameters/mcts_parameters.py:5),
neither of which is installed here.  The reference's own tests mock ``env.step``
(/root/reference/tests/agents/test_mcts.py:153), so **env parity is unpinned**;
this file restates the published Gymnasium 1.x classic-control / toy-text
definitions (equations in SURVEY.md section 8c) and is the spec the CUDA env
kernels are checked against.

Reference call sites this replaces: ``env.step`` at
/root/reference/islaMcts/agents/mcts.py:79, mcts_hash.py:93,
abstract_mcts.py:85; ``action_space.sample`` at
/root/reference/islaMcts/utils/action_selection_functions.py:47; the clone at
/root/reference/islaMcts/utils/mcts_utils.py:10-15 (objects are picklable).

Conventions: envs are used ``.unwrapped`` (no TimeLimit); ``api=4`` returns the
old ``(obs, reward, done, info)`` tuple that MctsHash/APW/SPW/DPW unpack
(mcts_hash.py:93), ``api=5`` the Gymnasium 5-tuple.
"""
from __future__ import annotations

import math

import numpy as np

ENV_CARTPOLE = 0
ENV_PENDULUM = 1
ENV_MOUNTAINCAR = 2
ENV_FROZENLAKE = 3

ENV_IDS = {
    "CartPole-v1": ENV_CARTPOLE,
    "Pendulum-v1": ENV_PENDULUM,
    "MountainCarContinuous-v0": ENV_MOUNTAINCAR,
    "FrozenLake-v1": ENV_FROZENLAKE,
}


class Discrete:
    """gymnasium.spaces.Discrete: sample() = uniform int in [0, n)."""

    def __init__(self, n: int, seed=None):
        self.n = int(n)
        self.shape = ()
        self.dtype = np.int64
        self.np_random = np.random.default_rng(seed)

    def seed(self, seed=None):
        self.np_random = np.random.default_rng(seed)

    def sample(self):
        return int(self.np_random.integers(self.n))


class Box:
    """gymnasium.spaces.Box (bounded): sample() = U[low, high] cast to dtype."""

    def __init__(self, low, high, shape, dtype=np.float32, seed=None):
        self.dtype = np.dtype(dtype)
        self.shape = tuple(shape)
        self.low = np.full(self.shape, low, dtype=self.dtype)
        self.high = np.full(self.shape, high, dtype=self.dtype)
        self.np_random = np.random.default_rng(seed)

    def seed(self, seed=None):
        self.np_random = np.random.default_rng(seed)

    def sample(self):
        return self.np_random.uniform(self.low, self.high, self.shape).astype(self.dtype)


# env.step calls counted across ALL instances (the reference steps pickled clones)
STEP_COUNTER = [0]


class _Env:
    env_id: int = -1
    spec_id: str = ""

    def __init__(self, api: int = 5):
        assert api in (4, 5)
        self.api = api
        self.n_steps = 0  # counted env.step calls (CPU baseline env-steps/s)

    @property
    def unwrapped(self):
        return self

    def _ret(self, obs, reward, terminated):
        self.n_steps += 1
        STEP_COUNTER[0] += 1
        if self.api == 4:
            return obs, reward, terminated, {}
        return obs, reward, terminated, False, {}


class CartPoleEnv(_Env):
    """CartPole-v1 (Gymnasium classic_control/cartpole.py, Euler integrator)."""

    env_id = ENV_CARTPOLE
    spec_id = "CartPole-v1"
    gravity = 9.8
    masscart = 1.0
    masspole = 0.1
    total_mass = masspole + masscart
    length = 0.5
    polemass_length = masspole * length
    force_mag = 10.0
    tau = 0.02
    theta_threshold_radians = 12 * 2 * math.pi / 360
    x_threshold = 2.4

    def __init__(self, api: int = 5, seed=None):
        super().__init__(api)
        self.action_space = Discrete(2, seed)
        self.np_random = np.random.default_rng(seed)
        self.state = None

    def reset(self, seed=None):
        if seed is not None:
            self.np_random = np.random.default_rng(seed)
        self.state = self.np_random.uniform(-0.05, 0.05, size=(4,)).astype(np.float64)
        obs = np.array(self.state, dtype=np.float32)
        return obs if self.api == 4 else (obs, {})

    def step(self, action):
        x, x_dot, theta, theta_dot = (float(v) for v in self.state)
        force = self.force_mag if int(action) == 1 else -self.force_mag
        costheta = math.cos(theta)
        sintheta = math.sin(theta)
        temp = (force + self.polemass_length * (theta_dot * theta_dot) * sintheta) / self.total_mass
        thetaacc = (self.gravity * sintheta - costheta * temp) / (
            self.length * (4.0 / 3.0 - self.masspole * (costheta * costheta) / self.total_mass)
        )
        xacc = temp - self.polemass_length * thetaacc * costheta / self.total_mass
        x = x + self.tau * x_dot
        x_dot = x_dot + self.tau * xacc
        theta = theta + self.tau * theta_dot
        theta_dot = theta_dot + self.tau * thetaacc
        self.state = np.array((x, x_dot, theta, theta_dot), dtype=np.float64)
        terminated = bool(
            x < -self.x_threshold
            or x > self.x_threshold
            or theta < -self.theta_threshold_radians
            or theta > self.theta_threshold_radians
        )
        return self._ret(np.array(self.state, dtype=np.float32), 1.0, terminated)


class PendulumEnv(_Env):
    """Pendulum-v1 (Gymnasium order: clip the speed, then advance the angle)."""

    env_id = ENV_PENDULUM
    spec_id = "Pendulum-v1"
    max_speed = 8.0
    max_torque = 2.0
    dt = 0.05
    g = 10.0
    m = 1.0
    l = 1.0

    def __init__(self, api: int = 5, seed=None):
        super().__init__(api)
        self.action_space = Box(-self.max_torque, self.max_torque, (1,), np.float32, seed)
        self.np_random = np.random.default_rng(seed)
        self.state = None

    def _obs(self):
        th, thdot = self.state
        return np.array([math.cos(th), math.sin(th), thdot], dtype=np.float32)

    def reset(self, seed=None):
        if seed is not None:
            self.np_random = np.random.default_rng(seed)
        self.state = self.np_random.uniform([-math.pi, -1.0], [math.pi, 1.0]).astype(np.float64)
        return self._obs() if self.api == 4 else (self._obs(), {})

    def step(self, u):
        th, thdot = (float(v) for v in self.state)
        u = float(np.asarray(u, dtype=np.float64).reshape(-1)[0])
        u = min(max(u, -self.max_torque), self.max_torque)
        ang = math.fmod(th + math.pi, 2 * math.pi)
        if ang < 0:  # python-style modulo (numpy % on floats)
            ang += 2 * math.pi
        ang -= math.pi
        costs = ang * ang + 0.1 * (thdot * thdot) + 0.001 * (u * u)
        newthdot = thdot + (3 * self.g / (2 * self.l) * math.sin(th) + 3.0 / (self.m * self.l**2) * u) * self.dt
        newthdot = min(max(newthdot, -self.max_speed), self.max_speed)
        newth = th + newthdot * self.dt
        self.state = np.array([newth, newthdot], dtype=np.float64)
        return self._ret(self._obs(), -costs, False)


class MountainCarContinuousEnv(_Env):
    """MountainCarContinuous-v0; state is rounded to float32 at the end of each step."""

    env_id = ENV_MOUNTAINCAR
    spec_id = "MountainCarContinuous-v0"
    min_action = -1.0
    max_action = 1.0
    min_position = -1.2
    max_position = 0.6
    max_speed = 0.07
    goal_position = 0.45
    goal_velocity = 0.0
    power = 0.0015

    def __init__(self, api: int = 5, seed=None):
        super().__init__(api)
        self.action_space = Box(self.min_action, self.max_action, (1,), np.float32, seed)
        self.np_random = np.random.default_rng(seed)
        self.state = None

    def reset(self, seed=None):
        if seed is not None:
            self.np_random = np.random.default_rng(seed)
        self.state = np.array([self.np_random.uniform(-0.6, -0.4), 0.0], dtype=np.float32)
        obs = np.array(self.state, dtype=np.float32)
        return obs if self.api == 4 else (obs, {})

    def step(self, action):
        position = float(self.state[0])
        velocity = float(self.state[1])
        a0 = float(np.asarray(action, dtype=np.float64).reshape(-1)[0])
        force = min(max(a0, self.min_action), self.max_action)
        velocity += force * self.power - 0.0025 * math.cos(3 * position)
        velocity = min(max(velocity, -self.max_speed), self.max_speed)
        position += velocity
        position = min(max(position, self.min_position), self.max_position)
        if position == self.min_position and velocity < 0:
            velocity = 0.0
        terminated = bool(position >= self.goal_position and velocity >= self.goal_velocity)
        reward = 100.0 if terminated else 0.0
        reward -= (a0 * a0) * 0.1
        self.state = np.array([position, velocity], dtype=np.float32)
        return self._ret(np.array(self.state, dtype=np.float32), reward, terminated)


class FrozenLakeEnv(_Env):
    """FrozenLake-v1 4x4.  0 LEFT, 1 DOWN, 2 RIGHT, 3 UP.  is_slippery=True: the action carried out is a-1, a or a+1
    (mod 4), each with probability 1/3, chosen like gym's categorical_sample: argmax(cumsum(p) > np_random.random()),
    one draw of the env's own generator per step (terminal cells included)."""

    env_id = ENV_FROZENLAKE
    spec_id = "FrozenLake-v1"
    MAP = ("SFFF", "FHFH", "FFFH", "HFFG")

    def __init__(self, api: int = 5, seed=None, is_slippery: bool = False):
        super().__init__(api)
        self.desc = np.asarray([list(r) for r in self.MAP], dtype="U1")
        self.nrow, self.ncol = self.desc.shape
        self.action_space = Discrete(4, seed)
        self.is_slippery = bool(is_slippery)
        self.np_random = np.random.default_rng(seed)
        self.s = 0

    @property
    def state(self):
        return self.s

    def reset(self, seed=None):
        self.s = 0
        return int(self.s) if self.api == 4 else (int(self.s), {})

    def step(self, a):
        row, col = divmod(int(self.s), self.ncol)
        u = self.np_random.random() if self.is_slippery else 0.0
        if self.desc[row, col] in "GH":
            # stepping from a terminal cell: stays put, reward 0, terminated
            return self._ret(int(self.s), 0.0, True)
        a = int(a)
        if self.is_slippery:
            i = int(np.argmax(np.cumsum([1.0 / 3.0] * 3) > u))
            a = (a - 1 + i) % 4
        if a == 0:
            col = max(col - 1, 0)
        elif a == 1:
            row = min(row + 1, self.nrow - 1)
        elif a == 2:
            col = min(col + 1, self.ncol - 1)
        elif a == 3:
            row = max(row - 1, 0)
        self.s = row * self.ncol + col
        letter = self.desc[row, col]
        return self._ret(int(self.s), float(letter == "G"), bool(letter in "GH"))


ENV_CLASSES = {
    ENV_CARTPOLE: CartPoleEnv,
    ENV_PENDULUM: PendulumEnv,
    ENV_MOUNTAINCAR: MountainCarContinuousEnv,
    ENV_FROZENLAKE: FrozenLakeEnv,
}


def make(name_or_id, api: int = 5, seed=None) -> _Env:
    env_id = ENV_IDS[name_or_id] if isinstance(name_or_id, str) else int(name_or_id)
    return ENV_CLASSES[env_id](api=api, seed=seed)


def env_state_vector(env: _Env) -> np.ndarray:
    """Root state as the float64 vector the engines take (FrozenLake: [s])."""
    if env.env_id == ENV_FROZENLAKE:
        return np.array([float(env.s)], dtype=np.float64)
    return np.asarray(env.state, dtype=np.float64).copy()


def set_env_state(env: _Env, vec) -> None:
    if env.env_id == ENV_FROZENLAKE:
        env.s = int(vec[0])
    elif env.env_id == ENV_MOUNTAINCAR:
        env.state = np.asarray(vec, dtype=np.float32).copy()
    else:
        env.state = np.asarray(vec, dtype=np.float64).copy()
