"""Whole episodes on the device: fit -> real step -> new root, for many episodes in lock-step, without host round trips.

Mirrors the episode loop of the reference's CLIs (sweep.py:166-227, sweep_vanilla.py, README.md:99-121): every real step
builds a FRESH tree per episode (``agent = create_agent()`` inside the loop, sweep.py:186), takes ``agent.fit()``'s action
and applies it to the real env.  Here the E episodes of an experiment (the reference runs ``--n_episodes`` of them one
after the other) advance together: one search launch for the E trees, one ``best`` epilogue, one env-step kernel that
applies the chosen actions, accumulates rewards and freezes finished episodes.  The host only enqueues launches; it
looks at the ``alive`` flags every few steps to stop early.
"""
from __future__ import annotations

import ctypes as C

import numpy as np
import torch

from . import _lib
from . import engine as E
from ._lib import check


class EpisodeBatch:
    def __init__(self, env: str | int, agent: int = _lib.AGENT_VANILLA, n_episodes: int = 100, n_sim: int = 1000,
                 max_steps: int = 100, seed: int = 0, device="cuda:0", action_grid=None, timeout_reward: float | None = -1000.0,
                 reuse_tree: bool = False, **engine_kwargs):
        self.env_id = E.ENV_NAMES[env] if isinstance(env, str) else int(env)
        self.device = torch.device(device)
        self.n, self.n_sim, self.max_steps, self.seed = int(n_episodes), int(n_sim), int(max_steps), int(seed)
        self.grid = tuple(int(x) for x in action_grid) if action_grid else (0, 0)
        # the reference's loop books reward = -1000 for the step that reaches the limit (sweep.py:206-209); None = env reward
        self.timeout_reward = float("nan") if timeout_reward is None else float(timeout_reward)
        # reuse_tree (opt-in build extension): instead of a fresh tree per real step (the reference's loop, sweep.py:186 --
        # the default here), the subtree below the chosen root action is kept and searched on (isla_engine_reroot;
        # thread-per-tree engine: vanilla search on CartPole / FrozenLake).  The search then continues from the state the
        # TREE stored for that child (computed in the search's precision), the real env advances in float64 beside it.
        self.reuse_tree = bool(reuse_tree)
        if self.reuse_tree:
            engine_kwargs.setdefault("kernel", _lib.KERNEL_THREAD_PER_TREE)
        self.engine = E.Engine(self.env_id, agent=agent, n_trees=self.n, sim_capacity=self.n_sim * (8 if self.reuse_tree else 1) + 2,
                               seed=seed, device=device, action_grid=action_grid, **engine_kwargs)
        self.S = E.STATE_DIM[self.env_id]

    def run(self, initial_states, check_every: int = 8) -> dict:
        """initial_states: [n_episodes, state_dim] float64 env states.  Returns host arrays: total_reward, n_steps, final
        states, and the per-step actions (rank of the chosen root child is not kept)."""
        dev, eng, lib = self.device, self.engine, _lib.load()
        states = torch.as_tensor(initial_states, dtype=torch.float64).reshape(self.n, self.S).to(dev).clone()
        alive = torch.ones(self.n, dtype=torch.uint8, device=dev)
        total = torch.zeros(self.n, dtype=torch.float64, device=dev)
        steps = torch.zeros(self.n, dtype=torch.int32, device=dev)
        actions = []
        for k in range(self.max_steps):
            if self.reuse_tree and k > 0:
                eng.reroot(act.to(torch.int32), self.n_sim)    # keep the subtree below the action just played (every tree)
            else:
                eng.reseed(self.seed + k, 0)                   # every real step searches with its own streams
                eng.set_roots(states)
            eng.run(self.n_sim)
            _, act = eng.best()
            actions.append(act)
            with torch.cuda.device(dev):
                check(lib.isla_episode_step(self.env_id, self.grid[0], self.grid[1], C.c_void_p(states.data_ptr()),
                                            C.c_void_p(act.data_ptr()), self.n, self.max_steps, self.timeout_reward, C.c_void_p(alive.data_ptr()),
                                            C.c_void_p(total.data_ptr()), C.c_void_p(steps.data_ptr()),
                                            C.c_void_p(E._stream_ptr(dev))))
            if (k + 1) % check_every == 0 and not bool(alive.any().item()):
                break
        cnt = eng.counters()
        if cnt["faults"]:
            raise _lib.IslaError(_lib.ERR_CAPACITY, f"device search reported faults: {cnt}")
        return dict(total_reward=total.cpu().numpy(), n_steps=steps.cpu().numpy(), states=states.cpu().numpy(),
                    actions=torch.stack(actions).cpu().numpy(), real_steps=int(steps.sum().item()))
