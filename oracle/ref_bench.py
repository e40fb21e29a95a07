"""TEST / BASELINE INFRASTRUCTURE ONLY -- the reference's OWN agents as the CPU arm of bench.py.

``stage()`` puts the reference's package (unmodified files) under the git-ignored ``baseline/_ref/`` so that it travels
to the GPU box with the repo snapshot (``/root/reference`` does not exist there).  pip cannot install it: the reference
has no setup.py / pyproject.toml (only requirements.txt), so the package directory is copied as it is -- the recorded
pip outcome is in DESIGN.md.  ``time_fit()`` builds the reference's real classes through oracle/ref_loader.py (stub
graphviz / gym / gymnasium modules, the two SURVEY 8c shims for SPW/DPW) on the float64 env restatement
(oracle/pyenvs.py, old 4-tuple step API) and times ``agent.fit()`` -- BASELINE.md section 3.
"""
from __future__ import annotations

import os
import shutil
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
STAGED = os.path.join(ROOT, "baseline", "_ref")
SOURCE = "/root/reference"


def stage(force: bool = False) -> str | None:
    """Copy /root/reference/islaMcts (the package, as it is) to baseline/_ref/islaMcts.  Returns the staged root or None."""
    src = os.path.join(SOURCE, "islaMcts")
    dst = os.path.join(STAGED, "islaMcts")
    if not os.path.isdir(src):
        return STAGED if os.path.isdir(dst) else None
    if force or not os.path.isdir(dst):
        if os.path.isdir(dst):
            shutil.rmtree(dst)
        os.makedirs(STAGED, exist_ok=True)
        shutil.copytree(src, dst, ignore=shutil.ignore_patterns("__pycache__", "*.pyc"))
    return STAGED


def reference_root() -> str | None:
    """Where the real reference can be imported from in this process (None: only the port is available)."""
    for r in (os.environ.get("ISLA_REFERENCE_ROOT"), SOURCE, STAGED):
        if r and os.path.isdir(os.path.join(r, "islaMcts", "agents")):
            return r
    return None


# the BASELINE.json configurations as the reference's own classes run them (HEAD semantics: Mcts / APW / DPW start their
# recursion at max_depth, i.e. zero-length rollouts -- SURVEY section 0 quirk 1; MctsHash plays budgeted rollouts)
CONFIGS = {
    "c1": dict(agent="Mcts", env="FrozenLake-v1", n_sim=100, C=0.5, gamma=1.0, max_depth=1000),
    "c2": dict(agent="MctsHash", env="CartPole-v1", n_sim=1000, C=1.0, gamma=0.99, max_depth=500),
    "c3": dict(agent="MctsApw", env="Pendulum-v1", n_sim=10000, C=1.0, gamma=0.99, max_depth=200, alpha=0.8, k=60),
    "c4": dict(agent="MctsDpw", env="MountainCarContinuous-v0", n_sim=10000, C=1.0, gamma=0.99, max_depth=500,
               alpha1=0.5, k1=1, alpha2=0.5, k2=1),
    "c5": dict(agent="MctsHash", env="CartPole-v1", n_sim=1000, C=1.0, gamma=0.99, max_depth=500),
}


def build_agent(R, sc: dict, seed: int):
    from oracle import pyenvs
    api = 5 if sc["agent"] == "Mcts" else 4   # mcts.py:79-82 indexes vals[0..2]; the others unpack 4 (mcts_hash.py:93)
    env = pyenvs.make(sc["env"], api=api, seed=seed)
    obs = env.reset()
    if api == 5:
        obs = obs[0]
    common = dict(root_data=obs, env=env, n_sim=sc["n_sim"], C=sc["C"], action_selection_fn=R.ucb1, gamma=sc["gamma"],
                  rollout_selection_fn=R.continuous_default_policy, max_depth=sc["max_depth"],
                  n_actions=getattr(env.action_space, "n", None), x_values=[], y_values=[], depths=[])
    a = sc["agent"]
    if a in ("Mcts", "MctsHash"):
        param = R.MctsParameters(**common)
    elif a == "MctsApw":
        param = R.PwParameters(**common, alpha=sc["alpha"], k=sc["k"], action_expansion_function=R.continuous_default_policy)
    elif a == "MctsDpw":
        common["action_selection_fn"] = R.ucb1_index
        param = R.DpwParams(**common, alphaApw=sc["alpha1"], kApw=sc["k1"], alphaSpw=sc["alpha2"], kSpw=sc["k2"],
                            state_variable="state")
    else:
        raise KeyError(a)
    return getattr(R, a)(param)


def time_fit(config: str, seed: int = 0, n_sim: int | None = None):
    """One fit() of the reference's own class for BASELINE config `config`; returns (seconds, sims, env_steps)."""
    root = reference_root()
    if root is None:
        raise RuntimeError("reference not staged (run oracle.ref_bench.stage() where /root/reference exists)")
    os.environ["ISLA_REFERENCE_ROOT"] = root
    from oracle import pyenvs, ref_loader
    ref_loader.REFERENCE_ROOT = root
    R = ref_loader.load()
    sc = dict(CONFIGS[config])
    if n_sim:
        sc["n_sim"] = int(n_sim)
    np.random.seed(seed)
    import random
    random.seed(seed)
    agent = build_agent(R, sc, seed)
    steps0 = pyenvs.STEP_COUNTER[0]
    t0 = time.perf_counter()
    try:
        agent.fit()
    except ValueError:
        # mcts_double_...:41 decodes a 4-byte float32 action key with dtype=float (HEAD bug): the search has
        # finished by then, only the returned action is lost
        pass
    dt = time.perf_counter() - t0
    return dt, sc["n_sim"], pyenvs.STEP_COUNTER[0] - steps0


def _worker(args):
    return time_fit(*args)


def time_many(config: str, fits: int, cores: int, n_sim: int | None = None):
    """`fits` independent fit()s on `cores` processes; returns (wall seconds, sims, env_steps)."""
    import multiprocessing as mp
    jobs = [(config, s, n_sim) for s in range(fits)]
    t0 = time.perf_counter()
    if cores == 1:
        res = [_worker(j) for j in jobs]
    else:
        with mp.get_context("fork").Pool(cores) as pool:
            res = pool.map(_worker, jobs, chunksize=1)
    wall = time.perf_counter() - t0
    return wall, sum(r[1] for r in res), sum(r[2] for r in res)


if __name__ == "__main__":
    print("staged:", stage(force=True))
    for c in ("c1", "c2", "c3", "c4"):
        dt, sims, steps = time_fit(c, 0, 2000 if c in ("c3", "c4") else None)
        print(c, "%.3f s  %.0f sims/s  %.0f env-steps/s" % (dt, sims / dt, steps / dt))
