"""Helpers shared by the oracle and GPU parity tests."""
import glob
import json
import os

import numpy as np

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")

ENV_IDS = {"CartPole-v1": 0, "Pendulum-v1": 1, "MountainCarContinuous-v0": 2, "FrozenLake-v1": 3,
           "CurveEnv": 4, "DiscreteCurveEnv": 4}
AGENTS = {"Mcts": 0, "MctsHash": 0, "MctsApw": 1, "MctsSpw": 2, "MctsDpw": 3}


def golden_names():
    return sorted(os.path.splitext(os.path.basename(p))[0] for p in glob.glob(os.path.join(GOLDEN_DIR, "*.npz")))


def load_golden(name):
    z = np.load(os.path.join(GOLDEN_DIR, name + ".npz"))
    g = {k: z[k] for k in z.files}
    g["config"] = json.loads(str(g["config"]))
    return g


def engine_kwargs(cfg):
    """Golden scenario -> the (env, agent, widening, ...) settings both engines take."""
    agent = AGENTS[cfg["agent"]]
    kw = dict(env_id=ENV_IDS[cfg["env"]], agent=agent, n_actions=cfg.get("n_actions") or 0,
              max_depth=cfg["max_depth"], budget_mode=0 if cfg["budget"] == "to_max_depth" else 1,
              C=cfg["C"], gamma=cfg["gamma"])
    if agent in (1, 2):
        kw.update(alpha1=cfg["alpha"], k1=cfg["k"])
    if agent == 3:
        kw.update(alpha1=cfg["alpha1"], k1=cfg["k1"], alpha2=cfg["alpha2"], k2=cfg["k2"])
    if cfg.get("grid"):
        kw.update(grid0=cfg["grid"][0], grid1=cfg["grid"][1], n_actions=cfg["grid"][0] * cfg["grid"][1])
    if cfg.get("expansion") == "voo":
        kw.update(expansion=1, epsilon=cfg["epsilon"], voo_n_samples=cfg["voo_n_samples"],
                  voo_max_try=cfg["voo_max_try"])
    if cfg.get("expansion") == "genetic2":
        kw.update(expansion=2, epsilon=cfg["epsilon"])
    if cfg.get("expansion") == "genetic":
        kw.update(expansion=3, epsilon=cfg["epsilon"], voo_n_samples=cfg["voo_n_samples"])
    return kw


def lake_manhattan_table():
    """[16, 4] prior knowledge of examples/example_lake.py:22-66 (Manhattan distance to the goal after the move)."""
    def dist(r, c, a):
        nr, nc = (r, c - 1) if a == 0 else (r + 1, c) if a == 1 else (r, c + 1) if a == 2 else (r - 1, c)
        if nc < 0 or nc > 3 or nr < 0 or nr > 3:
            nr, nc = r, c
        return abs(3 - nc) + abs(3 - nr)
    return np.array([[dist(s // 4, s % 4, a) for a in range(4)] for s in range(16)], dtype=np.float64)


def assert_tree_equal(got, g, rtol=1e-9, atol=1e-9, label=""):
    """Canonical DFS serialisations: integers bit-exact, totals/actions to tolerance."""
    for k in ("kind", "depth", "count", "terminal", "fanout"):
        exp = g["tree_" + k]
        assert len(got[k]) == len(exp), f"{label}{k}: {len(got[k])} records vs {len(exp)}"
        bad = np.flatnonzero(np.asarray(got[k]).astype(np.int64) != exp.astype(np.int64))
        assert bad.size == 0, f"{label}{k} differs at records {bad[:8]}: {np.asarray(got[k])[bad[:8]]} vs {exp[bad[:8]]}"
    np.testing.assert_allclose(got["total"], g["tree_total"], rtol=rtol, atol=atol, err_msg=label + "total")
    np.testing.assert_allclose(got["action"], g["tree_action"], rtol=0, atol=1e-12, err_msg=label + "action")
    if "tree_action2" in g and "action2" in got:
        np.testing.assert_allclose(got["action2"], g["tree_action2"], rtol=0, atol=1e-12, err_msg=label + "action2")
