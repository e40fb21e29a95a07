"""The C oracle replays the reference's decision traces and must rebuild the reference's trees.

Golden fixtures: tests/golden/*.npz, produced by oracle/gen_golden.py from the reference's own
agents (/root/reference/islaMcts/agents/*.py, unmodified).  Integers bit-exact; float64 totals to
1e-9 (the oracle and CPython do the same IEEE operations; libm pow/log may differ in the last ulp).
"""
import numpy as np
import pytest

from oracle import coracle
from tests.golden_util import assert_tree_equal, engine_kwargs, golden_names, load_golden


@pytest.mark.parametrize("name", golden_names())
def test_oracle_replays_reference_trace(name):
    g = load_golden(name)
    cfg = g["config"]
    kw = engine_kwargs(cfg)
    conf = coracle.make_config(kw.pop("env_id"), C_=kw.pop("C"), **kw)
    tree = coracle.Tree(conf, g["root_state"])
    if cfg.get("rollout") == "grid":
        from tests.golden_util import lake_manhattan_table
        assert tree.set_heuristic(lake_manhattan_table()) == 0
    if "noise" in g:                                    # slippery FrozenLake: the cloned env's per-step draws
        assert tree.set_noise(g["noise"]) == 0
    kinds, vals = g["trace_kind"], g["trace_value"]
    rc = tree.run(cfg["n_sim"], (kinds, vals))
    assert rc == 0, f"trace fault {rc} at {tree.trace_pos()}/{len(kinds)}"
    # everything but the final arg-max tie-break pick is consumed by the simulations
    assert tree.trace_pos() == len(kinds) - 1
    # the end of fit(): arg-max pick (and, for spw/dpw, the persistent re-ordering of root.actions)
    rc, rank, act = tree.best()
    assert rc == 0 and tree.trace_pos() == len(kinds)
    assert act == pytest.approx(float(g["fit_action"]), abs=1e-12)
    assert_tree_equal(tree.serialise(), g)
    a, na, tot = tree.root_children()
    np.testing.assert_array_equal(na, g["root_na"])
    np.testing.assert_allclose(tot, g["root_total"], rtol=1e-9, atol=1e-9)
    np.testing.assert_allclose(a, g["root_action"], rtol=0, atol=1e-12)
    if "root_action2" in g:
        np.testing.assert_allclose(tree.root_action2, g["root_action2"], rtol=0, atol=1e-12)
    # the sweep loop's depth metrics, computed by the reference's own get_depth_max / get_depth_mean
    dmax, dmean = tree.depth_stats()
    np.testing.assert_array_equal(dmax, g["root_depth_max"])
    np.testing.assert_allclose(dmean, g["root_depth_mean"], rtol=1e-12)
    # counted env.step calls of the reference run (tree-level replay + rollouts); the reference's own
    # CurveEnv is not instrumented (-1)
    if int(g["env_steps"]) >= 0:
        assert tree.counters()["env_steps"] == int(g["env_steps"])
    # the invariants SURVEY.md 8(a) lists for the reference's trees
    assert tree.serialise()["count"][0] == cfg["n_sim"]          # root ns
    if cfg["agent"] in ("Mcts", "MctsHash", "MctsApw"):
        assert na.sum() == cfg["n_sim"]                          # spw/dpw skip na on the descend branch


def test_known_answer_values_of_reference_unit_tests():
    """Known-answer arithmetic of the reference's own unit tests
    (/root/reference/tests/agents/test_mcts_hash.py:149-223, test_abstract_mcts.py:28-66):
    gamma=0.9, r=5: terminal backup -> total 5, na 1, child ns 1; q = total/na."""
    # FrozenLake from the cell left of the goal (s=14): RIGHT reaches G with r=1, terminal.
    conf = coracle.make_config(coracle.ENV_FROZENLAKE, n_actions=4, max_depth=10, C_=0.5, gamma=0.9)
    tree = coracle.Tree(conf, [14.0])
    kinds = np.array([0], np.uint8)      # unvisited pick: position 2 of [0,1,2,3] = RIGHT
    vals = np.array([2.0])
    assert tree.run(1, (kinds, vals)) == 0
    ser = tree.serialise()
    # root(ns=1,total=1) -> action RIGHT(na=1,total=1) -> terminal state(ns=1,total=0)
    assert ser["count"].tolist() == [1, 1, 1]
    assert ser["total"].tolist() == [1.0, 1.0, 0.0]
    assert ser["terminal"].tolist() == [0, 0, 1]
    # rollout returns 0 when no steps are allowed (test_abstract_mcts.py:44-66): budget zero mode
    conf = coracle.make_config(coracle.ENV_CARTPOLE, n_actions=2, max_depth=5, budget_mode=1, C_=1.0, gamma=0.9)
    tree = coracle.Tree(conf, [0.0, 0.0, 0.0, 0.0])
    assert tree.run(1, (np.array([0], np.uint8), np.array([1.0]))) == 0
    ser = tree.serialise()
    assert ser["total"].tolist() == [1.0, 1.0, 1.0]   # G = r + gamma*0
    assert tree.counters()["rollout_steps"] == 0


def test_oracle_curve_env_matches_reference_step_vectors():
    """The reference's own CurveEnv (environments/curve_env.py:121-177) stepped on 1090 random
    (state, action) pairs; fixture tests/golden/env/curve_env_steps.npz (oracle/gen_golden.py).
    State = (x, y, velocity, angle, n_step).  Terminal flags bit-exact, floats to 1e-12."""
    import os
    from tests.golden_util import GOLDEN_DIR
    z = np.load(os.path.join(GOLDEN_DIR, "env", "curve_env_steps.npz"))
    for s, a, ns, r, t in zip(z["state"], z["action"], z["next_state"], z["reward"], z["terminal"]):
        got_ns, got_r, got_t, _ = coracle.env_step(coracle.ENV_CURVE, s, a)
        assert got_t == bool(t)
        np.testing.assert_allclose(got_ns, ns, rtol=1e-12, atol=1e-12)
        assert got_r == pytest.approx(float(r), rel=1e-12)
