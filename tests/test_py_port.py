"""The pure-Python port (CPU baseline) makes the reference's exact draws and builds its exact tree."""
import numpy as np
import pytest

from oracle import py_port, pyenvs
from oracle import trace as T
from tests.golden_util import load_golden


@pytest.mark.parametrize("name", ["hash_cartpole", "hash_cartpole_shallow", "mcts_lake"])
def test_port_reproduces_reference_run(name):
    g = load_golden(name)
    sc = g["config"]
    hash_keys = sc["agent"] == "MctsHash"
    env = pyenvs.make(sc["env"], api=4 if hash_keys else 5, seed=sc["env_seed"])
    obs = env.reset()
    if not hash_keys:
        obs = obs[0]
    tr = T.Trace()
    env.action_space = T.RecordingSpace(env.action_space, tr)
    port = py_port.VanillaPort(env, obs, sc["n_sim"], sc["C"], sc["gamma"], sc["max_depth"], sc["n_actions"], hash_keys)
    if sc["budget"] == "zero":
        pass  # Mcts at HEAD: depth starts at max_depth and is passed through unchanged -> zero-length rollouts
    with T.recording(tr, sc["rng_seed"]):
        action = port.fit()
    kinds, vals = tr.arrays()
    np.testing.assert_array_equal(kinds, g["trace_kind"])      # same sequence of random calls ...
    np.testing.assert_array_equal(vals, g["trace_value"])      # ... with the same outcomes
    ser = T.serialise_reference_tree(port.root)
    for k, v in ser.items():
        if v.dtype.kind == "f":
            np.testing.assert_allclose(v, g["tree_" + k], rtol=1e-12, atol=1e-12)
        else:
            np.testing.assert_array_equal(v, g["tree_" + k])
    assert float(action) == float(g["fit_action"])
