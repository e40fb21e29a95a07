"""Checks that need the reference tree itself (/root/reference in the build container, or the copy staged under
baseline/_ref): the committed golden fixtures are what the reference produces TODAY, byte for byte, and the reference's
own classes can be timed as the CPU arm.  Skipped where neither exists."""
import os

import numpy as np
import pytest

from oracle import ref_bench

pytestmark = pytest.mark.skipif(ref_bench.reference_root() is None, reason="reference tree not present")

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize("name", ["hash_cartpole", "mcts_lake", "apw_pendulum_c3", "dpw_mountaincar_c4", "spw_cartpole_descend",
                                  "hash_lake_slippery"])
def test_golden_fixture_regenerates_byte_for_byte(name):
    """oracle/gen_golden.py re-run against the live reference: every array of the committed fixture is reproduced exactly
    (decision trace, canonical tree, root statistics, depth metrics, returned action, env-step count)."""
    from oracle import gen_golden, ref_loader
    ref_loader.REFERENCE_ROOT = ref_bench.reference_root()
    fresh = gen_golden.run_scenario(name, gen_golden.SCENARIOS[name])
    stored = np.load(os.path.join(ROOT, "tests", "golden", name + ".npz"), allow_pickle=False)
    assert set(stored.files) == set(fresh)
    for k in stored.files:
        a, b = np.asarray(stored[k]), np.asarray(fresh[k])
        assert a.dtype == b.dtype and a.shape == b.shape, k
        assert a.tobytes() == b.tobytes(), f"{name}: array {k} differs from the committed fixture"


def test_reference_classes_run_as_cpu_arm():
    """bench.py --impl reference times the reference's own MctsHash.fit through this helper."""
    dt, sims, steps = ref_bench.time_fit("c5", seed=0, n_sim=60)
    assert sims == 60 and steps > 60 and dt > 0
    dt, sims, steps = ref_bench.time_fit("c4", seed=0, n_sim=50)       # DPW with the two shims; HEAD's decode bug is caught
    assert sims == 50 and steps == 50
