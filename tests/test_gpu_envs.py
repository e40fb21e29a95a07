"""CUDA env dynamics and rollouts vs the CPU oracle (called through the C ABI)."""
import numpy as np
import pytest

from oracle import coracle

pytestmark = pytest.mark.gpu

ENVS = [0, 1, 2, 3, 4]


def _curve_fixture():
    import os
    from tests.golden_util import GOLDEN_DIR
    return np.load(os.path.join(GOLDEN_DIR, "env", "curve_env_steps.npz"))


def _random_states(env_id, n, rng):
    if env_id == 0:
        s = rng.uniform([-2.3, -2.0, -0.2, -2.5], [2.3, 2.0, 0.2, 2.5], size=(n, 4))
        a = rng.integers(0, 2, n).astype(np.float64)
    elif env_id == 1:
        s = rng.uniform([-12.0, -8.0], [12.0, 8.0], size=(n, 2))
        a = rng.uniform(-2.5, 2.5, n).astype(np.float32).astype(np.float64)
    elif env_id == 2:
        s = rng.uniform([-1.2, -0.07], [0.6, 0.07], size=(n, 2)).astype(np.float32).astype(np.float64)
        a = rng.uniform(-1.2, 1.2, n).astype(np.float32).astype(np.float64)
    elif env_id == 3:
        s = rng.integers(0, 16, size=(n, 1)).astype(np.float64)
        a = rng.integers(0, 4, n).astype(np.float64)
    else:
        # CurveEnv: poses the reference's own env visited (fixture), fresh actions (half of them braking right turns)
        fix = _curve_fixture()["state"]
        s = fix[rng.integers(0, len(fix), n)].copy()
        a = rng.uniform([-5.0, -30.0], [5.0, 30.0], size=(n, 2))
        a[::2] = rng.uniform([-5.0, -30.0], [0.0, 0.0], size=(len(a[::2]), 2))
    return s, a


@pytest.mark.parametrize("env_id", ENVS)
def test_env_step_f64_matches_oracle(env_id):
    from islamcts_b200 import engine as E
    rng = np.random.default_rng(env_id)
    s, a = _random_states(env_id, 4096, rng)
    ns, r, t = E.env_step(env_id, s, a, precision=1)
    ns, r, t = ns.cpu().numpy(), r.cpu().numpy(), t.cpu().numpy()
    for i in range(len(s)):
        ens, er, et, _ = coracle.env_step(env_id, s[i], a[i])
        # same IEEE operations in the same order; CUDA and glibc sin/cos may differ in the last ulp
        np.testing.assert_allclose(ns[i], ens, rtol=1e-13, atol=1e-15)
        assert r[i] == pytest.approx(er, rel=1e-13, abs=1e-15)
        assert bool(t[i]) == et


@pytest.mark.parametrize("env_id", ENVS)
def test_env_step_f32_within_1e5(env_id):
    """north_star: continuous env states and rewards under identical fed-in actions within 1e-5 relative (fp32)."""
    from islamcts_b200 import engine as E
    rng = np.random.default_rng(10 + env_id)
    s, a = _random_states(env_id, 4096, rng)
    s = s.astype(np.float32).astype(np.float64)  # feed values exactly representable in fp32
    ns, r, t = E.env_step(env_id, s, a, precision=0)
    ns, r, t = ns.cpu().numpy().astype(np.float64), r.cpu().numpy().astype(np.float64), t.cpu().numpy()
    mism = 0
    for i in range(len(s)):
        ens, er, et, _ = coracle.env_step(env_id, s[i], a[i])
        if env_id == 3:
            assert ns[i][0] == ens[0] and r[i] == er and bool(t[i]) == et   # FrozenLake: bit-exact
            continue
        scale = np.maximum(np.abs(ens), np.abs(s[i]))  # relative to the state magnitude
        assert np.all(np.abs(ns[i] - ens) <= 1e-5 * scale + 1e-7), (i, ns[i], ens)
        assert abs(r[i] - er) <= 1e-5 * max(1.0, abs(er))
        mism += int(bool(t[i]) != et)
    assert mism <= 2  # a threshold crossing decided differently in fp32 needs a state within 1e-7 of it


@pytest.mark.parametrize("env_id", ENVS)
def test_scripted_trajectories_short_horizon(env_id):
    """Same fed-in action sequence for 25 steps from reset states; fp32 device vs float64 oracle."""
    from islamcts_b200 import engine as E
    import torch
    rng = np.random.default_rng(20 + env_id)
    n, T = 512, 25
    s0 = np.stack([coracle.env_reset(env_id, 7, i) for i in range(n)])
    _, acts = _random_states(env_id, n * T, rng)
    acts = acts.reshape(T, n, -1).squeeze(-1) if acts.ndim == 1 else acts.reshape(T, n, -1)
    dev_s = torch.as_tensor(s0, dtype=torch.float32)
    cpu_s = s0.astype(np.float32).astype(np.float64)
    alive = np.ones(n, bool)
    for k in range(T):
        ns, r, t = E.env_step(env_id, dev_s, acts[k], precision=0)
        dev_s = ns
        got = ns.cpu().numpy().astype(np.float64)
        for i in np.flatnonzero(alive):
            ens, er, et, _ = coracle.env_step(env_id, cpu_s[i], acts[k][i])
            cpu_s[i] = ens
            tol = 1e-5 * np.maximum(1.0, np.abs(ens)) * (1 + k)  # round-off grows along the trajectory
            assert np.all(np.abs(got[i] - ens) <= tol), (env_id, k, i, got[i], ens)
            if et or bool(t[i].item()):
                alive[i] = False


@pytest.mark.parametrize("env_id,budget,gamma,grid", [(0, 500, 0.99, None), (1, 200, 0.99, None), (2, 500, 0.99, None),
                                                      (3, 1000, 1.0, None), (0, 7, 0.9, None), (4, 40, 0.95, None),
                                                      (4, 40, 0.95, (5, 7)), (4, 12, 0.9, (2, 2)), (4, 12, 0.9, (1, 2))])
def test_rollout_f64_matches_oracle_streams(env_id, budget, gamma, grid):
    from islamcts_b200 import engine as E
    n = 1500
    if env_id == 4:
        s0 = _random_states(4, n, np.random.default_rng(4))[0]
    else:
        s0 = np.stack([coracle.env_reset(env_id, 3, i) for i in range(n)])
    ret, steps = E.rollout(env_id, s0, budget=budget, gamma=gamma, seed=99, sim=5, precision=1, action_grid=grid)
    ret, steps = ret.cpu().numpy(), steps.cpu().numpy()
    if env_id == 4 and budget == 40:
        assert steps.max() > 3 and (steps < budget).mean() > 0.5   # the car does leave the track in most rollouts
    for i in range(n):
        eret, esteps = coracle.rollout(env_id, s0[i], budget, gamma, 99, i, 5, 0, grid=grid or (0, 0))
        assert steps[i] == esteps, (i, steps[i], esteps)
        assert ret[i] == pytest.approx(eret, rel=1e-11, abs=1e-11)


@pytest.mark.parametrize("env_id", [0, 3])
@pytest.mark.parametrize("n,budget", [(1, 5), (31, 0), (33, 1), (64, 3), (2049, 500)])
def test_rollout_queue_edges(env_id, n, budget):
    """The persistent kernels' start queues at their edges: fewer trajectories than one production (32), a ragged last
    production, budget 0 (a queued start finishes without a step) and budgets shorter than the warp's meeting period."""
    from islamcts_b200 import engine as E
    s0 = np.stack([coracle.env_reset(env_id, 8, i) for i in range(n)])
    for precision in (1, 0):
        ret, steps = E.rollout(env_id, s0, budget=budget, gamma=0.97, seed=123, sim=2, precision=precision)
        ret, steps = ret.cpu().numpy(), steps.cpu().numpy()
        for i in range(n):
            eret, esteps = coracle.rollout(env_id, s0[i], budget, 0.97, 123, i, 2, 0)
            if precision == 1 or env_id == 3:
                assert steps[i] == esteps and ret[i] == pytest.approx(eret, rel=1e-11, abs=1e-11), (i, steps[i], esteps)
            else:
                assert abs(int(steps[i]) - esteps) <= (1 if budget > 3 else 0)    # fp32 may cross the threshold a step apart


@pytest.mark.parametrize("env_id", [1, 2])
@pytest.mark.parametrize("budget", [0, 1, 2, 3, 5, 7, 202])
def test_rollout_four_step_blocks_at_their_edges(env_id, budget):
    """Pendulum / MountainCar rollouts run four predicated steps per Philox block (rollout_fast1d): budgets that end inside a
    block, and MountainCar episodes that reach the goal at every offset within a block, against the step-by-step oracle."""
    from islamcts_b200 import engine as E
    n = 257
    rng = np.random.default_rng(31 + env_id)
    if env_id == 2:
        # near the goal (0.45) and moving right: most of these terminate within the first 1-12 steps
        s0 = np.stack([rng.uniform(0.30, 0.449, n), rng.uniform(0.0, 0.06, n)], axis=1)
    else:
        s0 = np.stack([coracle.env_reset(env_id, 3, i) for i in range(n)])
    ret, steps = E.rollout(env_id, s0, budget=budget, gamma=0.97, seed=77, sim=4, precision=1)
    ret, steps = ret.cpu().numpy(), steps.cpu().numpy()
    for i in range(n):
        eret, esteps = coracle.rollout(env_id, s0[i], budget, 0.97, 77, i, 4, 0)
        assert steps[i] == esteps, (i, steps[i], esteps)
        assert ret[i] == eret or ret[i] == pytest.approx(eret, rel=1e-11, abs=1e-11)
    if env_id == 2 and budget >= 7:
        ended = steps[steps < budget]
        assert len(ended) > 20 and len(set((ended % 4).tolist())) == 4   # goals reached at every offset within a block


def test_rollout_reset_distribution_and_f32_agreement():
    from islamcts_b200 import engine as E
    n = 20000
    ret64, st64 = E.rollout(0, None, n=n, budget=500, gamma=0.99, seed=5, sim=0, precision=1)
    ret32, st32 = E.rollout(0, None, n=n, budget=500, gamma=0.99, seed=5, sim=0, precision=0)
    st64, st32 = st64.cpu().numpy(), st32.cpu().numpy()
    # device-side reset == oracle reset (first few) and fp32 rollouts end on the same step almost always
    for i in range(64):
        eret, esteps = coracle.rollout(0, coracle.env_reset(0, 5, i), 500, 0.99, 5, i, 0, 0)
        assert st64[i] == esteps
    assert (st64 != st32).mean() < 2e-3
    assert 15 < st64.mean() < 30   # random CartPole episodes last ~22 steps


@pytest.mark.parametrize("precision", [1, 0])
def test_curve_env_matches_reference_step_vectors(precision):
    """CurveEnv is reference code (environments/curve_env.py:121-177): the device env against step vectors recorded
    from the reference's own class (tests/golden/env/curve_env_steps.npz).  float64: terminal flags bit-exact, reals to
    1e-12; float32: 1e-5 relative, and a containment test decided differently needs a pose within ~1e-5 of the edge."""
    from islamcts_b200 import engine as E
    z = _curve_fixture()
    ns, r, t = E.env_step(4, z["state"], z["action"], precision=precision)
    ns, r, t = ns.cpu().numpy().astype(np.float64), r.cpu().numpy().astype(np.float64), t.cpu().numpy().astype(bool)
    if precision == 1:
        np.testing.assert_array_equal(t, z["terminal"])
        np.testing.assert_allclose(ns, z["next_state"], rtol=1e-12, atol=1e-12)
        np.testing.assert_allclose(r, z["reward"], rtol=1e-12)
    else:
        same = t == z["terminal"]
        assert (~same).sum() <= 2
        np.testing.assert_allclose(ns, z["next_state"], rtol=1e-5, atol=1e-5)
        np.testing.assert_allclose(r[same], z["reward"][same], rtol=1e-5)


def test_curve_containment_shortcut_equals_100_point_test():
    """The device decides containment from <= 6 of the reference's 100 linspace points (+ full loop when a margin is
    within 1e-9); on 200k random moves its verdict must equal the oracle's plain 100-point loop."""
    from islamcts_b200 import engine as E
    rng = np.random.default_rng(123)
    n = 200_000
    x = rng.uniform(-12.0, 10.0, n)
    lo, hi = -(1 / 50) * (3 * x ** 2) + 27.3, -(1 / 50) * x ** 2 + 27.5
    y = rng.uniform(lo - 0.2, hi + 0.2)
    s = np.stack([x, y, rng.uniform(0, 20, n), rng.uniform(0, 360, n), rng.integers(0, 40, n).astype(np.float64)], 1)
    a = rng.uniform([-5.0, -30.0], [5.0, 30.0], size=(n, 2))
    ns, r, t = E.env_step(4, s, a, precision=1)
    r, t = r.cpu().numpy(), t.cpu().numpy().astype(bool)
    sub = rng.choice(n, 20000, replace=False)
    for i in sub:
        _, er, et, _ = coracle.env_step(4, s[i], a[i])
        assert bool(t[i]) == et and r[i] == pytest.approx(er, rel=1e-12), (i, s[i], a[i])
    assert 0.05 < t.mean() < 0.95


@pytest.mark.parametrize("env_id,budget,gamma,tol", [(0, 500, 0.99, 1e-12), (1, 25, 0.99, 1e-4), (2, 25, 0.99, 1e-5),
                                                      (1, 200, 0.99, 5e-2), (2, 500, 0.99, 1e-4), (3, 1000, 1.0, 0.0)])
def test_rollout_returns_f32_vs_f64(env_id, budget, gamma, tol):
    """The float32 rollout path (what the benchmarks run) against the float64 one on the same Philox action streams:
    discounted returns on every trajectory that ends on the same step.  A single float32 step is within the north_star
    tolerance (1e-5 relative, test_env_step_*); a RETURN sums a trajectory along which the float32 state error grows
    (Pendulum near its unstable equilibrium: measured worst case 2.9e-5 after 25 steps), hence 1e-4 over 25 Pendulum
    steps, 1e-5 over 25 MountainCar steps; over the full C3 horizon (200 steps of a driven pendulum: trajectories that pass
    the unstable upright position amplify the float32 rounding, measured worst case 1.8e-2 of 8192) the bound is 5e-2 and
    the MEDIAN deviation must stay below 1e-5; 1e-4 over the C4 horizon.  CartPole's closed-form return and FrozenLake's
    integers agree exactly."""
    from islamcts_b200 import engine as E
    n = 8192
    ret64, st64 = E.rollout(env_id, None, n=n, budget=budget, gamma=gamma, seed=11, sim=3, precision=1)
    ret32, st32 = E.rollout(env_id, None, n=n, budget=budget, gamma=gamma, seed=11, sim=3, precision=0)
    same = (st64 == st32).cpu().numpy()
    assert same.mean() > 0.995
    r64, r32 = ret64.cpu().numpy()[same], ret32.cpu().numpy()[same]
    scale = np.maximum(np.abs(r64), 1.0)
    dev = np.abs(r32 - r64) / scale
    assert dev.max() <= tol, (dev.max(), np.argmax(dev))
    assert np.median(dev) <= 1e-5
    # and the float64 path is the oracle's (steps exact, returns 1e-11) on a few trajectories of this configuration
    for i in range(8):
        eret, esteps = coracle.rollout(env_id, coracle.env_reset(env_id, 11, i), budget, gamma, 11, i, 3, 0)
        assert int(st64[i]) == esteps and float(ret64[i]) == pytest.approx(eret, rel=1e-11, abs=1e-11)
