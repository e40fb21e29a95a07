"""Episode driver on the device (fit -> real step -> new root for many episodes at once) vs the same loop on the oracle."""
import numpy as np
import pytest

from oracle import coracle

pytestmark = pytest.mark.gpu


def _oracle_episodes(env_id, agent, states, n_sim, max_steps, seed, grid=(0, 0), timeout_reward=-1000.0, **cfg_kw):
    n = len(states)
    total, steps = np.zeros(n), np.zeros(n, np.int32)
    states = np.array(states, dtype=np.float64)
    for e in range(n):
        s = states[e].copy()
        for k in range(max_steps):
            t = coracle.Tree(coracle.make_config(env_id, agent=agent, seed=seed + k, tree_id=e, grid0=grid[0], grid1=grid[1], **cfg_kw), s)
            assert t.run(n_sim) == 0
            _, _, act = t.best()
            a = [act, t.best_action2]
            if grid[0]:   # DiscreteCurveEnv index -> (acceleration, steering)
                a = [np.linspace(-5, 5, grid[0])[int(act) // grid[1]], np.linspace(-30, 30, grid[1])[int(act) % grid[1]]]
            s, r, term, _ = coracle.env_step(env_id, s, a if env_id == 4 else act)
            steps[e] += 1
            if steps[e] >= max_steps and timeout_reward is not None:
                r = timeout_reward                      # sweep.py:206-209 of the reference
            total[e] += r
            if term:
                break
        states[e] = s
    return total, steps, states


@pytest.mark.parametrize("kernel", [1, 2])
def test_cartpole_episodes_equal_oracle_loop(kernel):
    from islamcts_b200.episodes import EpisodeBatch
    n, n_sim, max_steps, seed = 40, 120, 10, 900
    # near the failure boundary, so that some episodes end inside the horizon
    states = np.stack([coracle.env_reset(0, 3, i) for i in range(n)]) + np.array([0.0, 0.0, 0.17, 0.8])
    eb = EpisodeBatch("CartPole-v1", n_episodes=n, n_sim=n_sim, max_steps=max_steps, seed=seed, max_depth=40, C_=1.0,
                      gamma=0.95, precision=1, kernel=kernel)
    out = eb.run(states, check_every=4)
    etot, esteps, estates = _oracle_episodes(0, 0, states, n_sim, max_steps, seed, n_actions=2, max_depth=40, C_=1.0, gamma=0.95)
    np.testing.assert_array_equal(out["n_steps"], esteps)
    np.testing.assert_array_equal(out["total_reward"], etot)
    np.testing.assert_allclose(out["states"], estates, rtol=1e-12, atol=1e-12)
    assert 0 < (esteps < max_steps).sum() < n          # a mix of finished and running episodes


def test_curve_episodes_apw_box_and_vanilla_grid():
    from islamcts_b200.episodes import EpisodeBatch
    base = np.array([-5.70993649, 25.49759526, 2.5, 0.0, 3.0])
    states = np.stack([base + np.array([0.02 * i, 0.0, 0.1 * i, 0.0, 0.0]) for i in range(6)])
    eb = EpisodeBatch("CurveEnv", agent=1, n_episodes=6, n_sim=150, max_steps=6, seed=5, max_depth=20, C_=11.0, gamma=0.9,
                      alpha1=0.5, k1=3.0, precision=1)
    out = eb.run(states)
    etot, esteps, estates = _oracle_episodes(4, 1, states, 150, 6, 5, max_depth=20, C_=11.0, gamma=0.9, alpha1=0.5, k1=3.0)
    np.testing.assert_array_equal(out["n_steps"], esteps)
    np.testing.assert_allclose(out["total_reward"], etot, rtol=1e-12)
    np.testing.assert_allclose(out["states"], estates, rtol=1e-12, atol=1e-12)
    assert out["actions"].shape[1:] == (6, 2)
    eg = EpisodeBatch("DiscreteCurveEnv", agent=0, n_episodes=6, n_sim=150, max_steps=6, seed=5, max_depth=20, C_=11.0,
                      gamma=0.9, precision=1, action_grid=(5, 7))
    out = eg.run(states)
    etot, esteps, estates = _oracle_episodes(4, 0, states, 150, 6, 5, grid=(5, 7), n_actions=35, max_depth=20, C_=11.0, gamma=0.9)
    np.testing.assert_array_equal(out["n_steps"], esteps)
    np.testing.assert_allclose(out["total_reward"], etot, rtol=1e-12)
    np.testing.assert_allclose(out["states"], estates, rtol=1e-12, atol=1e-12)
