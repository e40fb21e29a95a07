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


@pytest.mark.parametrize("env_id,n_actions,prec", [(0, 2, 1), (0, 2, 0), (3, 4, 1)])
def test_reroot_keeps_the_chosen_subtree_bit_for_bit(env_id, n_actions, prec):
    """Tree reuse (opt-in build extension): isla_engine_reroot turns the subtree below the chosen root
    action into the tree.  Its DFS serialisation must be the old tree's subtree exactly -- every count, total, flag and
    action, one level up -- and the search continues on it: the root's visits are the kept ones plus the new simulations."""
    import torch
    from islamcts_b200.engine import Engine
    n_trees, n_sim = 12, 260
    roots = np.stack([coracle.env_reset(env_id, 4, i) for i in range(n_trees)])
    eng = Engine(env_id, n_trees=n_trees, sim_capacity=2 * n_sim + 2, max_depth=60, C_=1.0 if env_id == 0 else 0.5, gamma=0.97,
                 precision=prec, kernel=1, seed=8)
    eng.set_roots(roots).run(n_sim)
    old = [eng.export_tree(i) for i in range(n_trees)]
    _, act = eng.best()
    actions = act.to(torch.int32)
    actions[1] = (actions[1] + 1) % n_actions            # not only arg-max edges
    eng.reroot(actions, n_sim)
    assert eng.counters()["faults"] == 0
    kept = []
    for i in range(n_trees):
        o, n = old[i], eng.export_tree(i)
        a = int(actions[i])
        # the old tree's records below root action a: from its action record (exclusive) to the next depth-0 action record
        top = np.flatnonzero((o["kind"] == 1) & (o["depth"] == 0))
        j = top[list(o["action"][top]).index(float(a))]
        nxt = top[top > j]
        lo, hi = j + 1, (nxt[0] if len(nxt) else len(o["kind"]))
        if hi - lo <= 1 or o["terminal"][lo]:            # a leaf / terminal child: the new tree is a bare root
            assert len(n["kind"]) == 1 and n["count"][0] == 0
            kept.append(0)
            continue
        for k in ("kind", "terminal", "fanout"):
            np.testing.assert_array_equal(n[k][1:], o[k][lo + 1:hi])
        np.testing.assert_array_equal(n["depth"][1:], o["depth"][lo + 1:hi] - 1)
        np.testing.assert_array_equal(n["count"][1:], o["count"][lo + 1:hi])
        assert np.array_equal(n["total"][1:], o["total"][lo + 1:hi])           # bit for bit
        np.testing.assert_array_equal(n["action"][1:], o["action"][lo + 1:hi])
        assert n["count"][0] == o["count"][lo] - 1       # the re-rooted node is a root now: ns = sum of its edges' visits
        kept.append(int(n["count"][0]))
    eng.run(n_sim)
    assert eng.counters()["faults"] == 0
    na, _ = eng.root_stats()
    np.testing.assert_array_equal(na.sum(dim=1).cpu().numpy(), np.array(kept) + n_sim)


def test_episodes_with_tree_reuse_run_to_the_end():
    from islamcts_b200.episodes import EpisodeBatch
    n, n_sim, max_steps = 24, 200, 12
    states = np.stack([coracle.env_reset(0, 900, i) for i in range(n)])
    out = {}
    for reuse in (False, True):
        eb = EpisodeBatch("CartPole-v1", n_episodes=n, n_sim=n_sim, max_steps=max_steps, seed=5, max_depth=60, C_=1.0, gamma=0.97,
                          reuse_tree=reuse, timeout_reward=None)
        out[reuse] = eb.run(states)
        assert (out[reuse]["total_reward"] == out[reuse]["n_steps"]).all()      # CartPole pays 1 per real step
    # kept subtrees cannot make the balancing worse than chance; both modes keep most episodes alive over 12 steps
    assert out[True]["n_steps"].mean() >= 0.8 * out[False]["n_steps"].mean()
