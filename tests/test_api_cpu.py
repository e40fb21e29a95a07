"""Host-side API mirror: what can be checked without a GPU (construction, proxies, error behaviour).
Known-answer values come from the reference's own unit tests (/root/reference/tests/agents/)."""
import numpy as np
import pytest

from islamcts_b200.agents.abstract_mcts import AbstractActionNode, AbstractStateNode, identify_env
from islamcts_b200.agents.parameters.dpw_parameters import DpwParameters
from islamcts_b200.agents.parameters.mcts_parameters import MctsParameters
from islamcts_b200.agents.parameters.pw_parameters import PwParameters
from islamcts_b200.utils.action_selection_functions import continuous_default_policy, genetic_policy, genetic_policy2, ucb1, voo
from islamcts_b200.utils.agent_factory import get_agent


def _params(**kw):
    base = dict(root_data=np.zeros(4, np.float32), env=None, n_sim=10, C=1, action_selection_fn=ucb1, gamma=0.9,
                rollout_selection_fn=continuous_default_policy, max_depth=5, n_actions=2)
    base.update(kw)
    return MctsParameters(**base)


def test_q_value_known_answers():
    # /root/reference/tests/agents/test_abstract_mcts.py:28-40: q = total/na = 10/2; na == 0 -> ZeroDivisionError
    a = AbstractActionNode(data=1, param=None)
    a.total, a.na = 10, 2
    assert a.q_value == 5
    a.na = 0
    with pytest.raises(ZeroDivisionError):
        _ = a.q_value


def test_parameter_records_accept_both_reference_forms():
    p = _params()                                                          # README.md:67-77 form (no x/y/depths)
    assert p.x_values == [] and p.depths == []
    q = PwParameters(root_data=None, env=None, n_sim=1, C=1, action_selection_fn=ucb1, gamma=0.5,
                     rollout_selection_fn=continuous_default_policy, max_depth=5, n_actions=10, alpha=0, k=1,
                     action_expansion_function=continuous_default_policy, x_values=None, y_values=None, depths=None)  # sweep.py:93-110 form
    assert q.alpha == 0 and q.k == 1
    d = DpwParameters(root_data=None, env=None, n_sim=1, C=1, action_selection_fn=ucb1, gamma=0.5,
                      rollout_selection_fn=continuous_default_policy, max_depth=5, n_actions=10, alphaApw=0.5, kApw=1,
                      alphaSpw=0.1, kSpw=1, x_values=None, y_values=None, depths=None, state_variable="state")
    assert (d.alphaApw, d.kApw, d.alphaSpw, d.kSpw) == (0.5, 1, 0.1, 1)
    with pytest.raises(TypeError):
        MctsParameters(root_data=None)                                     # the core fields stay required


def test_agent_factory_dispatch_and_errors():
    # /root/reference/islaMcts/utils/agent_factory.py:16-46
    assert type(get_agent("vanilla", _params(root_data=0))).__name__ == "Mcts"
    assert type(get_agent("vanilla", _params())).__name__ == "MctsHash"
    with pytest.raises(NotImplementedError):
        get_agent("apw", _params(root_data=0))                              # hashable obs -> NotImplementedError
    with pytest.raises(NotImplementedError):
        get_agent("no_such_agent", _params())


def test_unknown_callables_and_envs_fail_loudly():
    from islamcts_b200.agents.mcts_hash import MctsHash

    class FakeCartPole:
        spec_id = "CartPole-v1"
        state = np.zeros(4)
        unwrapped = property(lambda self: self)

    ag = MctsHash(_params(env=FakeCartPole(), rollout_selection_fn=lambda node: 0))
    with pytest.raises(NotImplementedError):
        ag.fit()                                                            # a Python callable cannot run on the device
    with pytest.raises(NotImplementedError):
        identify_env(type("Acrobot", (), {})())
    g1 = genetic_policy(0.7, continuous_default_policy, 5)
    assert (g1.isla_kind, g1.epsilon, g1.n_samples) == ("genetic", 0.7, 5)
    with pytest.raises(NotImplementedError):
        genetic_policy(0.7, lambda node: 0, 5)                              # only the reference's default policy
    g2 = genetic_policy2(0.7, continuous_default_policy)
    assert (g2.isla_kind, g2.epsilon) == ("genetic2", 0.7)
    v = voo(0.7, continuous_default_policy, 5)
    assert (v.isla_kind, v.epsilon, v.n_samples, v.max_try) == ("voo", 0.7, 5, 1000)
    from islamcts_b200.utils.action_selection_functions import grid_policy
    gp = grid_policy({s: [s % 3, 1, 2, 0] for s in range(16)}, 4)
    assert gp.isla_kind == "grid" and gp.table.shape == (16, 4) and gp.table[5, 0] == 2.0
    with pytest.raises(ValueError):
        grid_policy({0: [1, 2, 3, 4], 2: [1, 2, 3, 4]}, 4)


def test_depth_statistics_of_proxies():
    # agents/abstract_mcts.py:120-142,190-209: root -> a0 -> s1 -> a1 -> s2 ; root -> a2 -> s3
    root, s1, s2, s3 = (AbstractStateNode(None, None) for _ in range(4))
    a0, a1, a2 = (AbstractActionNode(i, None) for i in range(3))
    root.actions = {0: a0, 1: a2}; a0.children = {0: s1}; s1.actions = {0: a1}; a1.children = {0: s2}; a2.children = {0: s3}
    assert root.get_depth_max(0) == 3
    assert sorted(root.get_depth_mean(0)) == [2, 3]
    assert a0.get_depth_max(0) == 2 and a0.get_depth_mean(0, True) == 2.0


def test_compat_package_paths():
    from oracle import ref_loader
    ref_loader.unload()
    from islaMcts.agents.mcts_hash import MctsHash
    from islaMcts.agents.mcts_action_progressive_widening_hash import MctsActionProgressiveWideningHash  # noqa: F401
    from islaMcts.agents.parameters.mcts_parameters import MctsParameters as P2
    from islaMcts.utils.action_selection_functions import ucb1 as u2
    import islamcts_b200.agents.mcts_hash as real
    assert MctsHash is real.MctsHash and u2 is ucb1
    # the compat parameter records ARE the package's records, except that code imported unchanged from the reference gets the
    # reference's HEAD semantics (zero-length rollouts in Mcts / APW / SPW / DPW) unless it says otherwise
    assert issubclass(P2, MctsParameters)
    kw = dict(root_data=None, env=None, n_sim=1, C=1, action_selection_fn=ucb1, gamma=1.0, rollout_selection_fn=None, max_depth=1, n_actions=2)
    assert P2(**kw).depth_mode == "reference_head" and MctsParameters(**kw).depth_mode is None
    from islaMcts.agents.parameters.pw_parameters import PwParameters as Pw2
    from islaMcts.agents.parameters.dpw_parameters import DpwParameters as Dpw2
    assert Pw2(**kw, alpha=0.5, k=1, action_expansion_function=None).depth_mode == "reference_head"
    assert Dpw2(**kw, alphaApw=0.5, kApw=1, alphaSpw=0.5, kSpw=1).depth_mode == "reference_head"


def test_curve_env_host_mirror_without_gpu():
    """islaMcts.environments.curve_env import path, spaces, action table and env identification (no device call)."""
    import itertools
    from islaMcts.environments.curve_env import CurveEnv, DiscreteCurveEnv
    from islamcts_b200 import _lib
    from islamcts_b200.agents.abstract_mcts import env_action_info, env_root_state, identify_env
    env = CurveEnv()
    assert env.action_space.shape == (2,) and env.action_space.dtype == np.float64
    np.testing.assert_array_equal(env.action_space.low, [-5.0, -30.0])
    np.testing.assert_array_equal(env.reset(), [-11.0, 21.0, 10.0, 90.0])
    assert identify_env(env) == _lib.ENV_CURVE and env_action_info(env, _lib.ENV_CURVE) == (0, None)
    np.testing.assert_array_equal(env_root_state(env, _lib.ENV_CURVE), [-11.0, 21.0, 10.0, 90.0, 0.0])
    assert env.lower_bound(0.0) == 27.3 and env.higher_bound(0.0) == 27.5
    d = DiscreteCurveEnv([5, 19])
    assert d.action_space.n == 95 and len(d.actions) == 95
    assert d.actions == list(itertools.product(np.linspace(-5, 5, 5), np.linspace(-30, 30, 19)))
    assert env_action_info(d, _lib.ENV_CURVE) == (95, (5, 19))
    assert env_action_info(DiscreteCurveEnv([1, 4]), _lib.ENV_CURVE) == (4, (1, 4))
    assert env_action_info(DiscreteCurveEnv([4, 1]), _lib.ENV_CURVE) == (4, (4, 1))
    big = DiscreteCurveEnv([10, 20])
    with pytest.raises(NotImplementedError):
        env_action_info(big, _lib.ENV_CURVE)
