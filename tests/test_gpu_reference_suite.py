"""The reference's own test file (tests/test-envs.py:1-139) replayed against this package: every registered id is made,
checked for gym-API conformance (what stable-baselines3's check_env verifies, which is not installed here), driven by the
msy and escapement policies through env.simulate, and the two numeric assertions of test_tipping are kept verbatim."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def check_env(env):
    """The checks of stable_baselines3.common.env_checker.check_env that apply to an old-gym (4-tuple) environment."""
    from gym_fishing_b200.spaces import Box, Discrete

    assert isinstance(env.observation_space, Box) and env.observation_space.shape == (1,)
    assert isinstance(env.action_space, (Box, Discrete))
    obs = env.reset()
    assert isinstance(obs, np.ndarray) and obs.shape == (1,) and np.isfinite(obs).all()
    for _ in range(3):
        action = env.action_space.sample()
        out = env.step(action)
        assert isinstance(out, tuple) and len(out) == 4
        obs, reward, done, info = out
        assert isinstance(obs, np.ndarray) and obs.shape == (1,)
        assert isinstance(reward, (float, int, np.floating)) and isinstance(done, bool) and isinstance(info, dict)
    env.reset()


def drive(env, output_prefix):
    import gym_fishing_b200 as gf

    for policy in (gf.msy, gf.escapement):
        model = policy(env)
        df = env.simulate(model)
        assert list(df.columns) == ["time", "state", "action", "reward", "rep"] and 1 <= len(df) <= env.Tmax
        assert np.isfinite(df.state.values).all() and (df.reward.values >= 0).all()
        with pytest.raises(RuntimeError, match="matplotlib"):       # plotting needs matplotlib, which is not installed
            env.plot(df, output_prefix + "-test.png")
    return df


def test_discrete():
    import gym_fishing_b200 as gf

    np.random.seed(0)
    env = gf.make("fishing-v0")
    check_env(env)
    drive(env, "v0")


def test_cts():
    import gym_fishing_b200 as gf

    env = gf.make("fishing-v1")
    check_env(env)
    drive(env, "v1")


@pytest.mark.parametrize("env_id", ["fishing-v5", "fishing-v6", "fishing-v7", "fishing-v8", "fishing-v9"])
def test_lognormal_growth_models(env_id):
    import gym_fishing_b200 as gf

    env = gf.make(env_id, sigma=0)
    check_env(env)
    df = drive(env, env_id)
    assert df.reward.sum() > 0


def test_tipping():
    import gym_fishing_b200 as gf

    np.random.seed(0)
    env = gf.make("fishing-v2", sigma=0, init_state=0.75)
    check_env(env)
    env.reset()
    # increases above tipping point
    obs, reward, done, info = env.step(env.get_action(0))
    assert env.get_fish_population(obs) >= 0.75
    # decreases below the tipping point
    env.init_state = 0.3
    env.reset()
    obs, reward, done, info = env.step(env.get_action(0))
    assert env.get_fish_population(obs) <= 0.3
    drive(env, "tip")


def test_nonstationary():
    import gym_fishing_b200 as gf

    env = gf.make("fishing-v10", sigma=0, alpha=-0.007)
    check_env(env)
    r0 = env.params["r"]
    drive(env, "ns")
    assert env.params["r"] < r0          # r keeps drifting and is never reset (growth_models.py:147-150)


def test_model_uncertainty():
    import gym_fishing_b200 as gf

    np.random.seed(0)
    env = gf.make("fishing-v11")
    check_env(env)
    for policy in (gf.msy, gf.escapement):
        df = env.simulate(policy(env), reps=10)
        assert df.rep.max() == 9
    assert env.model in env.models


def test_model_error_v4_which_upstream_never_tests():
    import gym_fishing_b200 as gf

    np.random.seed(1)
    env = gf.make("fishing-v4")
    check_env(env)
    K1 = env.K
    obs = env.reset()
    assert obs[0] == 0.75 and env.K != K1      # un-normalised reset state, parameters redrawn (fishing_model_error.py:41-48)
    drive(env, "v4")
