#!/usr/bin/env python
"""The reference's own RL script (tests/test-PPO.py:9-21, examples/PPO.py of boettiger-lab/gym_fishing) with one changed line:
the environment is the B200 batch instead of a single gym env.  Needs stable-baselines3 (NOT in the build image -- this file is
documentation of the drop-in, `tests/test_sb3_dropin.py` checks the same contract against a stand-in package):

    from stable_baselines3 import PPO
    from stable_baselines3.common.evaluation import evaluate_policy

FishingVecEnv subclasses stable_baselines3.common.vec_env.VecEnv, so PPO takes it as its vectorised environment as it is
(no DummyVecEnv around it); io="numpy" hands the learner numpy arrays, sb3_dtypes=True the float32 observations / rewards it
would get from DummyVecEnv anyway; monitor=True adds VecMonitor's infos[i]["episode"] from inside the step kernel."""
import sys

import gym_fishing_b200 as gf


def main(argv=None):
    try:
        from stable_baselines3 import PPO
        from stable_baselines3.common.evaluation import evaluate_policy
    except ImportError:
        print("stable-baselines3 is not installed: nothing to run (see the module docstring)")
        return None
    n_envs = 4096
    env = gf.FishingVecEnv("fishing-v1", num_envs=n_envs, io="numpy", sb3_dtypes=True, sigma=0.1, seed=0)
    model = PPO("MlpPolicy", env, n_steps=64, batch_size=n_envs * 8, verbose=0)      # tests/test-PPO.py:12
    model.learn(total_timesteps=n_envs * 64 * 20)                                   # :13
    eval_env = gf.FishingVecEnv("fishing-v1", num_envs=256, io="numpy", sb3_dtypes=True, sigma=0.1, seed=1)
    mean_reward, std_reward = evaluate_policy(model, eval_env, n_eval_episodes=256)  # :20
    print(f"mean episode reward {mean_reward:.3f} +- {std_reward:.3f}")
    # the reference's own VecEnv consumer: simulate_mdp_vec (gym_fishing/envs/shared_env.py:57-79)
    df = gf.shared_env.simulate_mdp_vec(eval_env, model, 256)
    print(df.groupby("rep").reward.sum().describe())
    return mean_reward


if __name__ == "__main__":
    main(sys.argv[1:])
