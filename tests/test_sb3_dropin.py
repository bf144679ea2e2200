"""Row f3 of SURVEY.md section 8: FishingVecEnv as a stable-baselines3 / gym drop-in.  Neither package is in the image, so
the checks run against stand-ins on sys.path (tests/stubs/stable_baselines3: the abstract VecEnv interface;
oracle/refstub/gym: the registry and spaces the reference itself imports), in a subprocess so that the stand-ins never leak
into the other tests.  What the reference does with these packages: tests/test-PPO.py:9-21, tests/test-DQN.py:9-24
(PPO / DQN("MlpPolicy", env)), gym_fishing/envs/shared_env.py:57-79 (simulate_mdp_vec on a VecEnv)."""
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
STUBS = [os.path.join(ROOT, "tests", "stubs"), os.path.join(ROOT, "oracle", "refstub")]


def run_with_stubs(code):
    env = dict(os.environ)
    env["PYTHONPATH"] = os.pathsep.join(STUBS + [ROOT] + ([env["PYTHONPATH"]] if env.get("PYTHONPATH") else []))
    res = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, env=env, timeout=600)
    assert res.returncode == 0, res.stdout[-3000:] + res.stderr[-3000:]
    return res.stdout


CLASS_CHECKS = r'''
import inspect
import stable_baselines3.common.vec_env as sb3
import gym
import gym_fishing_b200 as gf
from gym_fishing_b200 import spaces

# a VecEnv subclass with nothing abstract left
assert issubclass(gf.FishingVecEnv, sb3.VecEnv)
assert gf.FishingVecEnv.__abstractmethods__ == frozenset(), gf.FishingVecEnv.__abstractmethods__
# same call signatures as the abstract interface (parameter names and order)
for name in ("reset", "step_async", "step_wait", "close", "get_attr", "set_attr", "env_method", "env_is_wrapped", "seed"):
    want = list(inspect.signature(getattr(sb3.VecEnv, name)).parameters)
    have = list(inspect.signature(getattr(gf.FishingVecEnv, name)).parameters)
    assert have[:len(want)] == want, (name, have, want)
# the spaces are the installed gym's classes, so isinstance checks in third-party code hold
assert spaces.BACKEND == "gym" and spaces.Box is gym.spaces.Box
# importing the package registered every id with gym, entry points resolving to the product classes
import importlib
reg = gym.envs.registration._REGISTRY
for v in (0, 1, 2, 4, 5, 6, 7, 8, 9, 10, 11):
    ep = reg[f"fishing-v{v}"][0]
    mod, cls = ep.split(":")
    assert mod.startswith("gym_fishing_b200") and isinstance(getattr(importlib.import_module(mod), cls), type)
assert "fishing-v3" not in reg
print("class checks ok")
'''


def test_vecenv_is_an_sb3_subclass_and_ids_reach_gym():
    assert "class checks ok" in run_with_stubs(CLASS_CHECKS)


INSTANCE_CHECKS = r'''
import numpy as np
import torch
import stable_baselines3.common.vec_env as sb3
import gym
import gym_fishing_b200 as gf

# gym.make goes through gym's registry to the product class (the reference: gym.make("fishing-v1"), README.md:35-38)
env = gym.make("fishing-v1", sigma=0.0)
obs = env.reset()
assert obs.shape == (1,) and isinstance(env, gf.BaseFishingEnv)
o, r, d, info = env.step(np.array([-0.8]))
assert o.shape == (1,) and not d and info == {}

n = 512
venv = gf.FishingVecEnv("fishing-v1", num_envs=n, sigma=0.02, Tmax=9, monitor=True, seed=3)
assert isinstance(venv, sb3.VecEnv) and venv.sb3_base_initialised and venv.num_envs == n
assert isinstance(venv.action_space, gym.spaces.Box)
obs = venv.reset()
assert tuple(obs.shape) == (n, 1)
ret = torch.zeros(n, dtype=torch.float64, device="cuda"); length = torch.zeros(n, dtype=torch.int64, device="cuda")
g = torch.Generator(device="cuda").manual_seed(0)
seen = 0
for t in range(40):
    a = torch.rand(n, 1, dtype=torch.float64, device="cuda", generator=g) * 0.6 - 1.0
    venv.step_async(a)                                   # the abstract interface: step = step_async + step_wait
    obs, rew, done, infos = venv.step_wait()
    ret += rew; length += 1
    idx, r, l = infos.episode_arrays()                   # every finished episode at once
    want = done.nonzero().reshape(-1).cpu().numpy()
    assert np.array_equal(idx, want)
    if len(idx):
        assert np.array_equal(r, ret[done].cpu().numpy()) and np.array_equal(l, length[done].cpu().numpy())
        j = int(idx[0])
        assert infos[j]["episode"] == {"r": float(r[0]), "l": int(l[0])}
        ti, tobs = infos.terminal_observations()
        assert np.array_equal(ti, idx) and tobs.shape == (len(idx), 1)
        assert float(infos[j]["terminal_observation"][0]) == float(tobs[0, 0])
        seen += len(idx)
    ret[done] = 0; length[done] = 0
assert seen > n                                          # several episodes per environment
assert venv.env_is_wrapped(object) == [False] * n
assert venv.get_attr("Tmax", indices=[0, 5]) == [9, 9]
assert venv.env_method("get_fish_population", np.array([-0.25]), indices=0)[0][0] == 0.75
venv.close()
print("instance checks ok")
'''


@pytest.mark.gpu
def test_vecenv_instance_under_the_sb3_interface():
    assert "instance checks ok" in run_with_stubs(INSTANCE_CHECKS)
