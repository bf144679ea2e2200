"""Throw-away stand-in for OpenAI gym 0.17.x, just large enough to import the UNMODIFIED reference.

TEST INFRASTRUCTURE ONLY (used by tests/golden/make_golden.py in the authoring container to execute
/root/reference and record golden vectors).  Nothing in the product package imports this.
It provides what gym_fishing/envs/base_fishing_env.py:1-3 and gym_fishing/envs/__init__.py:1 import.
"""
from . import spaces  # noqa: F401
from .envs.registration import make, register  # noqa: F401


class Env:
    metadata = {}
    action_space = None
    observation_space = None

    def step(self, action):
        raise NotImplementedError

    def reset(self):
        raise NotImplementedError

    def render(self, mode="human"):
        raise NotImplementedError

    def close(self):
        pass
