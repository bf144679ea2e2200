"""Environment classes and id registration (mirrors gym_fishing/envs/__init__.py:17-35 for the four ids in scope)."""
from ..registration import register
from .base import BaseFishingEnv, FishingCtsEnv, FishingEnv, FishingModelError, FishingTippingEnv

register(id="fishing-v0", entry_point="gym_fishing_b200.envs:FishingEnv")
register(id="fishing-v1", entry_point="gym_fishing_b200.envs:FishingCtsEnv")
register(id="fishing-v2", entry_point="gym_fishing_b200.envs:FishingTippingEnv")
register(id="fishing-v4", entry_point="gym_fishing_b200.envs:FishingModelError")

__all__ = ["BaseFishingEnv", "FishingEnv", "FishingCtsEnv", "FishingTippingEnv", "FishingModelError"]
