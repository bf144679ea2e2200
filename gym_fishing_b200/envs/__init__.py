"""Environment classes and id registration (mirrors gym_fishing/envs/__init__.py:17-35 for the four ids in scope)."""
from ..registration import register
from .base import BaseFishingEnv, FishingCtsEnv, FishingEnv, FishingModelError, FishingTippingEnv

register(id="fishing-v0", entry_point="gym_fishing_b200.envs:FishingEnv")
register(id="fishing-v1", entry_point="gym_fishing_b200.envs:FishingCtsEnv")
register(id="fishing-v2", entry_point="gym_fishing_b200.envs:FishingTippingEnv")
register(id="fishing-v4", entry_point="gym_fishing_b200.envs:FishingModelError")

from .growth_models import Allen, BevertonHolt, May, ModelUncertainty, Myers, NonStationary, Ricker  # noqa: E402

# "next row" f1: gym_fishing/envs/__init__.py:38-71
register(id="fishing-v5", entry_point="gym_fishing_b200.envs:Allen")
register(id="fishing-v6", entry_point="gym_fishing_b200.envs:BevertonHolt")
register(id="fishing-v7", entry_point="gym_fishing_b200.envs:May")
register(id="fishing-v8", entry_point="gym_fishing_b200.envs:Myers")
register(id="fishing-v9", entry_point="gym_fishing_b200.envs:Ricker")
register(id="fishing-v10", entry_point="gym_fishing_b200.envs:NonStationary")
register(id="fishing-v11", entry_point="gym_fishing_b200.envs:ModelUncertainty")

__all__ = ["Allen", "BevertonHolt", "May", "Myers", "Ricker", "NonStationary", "ModelUncertainty", "BaseFishingEnv", "FishingEnv", "FishingCtsEnv", "FishingTippingEnv", "FishingModelError"]
