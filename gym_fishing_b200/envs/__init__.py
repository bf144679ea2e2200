"""Environment classes and id registration (the ids of gym_fishing/envs/__init__.py:17-71; the reference has no
fishing-v3).  v0/v1/v2/v4 are the hot-path rows of SURVEY.md section 8a, v5..v11 the "next" row f1."""
from ..registration import register
from .base import BaseFishingEnv, FishingCtsEnv, FishingEnv, FishingModelError, FishingTippingEnv
from .growth_models import Allen, BevertonHolt, May, ModelUncertainty, Myers, NonStationary, Ricker

ENV_CLASSES = {
    0: FishingEnv, 1: FishingCtsEnv, 2: FishingTippingEnv, 4: FishingModelError,
    5: Allen, 6: BevertonHolt, 7: May, 8: Myers, 9: Ricker, 10: NonStationary, 11: ModelUncertainty,
}
for _version, _cls in ENV_CLASSES.items():
    register(f"fishing-v{_version}", f"{__name__}:{_cls.__name__}")

__all__ = ["BaseFishingEnv", "ENV_CLASSES"] + [c.__name__ for c in ENV_CLASSES.values()]
