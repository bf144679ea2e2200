"""gym_fishing_b200 -- B200-native batched rollout engine for the gym_fishing environments (fishing-v0/v1/v2/v4).

    import gym_fishing_b200 as gf
    env  = gf.make("fishing-v1", sigma=0.01)                      # single env, reference gym surface
    venv = gf.FishingVecEnv("fishing-v1", num_envs=1 << 20)       # lock-step batch on the GPU

All dynamics run in hand-written sm_100a CUDA kernels (gym_fishing_b200/csrc) behind a C ABI
(include/fishing_b200.h); importing the package never touches the GPU, using an environment without a CUDA device
raises."""
from . import shared_env, spaces  # noqa: F401
from .registration import make, register, spec_ids  # noqa: F401
from . import envs  # noqa: F401  (registers the ids)
from .envs import BaseFishingEnv, FishingCtsEnv, FishingEnv, FishingModelError, FishingTippingEnv  # noqa: F401
from .vec_env import FishingVecEnv  # noqa: F401
from .engine import EnvBatch, bmsy  # noqa: F401
from .models.policies import BMSY, escapement, msy  # noqa: F401
from .rollouts import escapement_sweep, rollout_stats  # noqa: F401

__version__ = "0.1.0"
