"""Classical fisheries policies with the reference's protocol `model.predict(obs, **kw) -> (action, state)`
(gym_fishing/models/policies.py:4-31,51-67).  The interactive `user_action` policy is out of scope.

BMSY runs on the device (fishing_bmsy: the 10001-point float32 grid, deterministic growth and first-argmax in one
kernel).  To keep numpy's global random stream in step with the reference when the env draws its noise there, BMSY
consumes the one normal draw the reference's vectorised population_draw() would have consumed (policies.py:63) and
resets the env (policies.py:66)."""
from __future__ import annotations

import numpy as np

from .. import _native as nat
from ..engine import bmsy as _bmsy_kernel


def BMSY(env):
    """Stock size with maximum deterministic surplus production.  fishing-v0/v1/v2/v4: the value the reference returns, bit
    for bit (np.float32, np.float64 for v4).  fishing-v5..v11: the same procedure, but the surplus curve is evaluated in
    float64 where the reference's float32 grid makes numpy's `log` / `power` float32; the curve is flat to ~1e-8 around its
    maximum, so S can land on a neighbouring point of the 10001-point grid (tests/test_gpu_lognormal.py bounds it by two grid
    steps).  Resets the environment."""
    kind = env._kind
    if kind in nat.LOGNORMAL_KINDS:
        # fishing-v5..v11: no fused kernel; same procedure as the reference through the env's own (device) population_draw
        grid = np.linspace(env.observation_space.low, env.observation_space.high, num=10001,
                           dtype=env.observation_space.dtype)
        x_0 = np.asarray([env.get_fish_population(g) for g in grid])
        env.fish_population = x_0
        keep = env.sigma
        env.sigma = 0                      # no effect on these models (they read params["sigma"]), kept for fidelity
        growth = env.population_draw() - x_0
        env.sigma = keep
        S = x_0[np.argmax(growth)]
        env.reset()
        return S
    K_f64 = kind == nat.V4 or isinstance(env.K, np.floating) or isinstance(env.r, np.floating)
    if getattr(env, "_noise_mode", "philox") == "numpy":
        env._draw()
    S, _ = _bmsy_kernel(kind, r=env.r, K=env.K, C=getattr(env, "C", 0.5), K_is_f64=K_f64)
    env.reset()
    return np.float64(S) if K_f64 else np.float32(S)


class msy:
    """Constant quota equal to the surplus production at BMSY."""

    def __init__(self, env, **kwargs):
        self.env = env
        self.S = BMSY(env)
        keep = env.sigma
        env.fish_population = self.S
        env.sigma = 0
        try:
            self.msy = env.population_draw() - self.S
        finally:
            env.sigma = keep
        env.reset()

    def predict(self, obs, **kwargs):
        return self.env.get_action(self.msy), obs


class escapement:
    """Constant escapement: harvest everything above S = BMSY."""

    def __init__(self, env, **kwargs):
        self.env = env
        self.S = BMSY(env)

    def predict(self, obs, **kwargs):
        x = self.env.get_fish_population(obs)
        return self.env.get_action(max(x - self.S, 0.0)), obs


class user_action:
    """Interactive policy: shows the stock and asks for a harvest quota on stdin (models/policies.py:34-48 of the
    reference).  Host-side only -- the quota goes through the env's own get_action like any other policy."""

    def __init__(self, env, **kwargs):
        self.env = env

    def predict(self, obs, **kwargs):
        fish_population = self.env.get_fish_population(obs)
        quota = input("fish population: " + str(fish_population) + ". Your harvest quota: ")
        return self.env.get_action(float(quota)), obs
