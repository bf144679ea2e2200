"""fishing-v5..v11: the log-normal growth family with the reference's constructor signatures
(gym_fishing/envs/growth_models.py:6-204), computed by the same CUDA step kernel as the other environments.

Faithful to two upstream quirks: (i) the growth functions read the env's `params` DICT, not the attributes, so
`env.sigma = 0` (what the policies do) does not silence them; (ii) NonStationary mutates params["r"] at every draw
and never resets it."""
from __future__ import annotations

import copy
import ctypes as ct

import numpy as np

from .. import _native as nat
from ..vec_env import V11_MODELS, V11_PARAMS
from .base import BaseFishingEnv, _O_STATE  # noqa: F401


class _LogNormalEnv(BaseFishingEnv):
    _model_index = 0

    def _fparams(self, **over):
        p = self.params[self.model] if self._kind == nat.V11 else self.params
        kw = dict(r=p["r"], K=self.K, sigma=p["sigma"], C_=p.get("C", 0.5), x0=self.init_state, Tmax=self.Tmax,
                  seed=self._seed, growth_K=p["K"], M=p.get("M", 1.0), theta=p.get("theta", 1.0), q=p.get("q", 1.0),
                  b=p.get("b", 1.0), a=p.get("a", 0.0), alpha=p.get("alpha", 0.0))
        if self._kind == nat.V11:
            kw["models"] = [(m, self.params[m]) for m in self.models]
        kw.update(over)
        return nat.make_params(self._kind, **kw)

    def _soa(self):
        s = self._slot
        from .base import _O_K, _O_YEARS  # the K slot doubles as the 4-byte model index for v11 (never both)

        return nat.FishingSoA(s.ptr(_O_STATE), s.ptr(_O_YEARS), None, None, s.ptr(_O_K) if self._kind == nat.V11 else None)

    def step(self, action):
        if self._kind == nat.V11:
            from .base import _O_K

            self._slot.i32[_O_K // 4] = self.models.index(self.model)
        out = super().step(action)
        if self._kind == nat.V10:          # growth_models.py:148: params["r"] += alpha, kept across resets
            self.params["r"] = self.params["r"] + self.params["alpha"]
        return out

    def population_draw(self):
        """Growth of self.fish_population (scalar or array).  np.random.lognormal(mu, sigma) draws one normal PER ELEMENT."""
        torch = self._slot.torch
        arr = np.asarray(self.fish_population)
        scalar = arr.ndim == 0
        work = np.ascontiguousarray(arr.reshape(-1), dtype=np.float64)
        if self._noise_mode == "numpy":
            z = np.array([self._draw() for _ in range(work.size)], dtype=np.float64)
        else:
            z = np.zeros(work.size)
        d_in = torch.from_numpy(work).to(self._slot.device)
        d_z = torch.from_numpy(z).to(self._slot.device)
        d_out = torch.empty_like(d_in)
        p = self._fparams()
        idx = self.models.index(self.model) if self._kind == nat.V11 else 0
        with torch.cuda.device(self._slot.device):
            rc = nat.lib().fishing_population_draw_ln(ct.byref(p), idx, work.size, d_in.data_ptr(), d_z.data_ptr(),
                                                      d_out.data_ptr(), self._slot.stream())
        nat.check(rc, "fishing_population_draw_ln")
        if self._kind == nat.V10:
            self.params["r"] = self.params["r"] + self.params["alpha"]
        out = d_out.cpu().numpy().reshape(arr.shape)
        self.fish_population = out[()] if scalar else out
        return self.fish_population


class Allen(_LogNormalEnv):
    _kind = nat.V5

    def __init__(self, r=0.3, K=1, C=0.5, sigma=0.0, init_state=0.75, Tmax=100, file=None, **engine_kw):
        super().__init__(params={"r": r, "K": K, "sigma": sigma, "C": C, "x0": init_state}, Tmax=Tmax, file=file, **engine_kw)


class BevertonHolt(_LogNormalEnv):
    _kind = nat.V6

    def __init__(self, r=0.3, K=1, sigma=0.0, init_state=0.75, Tmax=100, file=None, **engine_kw):
        super().__init__(params={"r": r, "K": K, "sigma": sigma, "x0": init_state}, Tmax=Tmax, file=file, **engine_kw)


class Myers(_LogNormalEnv):
    _kind = nat.V8

    def __init__(self, r=1.0, K=1.0, M=1.0, theta=3.0, sigma=0.0, init_state=1.5, Tmax=100, file=None, **engine_kw):
        super().__init__(params={"r": r, "K": K, "sigma": sigma, "theta": theta, "M": M, "x0": init_state}, Tmax=Tmax,
                         file=file, **engine_kw)


class May(_LogNormalEnv):
    _kind = nat.V7

    def __init__(self, r=0.7, K=1.5, M=1.5, q=3, b=0.15, sigma=0.0, a=0.2, init_state=0.75, Tmax=100, file=None,
                 **engine_kw):
        super().__init__(params={"r": r, "K": K, "sigma": sigma, "q": q, "b": b, "a": a, "M": M, "x0": init_state},
                         Tmax=Tmax, file=file, **engine_kw)


class Ricker(_LogNormalEnv):
    _kind = nat.V9

    def __init__(self, r=0.3, K=1, sigma=0.0, init_state=0.75, Tmax=100, file=None, **engine_kw):
        super().__init__(params={"r": r, "K": K, "sigma": sigma, "x0": init_state}, Tmax=Tmax, file=file, **engine_kw)


class NonStationary(_LogNormalEnv):
    _kind = nat.V10

    def __init__(self, r=0.8, K=1, sigma=0.0, alpha=-0.007, init_state=0.75, Tmax=100, file=None, **engine_kw):
        super().__init__(params={"r": r, "K": K, "sigma": sigma, "alpha": alpha, "x0": init_state}, Tmax=Tmax, file=file,
                         **engine_kw)


class ModelUncertainty(_LogNormalEnv):
    _kind = nat.V11

    def __init__(self, models=None, params=None, Tmax=100, file=None, **engine_kw):
        self.models = list(V11_MODELS if models is None else models)
        self.model = self.models[0]
        super().__init__(Tmax=Tmax, file=file, **engine_kw)          # base defaults: K = 1, x0 = 0.75
        self.model = np.random.choice(self.models) if self._noise_mode == "numpy" else self._philox_choice()
        self.params = copy.deepcopy(V11_PARAMS) if params is None else params

    def _philox_choice(self):
        self._epoch += 1
        return self.models[(self._seed * 2654435761 + self._epoch) % len(self.models)]

    def reset(self):
        self.state = np.array([self.init_state / self.K - 1])
        self.fish_population = self.init_state
        self.model = np.random.choice(self.models) if self._noise_mode == "numpy" else self._philox_choice()
        self.years_passed = 0
        self.reward = 0
        self.harvest = 0
        return self.state
