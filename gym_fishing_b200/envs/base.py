"""Single-environment front end with the reference's gym surface, computed by the CUDA engine.

`make("fishing-v1")` returns an object with the attribute and method surface of the reference's BaseFishingEnv
(gym_fishing/envs/base_fishing_env.py:16-164): reset() -> obs, step(action) -> (obs, reward, done, {}), render,
close, simulate / plot / policyfn / plot_policy, get_quota / get_action / get_fish_population / get_state /
harvest_draw / population_draw, and the mutable attributes K r sigma init_state Tmax fish_population state
years_passed reward harvest params action_space observation_space.

Every step()/reset()/population_draw() is ONE launch of the same kernels the batched engine uses, on a one-element
structure-of-arrays slot (one 128-byte upload, one kernel, one 128-byte download); there is no CPU implementation of
the dynamics.  Like the reference, attributes are read at call time, so `env.sigma = 0`, `env.init_state = 0.3`,
`env.state = ...` between calls behave the same.

Noise: the reference draws np.random.normal(0, 1) from numpy's GLOBAL generator once per step (even when sigma == 0),
twice per fishing-v4 construction/reset, once in BMSY (base_fishing_env.py:130, fishing_model_error.py:37-43,
policies.py:63).  With noise="numpy" (default) this class makes the same calls in the same order and hands the draws
to the kernel, so `np.random.seed(s)` reproduces the reference's trajectories bit for bit.  noise="philox" uses the
engine's own counter-based stream instead (what the batched VecEnv does).
"""
from __future__ import annotations

import csv
import ctypes as ct

import numpy as np

from .. import _native as nat
from ..spaces import Discrete, unit_box

try:  # pragma: no cover - gym is not installed in the build image
    import gymnasium as _gym  # type: ignore
    _EnvBase = _gym.Env
except Exception:  # noqa: BLE001
    try:
        import gym as _gym  # type: ignore
        _EnvBase = _gym.Env
    except Exception:  # noqa: BLE001
        _EnvBase = object

# byte offsets inside the 128-byte slot (one environment, structure-of-arrays with n = 1)
_O_STATE, _O_NOISE, _O_ACTION, _O_K, _O_R, _O_YEARS, _O_OBS, _O_REWARD, _O_TERM, _O_DONE, _O_STOCK, _O_HARV, _O_ZK, _O_ZR = (
    0, 8, 16, 24, 32, 40, 48, 56, 64, 72, 80, 88, 96, 104)


class _Slot:
    """128 bytes of page-locked host memory that the kernels read and write directly (zero-copy: under unified virtual
    addressing a cudaHostAlloc'ed buffer has the same address on the device), with typed numpy views over it.  A scalar
    step is then one kernel launch and one stream synchronisation -- no staging copies in either direction (measured: 35 -> 25 us per
    env.step call, most of what is left being the launch and the synchronisation)."""

    def __init__(self, device):
        torch = nat.require_cuda()
        nat.lib()
        self.torch = torch
        self.device = torch.device(device)
        if self.device.index is None:
            self.device = torch.device("cuda", torch.cuda.current_device())
        with torch.cuda.device(self.device):
            torch.cuda.current_stream()                      # make sure the device has a context before the allocation
            p = ct.c_void_p()
            nat.check(nat.lib().fishing_host_alloc(ct.byref(p), 128), "fishing_host_alloc")
        self._ptr = p
        self._buf = (ct.c_char * 128).from_address(p.value)
        raw = np.frombuffer(self._buf, dtype=np.uint8)
        raw[:] = 0
        self.f64 = raw.view(np.float64)
        self.f32 = raw.view(np.float32)
        self.i64 = raw.view(np.int64)
        self.i32 = raw.view(np.int32)
        self.u8 = raw
        self.base = p.value

    def ptr(self, off):
        return self.base + off

    def upload(self):
        """Nothing to copy: the kernel reads the slot in place (host writes are visible to a kernel launched later)."""

    def download(self):
        """Wait for the kernel; its writes to the slot are visible to the host after the stream synchronisation."""
        self.torch.cuda.current_stream(self.device).synchronize()

    def stream(self):
        return self.torch.cuda.current_stream(self.device).cuda_stream

    def __del__(self):  # pragma: no cover
        try:
            if self._ptr is not None:
                self.f64 = self.f32 = self.i64 = self.i32 = self.u8 = None
                nat.lib().fishing_host_free(self._ptr)
                self._ptr = None
        except Exception:  # noqa: BLE001 -- interpreter shutdown
            pass


class BaseFishingEnv(_EnvBase):
    metadata = {"render.modes": ["human"]}
    _kind = nat.V1

    def __init__(self, params=None, Tmax=100, file=None, *, device="cuda", noise="numpy", seed=0):
        if params is None:
            params = {"r": 0.3, "K": 1, "sigma": 0.0, "x0": 0.75}
        if noise not in ("numpy", "philox"):
            raise ValueError("noise must be 'numpy' or 'philox'")
        self._slot = _Slot(device)  # raises without a CUDA device / without the built library
        self.K = params["K"]
        self.r = params["r"]
        self.sigma = params["sigma"]
        self.init_state = params["x0"]
        self.params = params
        self.Tmax = Tmax
        self.file = file
        self.fish_population = self.init_state
        self.reward = 0
        self.harvest = 0
        self.years_passed = 0
        self.action = None  # last action; the reference's render() reads an attribute it never sets (shared_env.py:9)
        self._noise_mode = noise
        self._seed = int(seed)
        self._step_idx = 0
        self._epoch = 0
        if file is not None:
            self.write_obj = open(file, "w+")
        self.action_space = unit_box()
        self.observation_space = unit_box()
        self.state = np.array([self.init_state / self.K - 1])

    # ------------------------------------------------------------------------------------------------ engine glue
    def _is_discrete(self):
        return isinstance(self.action_space, Discrete)

    def _fparams(self, **over):
        """The parameter block for the next launch, from the attributes as they are NOW (callers mutate them between steps,
        like the reference's callers do).  The ctypes structure is rebuilt only when one of them changed: building it costs
        about as much as the kernel launch itself."""
        v4 = self._kind == nat.V4
        key = (self.r_mean if v4 else self.r, self.K_mean if v4 else self.K, self.sigma, getattr(self, "C", 0.5), self.init_state,
               self.Tmax, getattr(self, "n_actions", 100), getattr(self, "sigma_p", 0.1), self._seed)
        if not over:
            cached = getattr(self, "_p_cache", None)
            if cached is not None and cached[0] == key:
                return cached[1]
        kw = dict(zip(("r", "K", "sigma", "C_", "x0", "Tmax", "n_actions", "sigma_p", "seed"), key))
        kw.update(over)
        p = nat.make_params(self._kind, **kw)
        if not over:
            self._p_cache = (key, p)
        return p

    def _soa(self):
        s = self._slot
        per_env = self._kind == nat.V4
        return nat.FishingSoA(s.ptr(_O_STATE), s.ptr(_O_YEARS), s.ptr(_O_K) if per_env else None,
                              s.ptr(_O_R) if per_env else None, None)

    def _draw(self):
        """One standard normal from numpy's global generator, looked up at call time like the reference does."""
        return float(np.random.normal(0, 1))

    # ------------------------------------------------------------------------------------------------ gym API
    def step(self, action):
        s = self._slot
        discrete = self._is_discrete()
        if discrete:
            s.i64[_O_ACTION // 8] = int(action)
            adt = nat.ACT_I64
        else:
            a = np.asarray(action)
            if a.dtype == np.float32:  # numpy>=2 keeps the quota in float32 for a float32 action (NEP 50)
                s.f32[_O_ACTION // 4] = a.reshape(-1)[0]
                adt = nat.ACT_F32
            else:
                s.f64[_O_ACTION // 8] = np.asarray(a, dtype=np.float64).reshape(-1)[0]
                adt = nat.ACT_F64
        self.action = action
        s.f64[_O_STATE // 8] = self.state[0]
        s.i32[_O_YEARS // 4] = self.years_passed
        if self._kind == nat.V4:
            s.f64[_O_K // 8] = self.K
            s.f64[_O_R // 8] = self.r
        use_np = self._noise_mode == "numpy"
        if use_np:
            s.f64[_O_NOISE // 8] = self._draw()
        s.upload()
        p = self._fparams()
        with s.torch.cuda.device(s.device):
            rc = nat.lib().fishing_step(ct.byref(p), 1, 0, self._step_idx, s.ptr(_O_ACTION), adt,
                                        s.ptr(_O_NOISE) if use_np else None, None, self._soa(), s.ptr(_O_OBS),
                                        s.ptr(_O_REWARD), s.ptr(_O_DONE), None, s.ptr(_O_STOCK), s.ptr(_O_HARV), s.stream())
        nat.check(rc, "fishing_step")
        s.download()
        self._step_idx += 1
        self.state = np.array([s.f64[_O_OBS // 8]])
        self.fish_population = np.float64(s.f64[_O_STOCK // 8])
        self.harvest = float(s.f64[_O_HARV // 8])
        self.reward = float(s.f64[_O_REWARD // 8])
        self.years_passed = int(s.i32[_O_YEARS // 4])
        done = bool(s.u8[_O_DONE])
        return self.state, self.reward, done, {}

    def _device_reset(self, zK=None, zr=None):
        s = self._slot
        per_env = self._kind == nat.V4
        if per_env and zK is not None:
            s.f64[_O_ZK // 8], s.f64[_O_ZR // 8] = zK, zr
        s.upload()
        p = self._fparams()
        with s.torch.cuda.device(s.device):
            rc = nat.lib().fishing_reset(ct.byref(p), 1, 0, self._epoch, None,
                                         s.ptr(_O_ZK) if (per_env and zK is not None) else None, self._soa(), None,
                                         s.stream())
        nat.check(rc, "fishing_reset")
        s.download()
        self._epoch += 1
        if per_env:
            self.K = np.float64(s.f64[_O_K // 8])
            self.r = np.float64(s.f64[_O_R // 8])
        self.state = np.array([s.f64[_O_STATE // 8]])
        self.years_passed = int(s.i32[_O_YEARS // 4])

    def reset(self):
        self._device_reset()
        self.fish_population = self.init_state
        self.reward = 0
        self.harvest = 0
        return self.state

    def render(self, mode="human"):
        """One CSV row [years_passed, state, action, reward] (shared_env.py:8-12).  The reference raises here because
        it never assigns self.action; this implementation records the last action in step()."""
        row = [self.years_passed, self.state[0], self.action, self.reward]
        if self.file is not None:
            csv.writer(self.write_obj).writerow(row)
        return row

    def close(self):
        if self.file is not None:
            self.write_obj.close()

    def seed(self, seed=None):
        """Not in the reference (base_fishing_env.py:13).  Seeds the Philox stream, and numpy's global generator when
        this env draws its noise there."""
        self._seed = 0 if seed is None else int(seed)
        if self._noise_mode == "numpy" and seed is not None:
            np.random.seed(seed)
        return [self._seed]

    # ------------------------------------------------------------------------------------------------ helpers
    def get_quota(self, action):
        if self._is_discrete():
            return (action / self.n_actions) * self.K
        action = np.clip(action, self.action_space.low, self.action_space.high)[0]
        return (action + 1) * self.K

    def get_action(self, quota):
        if self._is_discrete():
            return round(quota * self.n_actions / self.K)
        return quota / self.K - 1

    def get_fish_population(self, state):
        self.fish_population = (state[0] + 1) * self.K
        return self.fish_population

    def get_state(self, fish_population):
        self.state = np.array([fish_population / self.K - 1])
        return self.state

    def harvest_draw(self, quota):
        self.harvest = min(self.fish_population, quota)
        self.fish_population = max(self.fish_population - self.harvest, 0.0)
        return self.harvest

    def population_draw(self):
        """Growth of self.fish_population (scalar or array, float32 or float64) with one scalar normal draw, on the
        device (fishing_population_draw)."""
        torch = self._slot.torch
        x = self.fish_population
        z = self._draw() if self._noise_mode == "numpy" else 0.0
        arr = np.asarray(x)
        scalar = arr.ndim == 0
        np_f64_params = isinstance(self.K, np.floating) or isinstance(self.r, np.floating)
        f32 = arr.dtype == np.float32 and not np_f64_params
        work = np.ascontiguousarray(arr.reshape(-1), dtype=np.float32 if f32 else np.float64)
        d_in = torch.from_numpy(work).to(self._slot.device)
        d_out = torch.empty_like(d_in)
        p = self._fparams(r=self.r, K=self.K)
        with torch.cuda.device(self._slot.device):
            rc = nat.lib().fishing_population_draw(ct.byref(p), work.size, d_in.data_ptr(), int(f32), z,
                                                   d_out.data_ptr(), self._slot.stream())
        nat.check(rc, "fishing_population_draw")
        out = d_out.cpu().numpy().reshape(arr.shape)
        self.fish_population = out[()] if scalar else out
        return self.fish_population

    # ------------------------------------------------------------------------------------------------ reporting
    def simulate(env, model, reps=1):
        """simulate_mdp (shared_env.py:29-54).  With this package's escapement / msy policy the repetitions run as one fused
        rollout on the device (shared_env.simulate_on_device: same table, same use of numpy's generator); any other model
        -- an SB3 agent, a user policy -- goes through the host loop over step()."""
        from ..shared_env import simulate_mdp, simulate_on_device

        df = simulate_on_device(env, model, reps)
        return df if df is not None else simulate_mdp(env, model, reps)

    def policyfn(env, model, reps=1):
        from ..shared_env import estimate_policyfn

        return estimate_policyfn(env, model, reps)

    def plot(self, df, output="results.png"):
        from ..shared_env import plot_mdp

        return plot_mdp(self, df, output)

    def plot_policy(self, df, output="results.png"):
        from ..shared_env import plot_policyfn

        return plot_policyfn(self, df, output)


class FishingEnv(BaseFishingEnv):
    """fishing-v0: discrete actions, quota = action / n_actions * K (gym_fishing/envs/fishing_env.py:6-24)."""
    _kind = nat.V0

    def __init__(self, r=0.3, K=1, sigma=0.0, n_actions=100, init_state=0.75, Tmax=100, file=None, **engine_kw):
        super().__init__(params={"r": r, "K": K, "sigma": sigma, "x0": init_state}, Tmax=Tmax, file=file, **engine_kw)
        self.n_actions = n_actions
        self.action_space = Discrete(n_actions)


class FishingCtsEnv(BaseFishingEnv):
    """fishing-v1: continuous quota (gym_fishing/envs/fishing_cts_env.py:4-12)."""
    _kind = nat.V1

    def __init__(self, r=0.3, K=1, sigma=0.0, init_state=0.75, Tmax=100, file=None, **engine_kw):
        super().__init__(params={"r": r, "K": K, "sigma": sigma, "x0": init_state}, Tmax=Tmax, file=file, **engine_kw)


class FishingTippingEnv(BaseFishingEnv):
    """fishing-v2: tipping-point growth x*exp(r(1-x/K)(x-C) + x*sigma*z) (gym_fishing/envs/fishing_tipping_env.py:6-35)."""
    _kind = nat.V2

    def __init__(self, r=0.3, K=1, C=0.5, sigma=0.0, init_state=0.75, Tmax=100, file=None, **engine_kw):
        super().__init__(params={"r": r, "K": K, "sigma": sigma, "C": C, "x0": init_state}, Tmax=Tmax, file=file,
                         **engine_kw)
        self.C = C


class FishingModelError(BaseFishingEnv):
    """fishing-v4: K and r redrawn (K first) at construction and at every reset; reset returns the UN-normalised
    initial state (gym_fishing/envs/fishing_model_error.py:6-48)."""
    _kind = nat.V4

    def __init__(self, K_mean=1.0, r_mean=0.3, price=1.0, sigma=0.0, sigma_p=0.1, init_state=0.75, Tmax=100, file=None,
                 **engine_kw):
        super().__init__(params={"r": r_mean, "K": K_mean, "sigma": sigma, "r_mean": r_mean, "K_mean": K_mean,
                                 "sigma_p": sigma_p, "x0": init_state}, Tmax=Tmax, file=file, **engine_kw)
        self.K_mean = K_mean
        self.r_mean = r_mean
        self.sigma_p = sigma_p
        state, years = self.state, self.years_passed
        self._redraw()
        self.state, self.years_passed = state, years  # construction draws parameters but keeps state = x0/K_mean - 1

    def _redraw(self):
        if self._noise_mode == "numpy":
            zK = self._draw()
            zr = self._draw()
            self._device_reset(zK, zr)
        else:
            self._device_reset()

    def reset(self):
        self._redraw()
        self.fish_population = self.init_state
        self.harvest = 0
        return self.state
