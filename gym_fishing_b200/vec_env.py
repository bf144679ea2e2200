"""FishingVecEnv -- the batched, lock-step environment class (new; the reference has none, it only *consumes* a
stable-baselines3 VecEnv in simulate_mdp_vec, gym_fishing/envs/shared_env.py:15-26,57-79).

It satisfies that consumer and the SB3 VecEnv duck type: num_envs, observation_space, action_space, reset(),
step(actions) / step_async + step_wait, get_attr, set_attr, env_method, seed, close, auto-reset on done with
infos[i]["terminal_observation"].  Two I/O modes:

  * io="torch" (default): actions / observations / rewards / dones are CUDA tensors; zero copies, the kernels run on
    the current torch stream.  `step_into(...)` writes straight into slices of a caller-owned rollout buffer.
  * io="numpy": host arrays in and out through the library's pinned, chunk-pipelined host front end
    (fishing_vec_* in include/fishing_b200.h) -- what a stable-baselines3 loop or the reference's own
    simulate_mdp_vec sees.

All dynamics run in the sm_100a kernels of libfishing_b200.so; there is no CPU path.
"""
from __future__ import annotations

import ctypes as ct
import os

import numpy as np

from . import _native as nat
from .engine import EnvBatch
from .spaces import Discrete, unit_box

try:  # a real stable-baselines3 (absent from the build image): be one of ITS VecEnvs, so that PPO("MlpPolicy", venv) takes
    # the object as it is instead of wrapping it in DummyVecEnv (tests/test-PPO.py:9-21 of the reference, pinned
    # stable-baselines3==0.9.0 in examples/requirements.txt:29)
    from stable_baselines3.common.vec_env import VecEnv as _VecEnvBase  # type: ignore
except Exception:  # noqa: BLE001
    _VecEnvBase = object

_CTOR_KEYS = {
    nat.V0: ("r", "K", "sigma", "n_actions", "init_state", "Tmax"),
    nat.V1: ("r", "K", "sigma", "init_state", "Tmax"),
    nat.V2: ("r", "K", "C", "sigma", "init_state", "Tmax"),
    nat.V4: ("K_mean", "r_mean", "price", "sigma", "sigma_p", "init_state", "Tmax"),
    # log-normal growth family, gym_fishing/envs/growth_models.py:7-16,29-31,44-53,76-87,110-112,125-134,154-182
    nat.V5: ("r", "K", "C", "sigma", "init_state", "Tmax"),
    nat.V6: ("r", "K", "sigma", "init_state", "Tmax"),
    nat.V7: ("r", "K", "M", "q", "b", "sigma", "a", "init_state", "Tmax"),
    nat.V8: ("r", "K", "M", "theta", "sigma", "init_state", "Tmax"),
    nat.V9: ("r", "K", "sigma", "init_state", "Tmax"),
    nat.V10: ("r", "K", "sigma", "alpha", "init_state", "Tmax"),
    nat.V11: ("models", "params", "Tmax"),
}

# constructor defaults of the log-normal family (same file:lines)
_LN_DEFAULTS = {
    nat.V5: dict(r=0.3, K=1, C=0.5, sigma=0.0, init_state=0.75),
    nat.V6: dict(r=0.3, K=1, sigma=0.0, init_state=0.75),
    nat.V7: dict(r=0.7, K=1.5, M=1.5, q=3, b=0.15, sigma=0.0, a=0.2, init_state=0.75),
    nat.V8: dict(r=1.0, K=1.0, M=1.0, theta=3.0, sigma=0.0, init_state=1.5),
    nat.V9: dict(r=0.3, K=1, sigma=0.0, init_state=0.75),
    nat.V10: dict(r=0.8, K=1, sigma=0.0, alpha=-0.007, init_state=0.75),
}
V11_MODELS = ["allen", "beverton_holt", "myers", "may", "ricker"]
V11_PARAMS = {
    "allen": {"r": 0.3, "K": 1.0, "sigma": 0.0, "C": 0.5, "x0": 0.75},
    "beverton_holt": {"r": 0.3, "K": 1, "sigma": 0.0, "x0": 0.75},
    "myers": {"r": 1.0, "K": 1.0, "M": 1.0, "theta": 3.0, "sigma": 0.0, "x0": 1.5},
    "may": {"r": 0.7, "K": 1.5, "M": 1.5, "q": 3, "b": 0.15, "sigma": 0.0, "a": 0.2, "x0": 0.75},
    "ricker": {"r": 0.3, "K": 1, "sigma": 0.0, "x0": 0.75},
}


def growth_kwargs(kind, kw):
    """Extra EnvBatch keyword arguments (`params`-dict entries) for fishing-v5..v11."""
    if kind not in nat.LOGNORMAL_KINDS:
        return {}
    if kind == nat.V11:
        models = list(kw.get("models", V11_MODELS))
        params = kw.get("params", V11_PARAMS)
        return dict(models=[(m, dict(params[m])) for m in models])
    d = dict(_LN_DEFAULTS[kind])
    d.update({k: v for k, v in kw.items() if k in d})
    out = dict(growth_K=d["K"])
    for k in ("M", "theta", "q", "b", "a", "alpha"):
        if k in d:
            out[k] = d[k]
    return out


def normalise_kwargs(kind, kw):
    """Map the reference's constructor keywords (fishing_env.py:7-16, fishing_cts_env.py:5-7,
    fishing_tipping_env.py:7-16, fishing_model_error.py:9-19) to engine parameters."""
    kw = dict(kw)
    kw.pop("file", None)
    bad = set(kw) - set(_CTOR_KEYS[kind])
    if bad:
        raise TypeError(f"unexpected keyword argument(s) {sorted(bad)} for this environment id")
    out = dict(r=0.3, K=1.0, sigma=0.0, C=0.5, x0=0.75, Tmax=100, n_actions=100, sigma_p=0.1)
    if kind == nat.V11:      # base-class defaults scale the state whatever the model (growth_models.py:183-189)
        out["Tmax"] = kw.get("Tmax", 100)
        out["K"] = 1
        return out
    if kind in _LN_DEFAULTS:
        d = dict(_LN_DEFAULTS[kind])
        d.update({k: v for k, v in kw.items() if k in d})
        out.update(r=d["r"], K=d["K"], sigma=d["sigma"], x0=d["init_state"], Tmax=kw.get("Tmax", 100))
        if "C" in d:
            out["C"] = d["C"]
        return out
    if kind == nat.V4:
        out["r"] = kw.get("r_mean", 0.3)
        out["K"] = kw.get("K_mean", 1.0)
    else:
        out["r"] = kw.get("r", 0.3)
        out["K"] = kw.get("K", 1)
    for k in ("sigma", "C", "Tmax", "n_actions", "sigma_p"):
        if k in kw:
            out[k] = kw[k]
    out["x0"] = kw.get("init_state", 0.75)
    return out


_ZC_MAX = int(os.environ.get("FISHING_VEC_ZC_MAX", 65536))   # keep in step with zero_copy_max_envs() in csrc/abi.cu


class LazyInfos:
    """Sequence of per-env info dicts materialised on access (a list of 2^20 dicts per step would dominate the run
    time).  infos[i] == {"terminal_observation": array([s])} if env i finished in this step, else {}."""

    def __init__(self, done, terminal, episode=None):
        self._done, self._terminal, self._episode = done, terminal, episode

    def __len__(self):
        return len(self._done)

    def __getitem__(self, i):
        if isinstance(i, slice):
            return [self[j] for j in range(*i.indices(len(self)))]
        if not bool(self._done[i]):
            return {}
        info = {}
        if self._terminal is not None:
            t = self._terminal[i]
            info["terminal_observation"] = t.reshape(1) if hasattr(t, "reshape") else np.array([t])
        if self._episode is not None:      # VecMonitor-style statistics of the episode that just ended
            r, l = self._episode
            info["episode"] = {"r": float(r[i]), "l": int(l[i])}
        return info

    def __iter__(self):
        return (self[i] for i in range(len(self)))

    # vectorised views: nobody has to walk 2^20 dictionaries to find the handful of episodes that ended
    def done_indices(self):
        """indices of the environments whose episode ended in this step (numpy int64 array)."""
        d = self._done
        if hasattr(d, "nonzero") and not isinstance(d, np.ndarray):      # torch tensor
            return d.nonzero().reshape(-1).cpu().numpy()
        return np.flatnonzero(np.asarray(d))

    def episode_arrays(self):
        """(indices, returns, lengths) of the episodes that ended in this step -- the content of every infos[i]["episode"]
        at once.  Needs monitor=True."""
        if self._episode is None:
            raise ValueError("episode statistics are kept with FishingVecEnv(..., monitor=True)")
        idx = self.done_indices()
        r, l = self._episode
        if hasattr(r, "cpu"):
            import torch

            j = torch.as_tensor(idx, device=r.device)
            return idx, r[j].cpu().numpy(), l[j].cpu().numpy()
        return idx, np.asarray(r)[idx], np.asarray(l)[idx]

    def terminal_observations(self):
        """(indices, terminal observations [k, 1]) of the episodes that ended in this step."""
        idx = self.done_indices()
        if self._terminal is None:
            return idx, np.empty((0, 1))
        t = self._terminal
        if hasattr(t, "cpu"):
            import torch

            return idx, t[torch.as_tensor(idx, device=t.device)].cpu().numpy().reshape(-1, 1)
        return idx, np.asarray([t[i] for i in idx], dtype=np.float64).reshape(-1, 1)


class _LazyTerminal:
    """Terminal observations of the latest numpy-mode step, copied from the device the first time one is read."""

    def __init__(self, env):
        self._env, self._step, self._arr = env, env._steps_done, None

    def __getitem__(self, i):
        if self._arr is None:
            env = self._env
            if env._steps_done != self._step:
                raise RuntimeError("terminal observations must be read before the next step() call")
            nat.check(nat.lib().fishing_vec_fetch_terminal(env._handle, env._h_term.array.ctypes.data),
                      "fishing_vec_fetch_terminal")
            self._arr = env._h_term.array
        return self._arr[i]


class _PinnedArray:
    """numpy view over page-locked host memory from fishing_host_alloc (freed with the owner)."""

    def __init__(self, n, dtype):
        dtype = np.dtype(dtype)
        self.nbytes = max(1, n * dtype.itemsize)
        p = ct.c_void_p()
        nat.check(nat.lib().fishing_host_alloc(ct.byref(p), self.nbytes), "fishing_host_alloc")
        self.ptr = p
        buf = (ct.c_char * self.nbytes).from_address(p.value)
        self.array = np.frombuffer(buf, dtype=dtype, count=n)

    def free(self):
        if self.ptr is not None:
            nat.lib().fishing_host_free(self.ptr)
            self.ptr = None


class FishingVecEnv(_VecEnvBase):
    """A subclass of stable_baselines3.common.vec_env.VecEnv when that package is importable (a plain class otherwise): the
    whole abstract surface -- reset, step_async, step_wait, close, get_attr, set_attr, env_method, env_is_wrapped, seed -- is
    implemented here, so an SB3 algorithm takes the object as its vectorised environment unchanged."""

    def __init__(self, env_id="fishing-v1", num_envs=1, *, device="cuda", io="torch", seed=0, auto_reset=True,
                 env_id_offset=0, monitor=False, sb3_dtypes=False, packed_done=False, **env_kwargs):
        """monitor=True (torch mode): keep per-env episode return / length on the device -- inside the step kernel -- and
        report them in infos[i]["episode"] = {"r", "l"} when an episode ends, like stable-baselines3's VecMonitor
        (infos.episode_arrays() returns all of them at once).
        sb3_dtypes=True: observations and rewards come back as float32, which is what a stable-baselines3 VecEnv hands
        to the learner (the observation space is declared float32 and DummyVecEnv stores rewards in a float32 buffer);
        the state itself stays float64 on the device, the step kernel writes the float32 copies.  In numpy mode this
        halves the PCIe traffic.
        packed_done=True (numpy mode, step_many): the done flags cross PCIe bit-packed and are unpacked on the host."""
        if env_id not in nat.KIND_OF_ID:
            raise KeyError(f"unknown environment id {env_id!r}")
        if io not in ("torch", "numpy"):
            raise ValueError("io must be 'torch' or 'numpy'")
        nat.require_cuda()
        nat.lib()
        self.env_id = env_id
        self.kind = nat.KIND_OF_ID[env_id]
        self.num_envs = int(num_envs)
        self.io = io
        self.cfg = normalise_kwargs(self.kind, env_kwargs)
        self._growth = growth_kwargs(self.kind, env_kwargs)
        self.auto_reset = bool(auto_reset)
        self._seed = int(seed)
        self.env_id_offset = int(env_id_offset)
        self.observation_space = unit_box()
        self.action_space = Discrete(self.cfg["n_actions"]) if self.kind == nat.V0 else unit_box()
        self.metadata = {"render.modes": ["human"]}
        self._pending = None
        self.sb3_dtypes = bool(sb3_dtypes)
        self.packed_done = bool(packed_done)
        self.monitor = bool(monitor)
        if self.monitor and io != "torch":
            raise ValueError("monitor=True needs io='torch'")
        if self.packed_done and (io != "numpy" or self.num_envs % 2):
            raise ValueError("packed_done=True needs io='numpy' and an even number of environments")
        if _VecEnvBase is not object:   # let stable-baselines3 set up its own bookkeeping (render_mode, reset_infos, ...)
            try:
                _VecEnvBase.__init__(self, self.num_envs, self.observation_space, self.action_space)
            except TypeError:
                pass
        self._steps_done = 0
        self._closed = False
        self._handle = None
        self.terminal_obs = None
        if io == "torch":
            self.batch = EnvBatch(self.kind, self.num_envs, device=device, env_id0=self.env_id_offset, seed=seed,
                                  auto_reset=auto_reset, r=self.cfg["r"], K=self.cfg["K"], sigma=self.cfg["sigma"],
                                  C=self.cfg["C"], x0=self.cfg["x0"], Tmax=self.cfg["Tmax"],
                                  n_actions=self.cfg["n_actions"], sigma_p=self.cfg["sigma_p"],
                                  **growth_kwargs(self.kind, env_kwargs))
            self.device = self.batch.device
            torch = self.batch.torch
            # two sets of output buffers used alternately: what step() returns stays valid until the next-but-one call
            self._obs2 = [self.batch.state, torch.empty_like(self.batch.state)]
            self._reward2 = [torch.empty(self.num_envs, dtype=torch.float64, device=self.device) for _ in range(2)]
            self._done2 = [torch.empty(self.num_envs, dtype=torch.uint8, device=self.device) for _ in range(2)]
            self._flip = 0
            self._act_dtype = torch.int64 if self.kind == nat.V0 else torch.float64
            self.terminal_obs = torch.full((self.num_envs,), float("nan"), dtype=torch.float64, device=self.device)
            # what the step kernel keeps besides the reference's outputs (fishing_step_ex): episode statistics and float32
            # copies of observations / rewards -- one prepared argument block per output set
            self._extra = None
            if self.monitor:
                self.ep_return = torch.zeros(self.num_envs, dtype=torch.float64, device=self.device)
                self.last_ep_return = torch.zeros_like(self.ep_return)
                self.last_ep_length = torch.zeros(self.num_envs, dtype=torch.int32, device=self.device)
            if self.sb3_dtypes:
                self._obs32 = [torch.empty(self.num_envs, dtype=torch.float32, device=self.device) for _ in range(2)]
                self._reward32 = [torch.empty(self.num_envs, dtype=torch.float32, device=self.device) for _ in range(2)]
            if self.monitor or self.sb3_dtypes:
                self._extra = []
                for k in range(2):
                    x = nat.FishingStepExtra()
                    if self.monitor:
                        x.ep_return, x.last_ep_return = self.ep_return.data_ptr(), self.last_ep_return.data_ptr()
                        x.last_ep_length = self.last_ep_length.data_ptr()
                    if self.sb3_dtypes:
                        x.out_obs32, x.out_reward32 = self._obs32[k].data_ptr(), self._reward32[k].data_ptr()
                    self._extra.append(x)
        else:
            import torch

            dev = torch.device(device)
            self.device = dev
            self._dev_index = dev.index if dev.index is not None else torch.cuda.current_device()
            self._create_handle()

    # ------------------------------------------------------------------------------------------------ numpy mode
    def _params(self):
        c = self.cfg
        return nat.make_params(self.kind, r=c["r"], K=c["K"], sigma=c["sigma"], C_=c["C"], x0=c["x0"], Tmax=c["Tmax"],
                               n_actions=c["n_actions"], sigma_p=c["sigma_p"], seed=self._seed,
                               flags=nat.AUTO_RESET if self.auto_reset else 0, **self._growth)

    def _create_handle(self):
        h = ct.c_void_p()
        p = self._params()
        nat.check(nat.lib().fishing_vec_create(ct.byref(p), self.num_envs, self.env_id_offset, self._dev_index,
                                               ct.byref(h)), "fishing_vec_create")
        self._handle = h
        n = self.num_envs
        self._h_sets = [(_PinnedArray(n, np.float64), _PinnedArray(n, np.float64), _PinnedArray(n, np.uint8)) for _ in range(2)]
        self._h_flip = 0
        self._h_obs, self._h_reward, self._h_done = self._h_sets[0]
        self._h_sets32 = [(_PinnedArray(n, np.float32), _PinnedArray(n, np.float32), _PinnedArray(n, np.uint8)) for _ in range(2)] \
            if self.sb3_dtypes else []
        self._h_term = _PinnedArray(n, np.float64)
        self._h_act = _PinnedArray(n, np.float64)  # staging for pageable / wrongly-typed actions (8 bytes per env)

    # ------------------------------------------------------------------------------------------------ VecEnv API
    def reset(self):
        """-> observations [num_envs, 1] float64 (BaseFishingEnv.reset, base_fishing_env.py:83-91; v4 redraws K, r)."""
        self._check_open()
        if self.io == "torch":
            self.batch.state = self._obs2[self._flip]
            s = self.batch.reset()
            if self.monitor:
                self.ep_return.zero_()
            return s.view(self.num_envs, 1).float() if self.sb3_dtypes else s.view(self.num_envs, 1)
        nat.check(nat.lib().fishing_vec_reset(self._handle, self._h_obs.array.ctypes.data), "fishing_vec_reset")
        obs = self._h_obs.array.reshape(self.num_envs, 1)
        return obs.astype(np.float32) if self.sb3_dtypes else obs

    def step_async(self, actions):
        self._pending = actions

    def step_wait(self):
        actions, self._pending = self._pending, None
        if actions is None:
            raise RuntimeError("step_wait() without step_async()")
        return self._step(actions)

    def step(self, actions):
        """-> (obs [n,1] float64, rewards [n] float64, dones [n] bool, infos)  -- BaseFishingEnv.step
        (base_fishing_env.py:60-81) for every environment, with SB3-style auto-reset when enabled."""
        return self._step(actions)

    def _step(self, actions):
        self._check_open()
        if self.io == "torch":
            return self._step_torch(actions)
        return self._step_numpy(actions)

    def _coerce_torch_actions(self, actions):
        torch = self.batch.torch
        if (torch.is_tensor(actions) and actions.dtype == self._act_dtype and actions.device == self.device
                and actions.is_contiguous() and actions.numel() == self.num_envs):
            return actions.view(self.num_envs)      # the common case: nothing to convert
        if not torch.is_tensor(actions):
            actions = torch.as_tensor(np.asarray(actions), device=self.device)
        if actions.device != self.device:
            actions = actions.to(self.device)
        actions = actions.reshape(self.num_envs)
        if self.kind == nat.V0:
            if actions.dtype not in (torch.int32, torch.int64):
                actions = actions.to(torch.int64)
        elif actions.dtype not in (torch.float64, torch.float32):
            actions = actions.to(torch.float64)
        return actions.contiguous()

    def _step_torch(self, actions):
        b = self.batch
        a = self._coerce_torch_actions(actions)
        self._flip ^= 1
        k = self._flip
        obs, rew, done = b.step_trusted(a, self._obs2[k], self._reward2[k], self._done2[k],
                                        self.terminal_obs if self.auto_reset else None,
                                        self._extra[k] if self._extra is not None else None)
        done_b = done.view(b.torch.bool)
        episode = (self.last_ep_return, self.last_ep_length) if self.monitor else None
        infos = LazyInfos(done, self.terminal_obs, episode)
        if self.sb3_dtypes:
            return self._obs32[k].view(self.num_envs, 1), self._reward32[k], done_b, infos
        return obs.view(self.num_envs, 1), rew, done_b, infos

    def step_into(self, actions, out_obs, out_reward, out_done, out_terminal=None, noise=None):
        """torch mode, rollout-buffer form: write this step's observations / rewards / dones straight into the given
        (contiguous, 1-D, length num_envs) slices, e.g. buf_obs[t], buf_reward[t], buf_done[t].  No allocation, no
        copy: the observation slice becomes the engine's state for the next step."""
        self._check_open()
        a = self._coerce_torch_actions(actions)
        self.batch.step(a, noise=noise, out_obs=out_obs.view(-1), out_reward=out_reward.view(-1),
                        out_done=out_done.view(-1), out_terminal=out_terminal)

    def _step_numpy(self, actions):
        n = self.num_envs
        a = np.asarray(actions)
        if self.kind == nat.V0:
            if a.dtype == np.int32:
                adt = nat.ACT_I32
            else:
                a = a.astype(np.int64, copy=False)
                adt = nat.ACT_I64
        else:
            if a.dtype == np.float32:
                adt = nat.ACT_F32
            else:
                a = a.astype(np.float64, copy=False)
                adt = nat.ACT_F64
        a = np.ascontiguousarray(a.reshape(n))
        self._h_flip ^= 1
        if self.sb3_dtypes:
            o32, r32, d8 = self._h_sets32[self._h_flip]
            rc = nat.lib().fishing_vec_step_many_f32(self._handle, 1, a.ctypes.data, adt, o32.array.ctypes.data,
                                                     r32.array.ctypes.data, d8.array.ctypes.data)
            nat.check(rc, "fishing_vec_step_many_f32")
            self._steps_done += 1
            done = d8.array.view(np.bool_)
            return (o32.array.reshape(n, 1), r32.array, done, LazyInfos(done, _LazyTerminal(self) if self.auto_reset else None))
        self._h_obs, self._h_reward, self._h_done = self._h_sets[self._h_flip]
        if n <= _ZC_MAX:   # small batch: stage the actions in page-locked memory so that the library can take its one-launch route
            staged = self._h_act.array.view(a.dtype)[:n]
            np.copyto(staged, a)
            a = staged
        rc = nat.lib().fishing_vec_step(self._handle, a.ctypes.data, adt, None, self._h_obs.array.ctypes.data,
                                        self._h_reward.array.ctypes.data, self._h_done.array.ctypes.data, None)
        nat.check(rc, "fishing_vec_step")
        self._steps_done += 1
        done = self._h_done.array.view(np.bool_)
        return (self._h_obs.array.reshape(n, 1), self._h_reward.array, done,
                LazyInfos(done, _LazyTerminal(self) if self.auto_reset else None))

    def step_many(self, actions, out=None):
        """T = actions.shape[0] lock-steps in one call (the loop of simulate_mdp_vec with a pre-computed action table).
        numpy mode: actions [T, n] host array -> (obs [T, n], rewards [T, n], dones [T, n]) host arrays, with the copies
        and kernels of consecutive steps pipelined (fishing_vec_step_many).  `out` may supply the three output arrays
        (e.g. page-locked ones).  torch mode: CUDA tensors in and out (fishing_step_many)."""
        self._check_open()
        n = self.num_envs
        if self.io == "torch":
            torch = self.batch.torch
            a = actions if torch.is_tensor(actions) else torch.as_tensor(np.asarray(actions), device=self.device)
            T = a.shape[0]
            a = a.reshape(T, n).contiguous()
            if out is None:
                out = (torch.empty((T, n), dtype=torch.float64, device=self.device),
                       torch.empty((T, n), dtype=torch.float64, device=self.device),
                       torch.empty((T, n), dtype=torch.uint8, device=self.device))
            self.batch.step_many(a, out[0], out[1], out[2])
            return out[0], out[1], out[2].view(torch.bool)
        a = np.asarray(actions)
        T = a.shape[0]
        if self.kind == nat.V0:
            adt = nat.ACT_I32 if a.dtype == np.int32 else nat.ACT_I64
            a = a if a.dtype == np.int32 else a.astype(np.int64, copy=False)
        else:
            adt = nat.ACT_F32 if a.dtype == np.float32 else nat.ACT_F64
            a = a if a.dtype == np.float32 else a.astype(np.float64, copy=False)
        a = np.ascontiguousarray(a.reshape(T, n))
        fdt = np.float32 if self.sb3_dtypes else np.float64
        done_shape = (T, (n + 7) // 8) if self.packed_done else (T, n)
        if out is None:
            out = (np.empty((T, n), dtype=fdt), np.empty((T, n), dtype=fdt), np.empty(done_shape, dtype=np.uint8))
        obs, rew, done = out
        if obs.dtype != fdt or rew.dtype != fdt:
            raise TypeError(f"output arrays must be {np.dtype(fdt).name}")
        if done.shape != done_shape or done.dtype != np.uint8:
            raise ValueError(f"the done array must be uint8 {done_shape}" + (" (bit-packed)" if self.packed_done else ""))
        wire = ((nat.WIRE_OBS_F32 | nat.WIRE_REWARD_F32) if self.sb3_dtypes else 0) | (nat.WIRE_DONE_BITS if self.packed_done else 0)
        if wire:
            rc = nat.lib().fishing_vec_step_many_wire(self._handle, T, a.ctypes.data, adt, obs.ctypes.data, rew.ctypes.data,
                                                      done.ctypes.data, wire)
        else:
            rc = nat.lib().fishing_vec_step_many(self._handle, T, a.ctypes.data, adt, obs.ctypes.data, rew.ctypes.data,
                                                 done.ctypes.data)
        nat.check(rc, "fishing_vec_step_many")
        self._steps_done += T
        if self.packed_done:      # [T, ceil(n/8)] bytes; FishingVecEnv.unpack_dones(done) gives the [T, n] bools
            return obs, rew, done
        return obs, rew, done.view(np.bool_)

    def unpack_dones(self, packed):
        """bit-packed done rows (packed_done=True) -> bool array [..., num_envs]."""
        return np.unpackbits(np.asarray(packed, dtype=np.uint8), axis=-1, count=self.num_envs, bitorder="little").view(np.bool_)

    # ------------------------------------------------------------------------------------------------ attributes
    def _indices(self, indices):
        if indices is None:
            return range(self.num_envs)
        if isinstance(indices, int):
            return [indices]
        return list(indices)

    _ARRAYS = {"state": (nat.ARRAY_STATE, np.float64), "years_passed": (nat.ARRAY_YEARS, np.int32),
               "K": (nat.ARRAY_K, np.float64), "r": (nat.ARRAY_R, np.float64), "model": (nat.ARRAY_MODEL, np.int32)}

    def _per_env_array(self, attr_name):
        """host copy of one per-environment array (numpy), whichever the io mode."""
        if self.io == "torch":
            b = self.batch
            t = {"state": b.state, "years_passed": b.years, "K": b.K_env, "r": b.r_env, "model": b.model}[attr_name]
            return None if t is None else t.cpu().numpy()
        which, dt = self._ARRAYS[attr_name]
        out = np.empty(self.num_envs, dtype=dt)
        rc = nat.lib().fishing_vec_fetch_array(self._handle, which, out.ctypes.data)
        if rc == -2:            # FISHING_E_KIND: this environment id has no such array
            return None
        nat.check(rc, "fishing_vec_fetch_array")
        return out

    def _store_per_env_array(self, attr_name, arr):
        if self.io == "torch":
            b = self.batch
            t = {"state": b.state, "years_passed": b.years, "K": b.K_env, "r": b.r_env, "model": b.model}[attr_name]
            t.copy_(b.torch.as_tensor(arr, device=self.device).to(t.dtype))
            return
        which, dt = self._ARRAYS[attr_name]
        arr = np.ascontiguousarray(arr, dtype=dt)
        nat.check(nat.lib().fishing_vec_store_array(self._handle, which, arr.ctypes.data), "fishing_vec_store_array")

    def get_attr(self, attr_name, indices=None):
        """list of per-env values (shared_env.py:73 calls env.get_attr("Tmax")[0])."""
        idx = self._indices(indices)
        c = self.cfg
        shared = {"Tmax": c["Tmax"], "sigma": c["sigma"], "init_state": c["x0"], "n_actions": c["n_actions"], "C": c["C"],
                  "sigma_p": c["sigma_p"], "K_mean": c["K"], "r_mean": c["r"]}
        if attr_name in shared:
            return [shared[attr_name] for _ in idx]
        if attr_name in ("K", "r"):
            if self.kind != nat.V4:
                return [c[attr_name] for _ in idx]
            t = self._per_env_array(attr_name)       # fishing-v4: the parameters every environment drew at its last reset
            return [float(t[i]) for i in idx]
        if attr_name in ("state", "years_passed", "model"):
            t = self._per_env_array(attr_name)
            if t is None:
                raise AttributeError(attr_name)
            return [np.array([t[i]]) if attr_name == "state" else int(t[i]) for i in idx]
        if attr_name == "fish_population":
            s, K = self._per_env_array("state"), self.get_attr("K", idx)
            return [float((s[i] + 1) * K[j]) for j, i in enumerate(idx)]
        raise AttributeError(attr_name)

    def set_attr(self, attr_name, value, indices=None):
        """Constructor parameters (r, K, sigma, C, Tmax, init_state, ...) are ONE value for the whole batch: they can be set
        for all environments (indices=None or every index); per-environment arrays (state, years_passed, fishing-v4's K and
        r, fishing-v11's model) can be set for any subset."""
        key = {"init_state": "x0", "K_mean": "K", "r_mean": "r"}.get(attr_name, attr_name)
        per_env = attr_name in ("state", "years_passed", "model") or (attr_name in ("K", "r") and self.kind == nat.V4)
        if per_env:
            idx = list(self._indices(indices))
            arr = self._per_env_array(attr_name)
            if arr is None:
                raise AttributeError(attr_name)
            vals = np.broadcast_to(np.asarray(value, dtype=arr.dtype).reshape(-1) if np.ndim(value) else np.asarray(value, dtype=arr.dtype),
                                   (len(idx),))
            arr[idx] = vals
            self._store_per_env_array(attr_name, arr)
            return
        if indices is not None and sorted(set(self._indices(indices))) != list(range(self.num_envs)):
            raise ValueError(f"{attr_name!r} is one value for the whole batch (the kernels read it as a launch parameter); "
                             "set it with indices=None, or use separate FishingVecEnv objects for separate parameter sets")
        if key not in self.cfg:
            raise AttributeError(attr_name)
        self.cfg[key] = value
        if self.io == "torch":
            self.batch.cfg["C_" if key == "C" else key] = value
        else:
            p = self._params()
            nat.check(nat.lib().fishing_vec_set_params(self._handle, ct.byref(p)), "fishing_vec_set_params")

    def env_method(self, method_name, *method_args, indices=None, **method_kwargs):
        """Only the pure unit conversions make sense on a batch (shared_env.py:17-19 calls get_fish_population)."""
        idx = self._indices(indices)
        K = self.get_attr("K", idx)
        n_act = self.cfg["n_actions"]
        out = []
        for j, _ in enumerate(idx):
            if method_name == "get_fish_population":
                out.append(np.array([(np.asarray(method_args[0], dtype=np.float64).reshape(-1)[0] + 1) * K[j]]))
            elif method_name == "get_state":
                out.append(np.array([method_args[0] / K[j] - 1]))
            elif method_name == "get_quota":
                a = method_args[0]
                out.append((a / n_act) * K[j] if self.kind == nat.V0 else
                           (float(np.clip(np.asarray(a, dtype=np.float64).reshape(-1)[0], -1, 1)) + 1) * K[j])
            elif method_name == "get_action":
                q = method_args[0]
                out.append(round(q * n_act / K[j]) if self.kind == nat.V0 else q / K[j] - 1)
            else:
                raise AttributeError(f"env_method({method_name!r}) is not available on the batched environment")
        return out

    def env_is_wrapped(self, wrapper_class, indices=None):
        return [False for _ in self._indices(indices)]

    def seed(self, seed=None):
        """Re-key the Philox noise streams (the reference has no seed(), base_fishing_env.py:13)."""
        self._seed = 0 if seed is None else int(seed)
        if self.io == "torch":
            self.batch.cfg["seed"] = self._seed
        else:
            self.close()
            self._closed = False
            self._create_handle()
        return [self._seed + i for i in range(min(self.num_envs, 8))]

    def render(self, mode="human"):
        """The reference's render() appends one row [years_passed, state, ...] per step of its single environment
        (base_fishing_env.py:93-98).  For the batch: the table of all environments, columns years_passed, state (the
        normalised observation) and fish_population -- one device-to-host copy of two arrays, nothing is drawn."""
        s = self._per_env_array("state")
        y = self._per_env_array("years_passed")
        K = np.asarray(self.get_attr("K"), dtype=np.float64)
        return np.column_stack([y.astype(np.float64), s, (s + 1) * K])

    def _check_open(self):
        if self._closed:
            raise RuntimeError("environment is closed")

    def close(self):
        if self._closed:
            return
        self._closed = True
        if self._handle is not None:
            nat.lib().fishing_vec_destroy(self._handle)
            self._handle = None
            for a in (*self._h_sets[0], *self._h_sets[1], *(x for st in self._h_sets32 for x in st), self._h_term, self._h_act):
                a.free()

    def __del__(self):  # pragma: no cover
        try:
            self.close()
        except Exception:  # noqa: BLE001
            pass
