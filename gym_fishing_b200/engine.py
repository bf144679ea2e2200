"""Thin torch-tensor wrappers over the C ABI (include/fishing_b200.h).  torch is used only for device memory and
streams: every computation happens in libfishing_b200.so's CUDA kernels, reached through ctypes with raw
pointers.  Nothing here falls back to the CPU.
"""
from __future__ import annotations

import ctypes as ct

from . import _native as nat

_ACT_OF_DTYPE = None


def _act_dtype(t):
    import torch

    global _ACT_OF_DTYPE
    if _ACT_OF_DTYPE is None:
        _ACT_OF_DTYPE = {torch.float64: nat.ACT_F64, torch.float32: nat.ACT_F32, torch.int32: nat.ACT_I32,
                         torch.int64: nat.ACT_I64}
    try:
        return _ACT_OF_DTYPE[t.dtype]
    except KeyError:
        raise TypeError(f"unsupported action dtype {t.dtype}; use float64/float32 (Box) or int32/int64 (Discrete)")


def _ptr(t):
    return None if t is None else t.data_ptr()


def _stream(device):
    import torch

    return torch.cuda.current_stream(device).cuda_stream


def _raw_stream_fn():
    """torch's direct accessor of the current stream's handle (a C call, ~0.3 us) when this torch has it; the public
    route through torch.cuda.current_stream() builds a Stream object and costs ~4 us per step() call."""
    import torch

    fn = getattr(torch._C, "_cuda_getCurrentRawStream", None)
    if fn is None:
        return lambda idx: torch.cuda.current_stream(idx).cuda_stream
    return fn


def _chk_dev(t, name, dtype=None, n=None):
    if not t.is_cuda:
        raise ValueError(f"{name} must be a CUDA tensor (the engine has no CPU path)")
    if not t.is_contiguous():
        raise ValueError(f"{name} must be contiguous")
    if dtype is not None and t.dtype != dtype:
        raise TypeError(f"{name} must be {dtype}, got {t.dtype}")
    if n is not None and t.numel() != n:
        raise ValueError(f"{name} must have {n} elements, got {t.numel()}")


class _Cfg(dict):
    """dict that counts its mutations, so that the ctypes parameter block is rebuilt only when something changed."""

    version = 0

    def __setitem__(self, k, v):
        self.version += 1
        super().__setitem__(k, v)

    def update(self, *a, **kw):
        self.version += 1
        super().update(*a, **kw)


class EnvBatch:
    """Structure-of-arrays state of `n` environments on one GPU plus the kernels that advance it.

    env_id0 is the global id of the first environment: Philox noise is keyed by global id, so sharding a batch over
    GPUs (gym_fishing_b200.distributed) does not change any trajectory.
    """

    def __init__(self, kind, n, *, device="cuda", env_id0=0, seed=0, auto_reset=False, r=0.3, K=1.0, sigma=0.0, C=0.5,
                 x0=0.75, Tmax=100, n_actions=100, sigma_p=0.1, **growth):
        """**growth: extra entries of the reference's `params` dict for fishing-v5..v11 (growth_K, M, theta, q, b, a, alpha,
        models=[(name, params), ...])."""
        torch = nat.require_cuda()
        nat.lib()
        self.torch = torch
        self._lib = nat.lib()
        self._raw_stream = _raw_stream_fn()
        self.kind = int(kind)
        self.n = int(n)
        self.device = torch.device(device)
        if self.device.type != "cuda":
            raise ValueError("device must be a CUDA device")
        if self.device.index is None:
            self.device = torch.device("cuda", torch.cuda.current_device())
        self.env_id0 = int(env_id0)
        self.step_idx = 0
        self.epoch = 0
        self.cfg = _Cfg(r=r, K=K, sigma=sigma, C_=C, x0=x0, Tmax=Tmax, n_actions=n_actions, sigma_p=sigma_p, seed=seed)
        self.cfg.update(growth)
        self._p_cache = (None, None, None)
        self.auto_reset = bool(auto_reset)
        f64 = dict(dtype=torch.float64, device=self.device)
        self.state = torch.empty(self.n, **f64)
        self.years = torch.zeros(self.n, dtype=torch.int32, device=self.device)
        self.K_env = torch.empty(self.n, **f64) if self.kind == nat.V4 else None
        self.r_env = torch.empty(self.n, **f64) if self.kind == nat.V4 else None
        self.model = torch.zeros(self.n, dtype=torch.int32, device=self.device) if self.kind == nat.V11 else None
        self._scratch = None

    # ------------------------------------------------------------------------------------------------ plumbing
    def params(self, **override):
        if not override:
            ver, auto, p = self._p_cache
            if ver == self.cfg.version and auto == self.auto_reset:
                return p
        cfg = dict(self.cfg)
        cfg.update(override)
        flags = cfg.pop("flags", nat.AUTO_RESET if self.auto_reset else 0)
        p = nat.make_params(self.kind, flags=flags, **cfg)
        if not override:
            self._p_cache = (self.cfg.version, self.auto_reset, p)
        return p

    def soa(self, state=None):
        s = self.state if state is None else state
        return nat.FishingSoA(s.data_ptr(), self.years.data_ptr(), _ptr(self.K_env), _ptr(self.r_env), _ptr(self.model))

    def _drift(self, steps):
        """fishing-v10: r <- r + alpha before every draw (growth_models.py:147-150); sequential adds like the reference."""
        if self.kind == nat.V10:
            r, alpha = float(self.cfg["r"]), float(self.cfg.get("alpha", 0.0))
            for _ in range(int(steps)):
                r = r + alpha
            self.cfg["r"] = r

    # ------------------------------------------------------------------------------------------------ kernels
    def reset(self, mask=None, reset_noise=None, out_obs=None, reset_model=None):
        """fishing_reset.  mask: uint8/bool [n] selecting environments; reset_noise: [2, n] (zK, zr) for v4;
        reset_model: int32 [n] injected model indices for v11."""
        torch = self.torch
        if reset_model is not None:
            _chk_dev(reset_model, "reset_model", torch.int32, self.n)
            if mask is not None:
                mask = mask.to(torch.uint8) if mask.dtype != torch.uint8 else mask
                _chk_dev(mask, "mask", torch.uint8, self.n)
            if out_obs is not None:
                _chk_dev(out_obs, "out_obs", torch.float64, self.n)
            p = self.params()
            with torch.cuda.device(self.device):
                rc = nat.lib().fishing_reset_models(ct.byref(p), self.n, self.env_id0, self.epoch, _ptr(mask),
                                                    reset_model.data_ptr(), self.soa(), _ptr(out_obs), _stream(self.device))
            nat.check(rc, "fishing_reset_models")
            self.epoch += 1
            return self.state
        if mask is not None:
            mask = mask.to(torch.uint8) if mask.dtype != torch.uint8 else mask
            _chk_dev(mask, "mask", torch.uint8, self.n)
        if reset_noise is not None:
            _chk_dev(reset_noise, "reset_noise", torch.float64, 2 * self.n)
        if out_obs is not None:
            _chk_dev(out_obs, "out_obs", torch.float64, self.n)
        p = self.params()
        with torch.cuda.device(self.device):
            rc = nat.lib().fishing_reset(ct.byref(p), self.n, self.env_id0, self.epoch, _ptr(mask), _ptr(reset_noise),
                                         self.soa(), _ptr(out_obs), _stream(self.device))
        nat.check(rc, "fishing_reset")
        self.epoch += 1
        return self.state

    def step(self, action, *, noise=None, reset_noise=None, out_obs=None, out_reward=None, out_done=None,
             out_terminal=None, out_stock=None, out_harvest=None, state_in=None):
        """fishing_step: one lock-step of all n environments.

        out_obs defaults to the state array itself (in place).  When out_obs is another buffer (e.g. slice t of a
        rollout buffer) pass it as `state_in` of the next call, or leave state_in=None to have the engine keep
        `self.state` pointed at the most recent observations without a copy."""
        torch = self.torch
        _chk_dev(action, "action", None, self.n)
        adt = _act_dtype(action)
        src = self.state if state_in is None else state_in
        if out_obs is None:
            out_obs = src
        if out_reward is None:
            out_reward = torch.empty(self.n, dtype=torch.float64, device=self.device)
        if out_done is None:
            out_done = torch.empty(self.n, dtype=torch.uint8, device=self.device)
        for nm, t, dt in (("out_obs", out_obs, torch.float64), ("out_reward", out_reward, torch.float64),
                          ("out_done", out_done, torch.uint8)):
            _chk_dev(t, nm, dt, self.n)
        if noise is not None:
            _chk_dev(noise, "noise", torch.float64, self.n)
        if reset_noise is not None:
            _chk_dev(reset_noise, "reset_noise", torch.float64, 2 * self.n)
        if out_terminal is not None:
            _chk_dev(out_terminal, "out_terminal", torch.float64, self.n)
        # the optional arrays reach the kernel as raw pointers too: same device / dtype / length checks
        if out_stock is not None:
            _chk_dev(out_stock, "out_stock", torch.float64, self.n)
        if out_harvest is not None:
            _chk_dev(out_harvest, "out_harvest", torch.float64, self.n)
        if state_in is not None:
            _chk_dev(state_in, "state_in", torch.float64, self.n)
        p = self.params()
        if torch.cuda.current_device() != self.device.index:
            torch.cuda.set_device(self.device)
        rc = self._lib.fishing_step(ct.byref(p), self.n, self.env_id0, self.step_idx, action.data_ptr(), adt,
                                    _ptr(noise), _ptr(reset_noise), self.soa(src), out_obs.data_ptr(),
                                    out_reward.data_ptr(), out_done.data_ptr(), _ptr(out_terminal),
                                    _ptr(out_stock), _ptr(out_harvest), self._raw_stream(self.device.index))
        if rc:
            nat.check(rc, "fishing_step")
        self._drift(1)
        self.step_idx += 1
        self.state = out_obs  # the newest observations ARE the state (no copy)
        return out_obs, out_reward, out_done

    def step_trusted(self, action, out_obs, out_reward, out_done, out_terminal=None, extra=None):
        """step() for a caller that owns correctly shaped, typed, contiguous CUDA buffers (FishingVecEnv's alternating
        output sets; `action` already coerced): no argument validation, no default allocations -- about 4 us less Python
        per call.  extra: a prepared _native.FishingStepExtra (episode statistics / float32 copies, fishing_step_ex)."""
        src = self.state
        p = self.params()
        if self.torch.cuda.current_device() != self.device.index:
            self.torch.cuda.set_device(self.device)
        if extra is None:
            rc = self._lib.fishing_step(ct.byref(p), self.n, self.env_id0, self.step_idx, action.data_ptr(), _act_dtype(action),
                                        None, None, self.soa(src), out_obs.data_ptr(), out_reward.data_ptr(), out_done.data_ptr(),
                                        _ptr(out_terminal), None, None, self._raw_stream(self.device.index))
        else:
            rc = self._lib.fishing_step_ex(ct.byref(p), self.n, self.env_id0, self.step_idx, action.data_ptr(), _act_dtype(action),
                                           None, None, self.soa(src), out_obs.data_ptr(), out_reward.data_ptr(),
                                           out_done.data_ptr(), _ptr(out_terminal), ct.byref(extra),
                                           self._raw_stream(self.device.index))
        if rc:
            nat.check(rc, "fishing_step")
        self._drift(1)
        self.step_idx += 1
        self.state = out_obs
        return out_obs, out_reward, out_done

    def step_many(self, actions, out_obs, out_reward, out_done, *, noise=None, out_terminal=None):
        """fishing_step_many: T = actions.shape[0] lock-steps into rollout buffers [T, n] (one launch per step, issued
        from C; capturable in a CUDA graph).  self.state keeps its own storage and holds the final state."""
        torch = self.torch
        T = actions.shape[0]
        _chk_dev(actions, "actions", None, T * self.n)
        adt = _act_dtype(actions)
        _chk_dev(out_obs, "out_obs", torch.float64, T * self.n)
        _chk_dev(out_reward, "out_reward", torch.float64, T * self.n)
        _chk_dev(out_done, "out_done", torch.uint8, T * self.n)
        if noise is not None:
            _chk_dev(noise, "noise", torch.float64, T * self.n)
        if out_terminal is not None:
            _chk_dev(out_terminal, "out_terminal", torch.float64, T * self.n)
        p = self.params()
        with torch.cuda.device(self.device):
            rc = nat.lib().fishing_step_many(ct.byref(p), self.n, self.env_id0, self.step_idx, T, actions.data_ptr(), adt,
                                             _ptr(noise), self.soa(), out_obs.data_ptr(), out_reward.data_ptr(),
                                             out_done.data_ptr(), _ptr(out_terminal), _stream(self.device))
        nat.check(rc, "fishing_step_many")
        self._drift(T)
        self.step_idx += T

    def noise(self, T, *, stream=nat.STREAM_STEP, step0=None):
        """fishing_noise_fill: the engine's own Philox/Box-Muller draws, [T, n] (or [T, 2, n] for the v4 streams)."""
        torch = self.torch
        shape = (T, self.n) if stream == nat.STREAM_STEP else (T, 2, self.n)
        out = torch.empty(shape, dtype=torch.float64, device=self.device)
        step0 = self.step_idx if step0 is None else step0
        with torch.cuda.device(self.device):
            rc = nat.lib().fishing_noise_fill(self.cfg["seed"] & 0xFFFFFFFFFFFFFFFF, stream, self.n, self.env_id0, step0, T,
                                              out.data_ptr(), _stream(self.device))
        nat.check(rc, "fishing_noise_fill")
        return out

    def rollout(self, T, *, policy="escapement", S=None, S_env=None, msy_q=None, actions=None, noise=None,
                noise_id_div=0, freeze_done=False, trace=False, trace_action=False, out_return=None,
                out_done_step=None):
        """fishing_rollout: T fused steps.  Returns dict(ret, done_step[, obs, reward, done, action])."""
        torch = self.torch
        if self.n % 2 or self.env_id0 % 2:
            raise ValueError("the fused rollout kernel owns environments in pairs: n and env_id0 must be even "
                             f"(got n={self.n}, env_id0={self.env_id0}); use step()/step_many() for odd batches")
        if noise_id_div and noise_id_div > 1 and (self.n % (2 * noise_id_div) or self.env_id0 % (2 * noise_id_div)):
            raise ValueError("with noise_id_div = m the batch must hold whole pairs of noise ids: n and env_id0 must be "
                             f"multiples of 2*m (got n={self.n}, env_id0={self.env_id0}, m={noise_id_div})")
        a = nat.FishingRolloutArgs()
        a.T = int(T)
        a.step0 = self.step_idx
        a.noise_id_div = int(noise_id_div)
        if policy == "actions":
            a.policy = nat.POLICY_ACTIONS
            _chk_dev(actions, "actions", None, T * self.n)
            a.action_dtype = _act_dtype(actions)
            a.actions = actions.data_ptr()
        elif policy == "escapement":
            a.policy = nat.POLICY_ESCAPEMENT
            if S_env is not None:
                _chk_dev(S_env, "S_env", torch.float64, self.n)
                a.S_env = S_env.data_ptr()
            elif S is None:
                raise ValueError("escapement policy needs S or S_env")
            a.S = float(S) if S is not None else 0.0
        elif policy == "msy":
            a.policy = nat.POLICY_MSY
            a.msy_q = float(msy_q)
        else:
            raise ValueError(f"unknown policy {policy!r}")
        if noise is not None:
            _chk_dev(noise, "noise", torch.float64, T * self.n)
            a.noise = noise.data_ptr()
        f64 = dict(dtype=torch.float64, device=self.device)
        ret = torch.empty(self.n, **f64) if out_return is None else out_return
        dstep = torch.empty(self.n, dtype=torch.int32, device=self.device) if out_done_step is None else out_done_step
        a.out_return = ret.data_ptr()
        a.out_done_step = dstep.data_ptr()
        out = dict(ret=ret, done_step=dstep)
        if trace:
            out["obs"] = torch.empty((T, self.n), **f64)
            out["reward"] = torch.empty((T, self.n), **f64)
            out["done"] = torch.empty((T, self.n), dtype=torch.uint8, device=self.device)
            a.trace_obs, a.trace_reward, a.trace_done = (out[k].data_ptr() for k in ("obs", "reward", "done"))
        if trace_action:
            out["action"] = torch.empty((T, self.n), **f64)
            a.trace_action = out["action"].data_ptr()
        flags = (nat.AUTO_RESET if self.auto_reset else 0) | (nat.FREEZE_DONE if freeze_done else 0)
        p = self.params(flags=flags)
        with torch.cuda.device(self.device):
            rc = nat.lib().fishing_rollout(ct.byref(p), self.n, self.env_id0, ct.byref(a), self.soa(), _stream(self.device))
        nat.check(rc, "fishing_rollout")
        self._drift(T)
        self.step_idx += int(T)
        return out

    def reduce_stats(self, ret, done_step=None, groups=1):
        """fishing_reduce_stats -> tensor [groups, 3] = (sum, sum of squares, number done) per group (i % groups)."""
        torch = self.torch
        n = ret.numel()
        _chk_dev(ret, "ret", torch.float64)
        nbytes = nat.lib().fishing_reduce_scratch_bytes(n, groups)
        if nbytes <= 0:
            raise ValueError("n must be a positive multiple of groups")
        if self._scratch is None or self._scratch.numel() < nbytes // 8:
            self._scratch = torch.empty(nbytes // 8, dtype=torch.float64, device=self.device)
        out = torch.empty((groups, 3), dtype=torch.float64, device=self.device)
        with torch.cuda.device(self.device):
            rc = nat.lib().fishing_reduce_stats(ret.data_ptr(), _ptr(done_step), n, groups, self._scratch.data_ptr(),
                                                out.data_ptr(), _stream(self.device))
        nat.check(rc, "fishing_reduce_stats")
        return out


def bmsy(kind, *, r, K, C=0.5, K_is_f64=False):
    """fishing_bmsy: (S, msy quota) as Python floats; float32-valued for v0/v1/v2 exactly like the reference."""
    nat.require_cuda()
    S, q = C_double(), C_double()
    rc = nat.lib().fishing_bmsy(int(kind), float(r), float(K), float(C), int(bool(K_is_f64)), ct.byref(S), ct.byref(q))
    nat.check(rc, "fishing_bmsy")
    return S.value, q.value


C_double = ct.c_double
