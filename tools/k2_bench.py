#!/usr/bin/env python
"""Kernel-only timings of the fused rollout kernels (K2 / sweep) and the K1 lock-step loop, for A/B runs of kernel
variants (FISHING_B200_LIB selects the build).  Prints one JSON line; includes an fp64 checksum of the returns so that
variants can be compared for bit-identity."""
from __future__ import annotations

import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

from gym_fishing_b200 import _native as nat  # noqa: E402
from gym_fishing_b200.engine import EnvBatch  # noqa: E402


def timed(fn, reps=5, warm=2):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    best = 1e30
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        out = fn()
        e1.record()
        torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    return best, out


def main():
    dev = torch.device("cuda", 0)
    n = int(os.environ.get("K2B_N", 1 << 20))
    T = int(os.environ.get("K2B_T", 1000))
    res = {"lib": os.path.basename(nat.LIB_PATH), "n": n, "T": T}

    def fused(kind, policy, sigma, S=None, n_=n, levels=0, **kw):
        b = EnvBatch(kind, n_, device=dev, r=0.3, K=kw.pop("K", 1.0), sigma=sigma, x0=0.75, Tmax=kw.pop("Tmax", 100), seed=0,
                     C=0.5, **kw)
        ret = torch.empty(n_, dtype=torch.float64, device=dev)
        ds = torch.empty(n_, dtype=torch.int32, device=dev)
        S_env = None
        if levels:
            S_env = (torch.arange(n_, device=dev) % levels).to(torch.float64) / (levels - 1)

        def run():
            b.reset()
            return b.rollout(T, policy=policy, S=None if levels else S, S_env=S_env, noise_id_div=levels, out_return=ret,
                             out_done_step=ds, freeze_done=kw.get("freeze", False))
        ms, out = timed(run)
        return {"ms": round(ms, 4), "env_steps_per_s": n_ * T / (ms * 1e-3), "checksum": float(out["ret"].sum()),
                "ret_hash": int(out["ret"].view(torch.int64).sum())}

    res["v1_escapement_sigma0.01"] = fused(nat.V1, "escapement", 0.01, S=0.4996)
    res["v1_escapement_sigma0.01_autoreset"] = fused(nat.V1, "escapement", 0.01, S=0.4996, auto_reset=True)
    res["v1_escapement_sigma0"] = fused(nat.V1, "escapement", 0.0, S=0.4996)
    res["v2_escapement_sigma0.05_T1000"] = fused(nat.V2, "escapement", 0.05, S=float.fromhex("0x1.9374bcp-1"), Tmax=1000)
    res["v1_K2_escapement_sigma0.01"] = fused(nat.V1, "escapement", 0.01, S=0.9, K=2.0)
    res["v1_sweep1024_sigma0.1"] = fused(nat.V1, "escapement", 0.1, n_=1024 * 1024, levels=1024, Tmax=1000)

    # K1 lock-step loop issued from C
    Tk = 200
    b = EnvBatch(nat.V1, n, device=dev, auto_reset=True, r=0.3, K=1.0, sigma=0.01, x0=0.75, Tmax=100, seed=0)
    acts = torch.rand(Tk, n, dtype=torch.float64, device=dev) * 2 - 1
    obs = torch.empty(Tk, n, dtype=torch.float64, device=dev)
    rew = torch.empty_like(obs)
    done = torch.empty(Tk, n, dtype=torch.uint8, device=dev)

    def k1():
        b.step_many(acts, obs, rew, done)
    b.reset()
    ms, _ = timed(k1)
    res["k1_v1_us_per_launch"] = round(ms * 1e3 / Tk, 3)
    res["k1_hash"] = int((obs[-1].view(torch.int64) >> 11).sum())
    # the same loop without process noise (no Philox, no Box-Muller): what the lock-step costs apart from the draws
    b = EnvBatch(nat.V1, n, device=dev, auto_reset=True, r=0.3, K=1.0, sigma=0.0, x0=0.75, Tmax=100, seed=0)
    b.reset()
    ms, _ = timed(k1)
    res["k1_v1_sigma0_us_per_launch"] = round(ms * 1e3 / Tk, 3)
    print(json.dumps(res))


if __name__ == "__main__":
    main()
