#!/usr/bin/env python
"""Runs BASELINE.json's configs 3, 4 and 5 (bench.py covers config 2, tests cover config 1) and prints one JSON line
per config (rank 0).  Launch plainly for one GPU or under torchrun for N:

    python benchmarks/run_configs.py [--scale 1.0] [--only cfg3,cfg4,cfg5]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 benchmarks/run_configs.py

Sizes are PER BOX as in BASELINE.json (cfg3: 2^24 envs, cfg5: 1024 x 16384) divided over the ranks present; with fewer
than 8 GPUs `--per-gpu` keeps the per-GPU share of the 8-GPU run instead (weak-scaling view).  All times are CUDA-event
times on the launching stream, max over ranks."""
from __future__ import annotations

import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

import gym_fishing_b200 as gf  # noqa: E402
from gym_fishing_b200 import _native as nat  # noqa: E402
from gym_fishing_b200.distributed import barrier_max_ms, init_from_env  # noqa: E402
from gym_fishing_b200.engine import EnvBatch  # noqa: E402


def timed(fn, dev, reps=3, warm=1):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        out = fn()
    e1.record()
    torch.cuda.synchronize()
    return barrier_max_ms(e0.elapsed_time(e1) / reps, dev), out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--only", default="cfg3,cfg4,cfg5,f1")
    ap.add_argument("--per-gpu", action="store_true", help="keep the per-GPU share of the 8-GPU configuration")
    ap.add_argument("--scale", type=float, default=1.0, help="shrink every size by this factor (smoke runs)")
    args = ap.parse_args()
    rank, world, local = init_from_env()
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    only = set(args.only.split(","))
    share = world / 8.0 if args.per_gpu else 1.0
    lines = []

    if "cfg3" in only:
        # fishing-v2 tipping point, fused 1000-step constant-escapement rollout, 2^24 envs per box
        n = max(512, int((1 << 24) * share * args.scale) // 512 * 512)
        S = float.fromhex("0x1.9374bcp-1")          # reference BMSY for the default tipping model (tests/golden/kat.json)
        for sigma in (0.05, 0.0):
            ms, res = timed(lambda: gf.rollout_stats("fishing-v2", n, 1000, policy="escapement", S=S, device=dev, sigma=sigma,
                                                     Tmax=1000, r=0.3, K=1, C=0.5, init_state=0.75), dev)
            lines.append({"config": "cfg3", "env": "fishing-v2", "n_envs_total": n, "T": 1000, "sigma": sigma, "n_gpus": world,
                          "ms": ms, "env_steps_per_s": n * 1000 / (ms * 1e-3), "note": "includes reset, K4 reduction and the collective",
                          "mean_return": float(res["mean"]), "std_return": float(res["std"]), "n_extinct": float(res["n_done"])})

    if "cfg4" in only:
        # fishing-v4 feeding a torch rollout buffer: K1 writes obs/reward/done straight into [n_steps, N] slices
        n = max(1024, int((1 << 22) * args.scale) // 1024 * 1024)
        n_steps = 128
        venv = gf.FishingVecEnv("fishing-v4", num_envs=n, device=dev, seed=0, env_id_offset=rank * n, K_mean=1.0, r_mean=0.3,
                                sigma=0.0, sigma_p=0.1, init_state=0.75, Tmax=100)
        obs_buf = torch.empty(n_steps + 1, n, dtype=torch.float64, device=dev)
        rew_buf = torch.empty(n_steps, n, dtype=torch.float64, device=dev)
        done_buf = torch.empty(n_steps, n, dtype=torch.uint8, device=dev)
        act_buf = torch.empty(n_steps, n, dtype=torch.float64, device=dev)
        policy = torch.nn.Sequential(torch.nn.Linear(1, 64), torch.nn.Tanh(), torch.nn.Linear(64, 64), torch.nn.Tanh(),
                                     torch.nn.Linear(64, 1), torch.nn.Tanh()).to(dev)
        fixed_actions = torch.rand(n_steps, n, dtype=torch.float64, device=dev) * 0.4 - 1.0

        def env_only():
            obs_buf[0].copy_(venv.reset().view(-1))
            for t in range(n_steps):
                venv.step_into(fixed_actions[t], obs_buf[t + 1], rew_buf[t], done_buf[t])

        @torch.no_grad()
        def env_and_policy():
            obs_buf[0].copy_(venv.reset().view(-1))
            for t in range(n_steps):
                a = policy(obs_buf[t].to(torch.float32).unsqueeze(1)).squeeze(1)       # fp32 MLP
                act_buf[t].copy_(a)                                                    # cast to fp64
                venv.step_into(act_buf[t], obs_buf[t + 1], rew_buf[t], done_buf[t])

        def env_only_c_loop():      # same kernels issued from C with programmatic dependent launch (no python per step)
            obs_buf[0].copy_(venv.reset().view(-1))
            venv.batch.step_many(fixed_actions, obs_buf[1:], rew_buf, done_buf)

        ms_env, _ = timed(env_only, dev)
        ms_c, _ = timed(env_only_c_loop, dev)
        # the same with a policy-like action distribution (small quotas: episodes mostly end at the time limit, ~1 % of the
        # environments reset per step instead of ~10 %); the per-warp cost of the divergent v4 reset path drops accordingly
        fixed_actions.uniform_(-1.0, -0.93)
        ms_c_gentle, _ = timed(env_only_c_loop, dev)
        done_rate_gentle = float(done_buf.float().mean())
        fixed_actions.uniform_(-1.0, -0.6)
        ms_pol, _ = timed(env_and_policy, dev)
        lines.append({"config": "cfg4", "env": "fishing-v4", "n_envs_per_gpu": n, "n_steps": n_steps, "n_gpus": world,
                      "ms_env_only": ms_env, "env_steps_per_s_env_only": world * n * n_steps / (ms_env * 1e-3),
                      "ms_env_only_c_loop": ms_c, "env_steps_per_s_env_only_c_loop": world * n * n_steps / (ms_c * 1e-3),
                      "ms_env_only_c_loop_gentle_actions": ms_c_gentle, "done_rate_gentle": done_rate_gentle,
                      "env_steps_per_s_env_only_c_loop_gentle_actions": world * n * n_steps / (ms_c_gentle * 1e-3),
                      "ms_env_plus_policy": ms_pol, "env_steps_per_s_env_plus_policy": world * n * n_steps / (ms_pol * 1e-3),
                      "bytes_per_env_step": 57, "done_rate_last": float(done_buf[-1].float().mean()),
                      "mean_reward_last": float(rew_buf[-1].mean())})
        venv.close()
        del obs_buf, rew_buf, done_buf, act_buf, fixed_actions
        torch.cuda.empty_cache()

    if "cfg5" in only:
        # escapement-policy grid sweep: 1024 levels x 16384 noise replicates per box, common random numbers, one collective
        levels = 1024
        reps = max(2 * world, int(16384 * share * args.scale) // (2 * world) * (2 * world))
        ms, res = timed(lambda: gf.escapement_sweep("fishing-v1", levels=levels, replicates=reps, T=1000, device=dev, seed=0,
                                                    r=0.3, K=1, sigma=0.1, init_state=0.75, Tmax=1000), dev, reps=2)
        best = int(torch.argmax(res["mean"]))
        lines.append({"config": "cfg5", "env": "fishing-v1", "levels": levels, "replicates_total": reps, "T": 1000, "n_gpus": world,
                      "ms": ms, "env_steps_per_s": levels * reps * 1000 / (ms * 1e-3), "best_S": float(res["S"][best]),
                      "best_mean_return": float(res["mean"][best]), "std_at_best": float(res["std"][best]),
                      "checksum_mean": float(res["mean"].sum())})

    if "f1" in only:
        # next-row f1: the log-normal growth family, K1 lock-step (step_many, PDL) and fused K2 with in-kernel escapement
        from gym_fishing_b200.vec_env import growth_kwargs, normalise_kwargs

        n, T = 1 << 20, 200
        acts = torch.rand(T, n, dtype=torch.float64, device=dev) * 0.3 - 1.0
        obs = torch.empty(T, n, dtype=torch.float64, device=dev)
        rew = torch.empty_like(obs)
        done = torch.empty(T, n, dtype=torch.uint8, device=dev)
        for env_id in ("fishing-v5", "fishing-v6", "fishing-v7", "fishing-v8", "fishing-v9", "fishing-v10", "fishing-v11"):
            kind = nat.KIND_OF_ID[env_id]
            kw = dict(sigma=0.05, Tmax=100) if kind != nat.V11 else dict(Tmax=100)
            cfg, gk = normalise_kwargs(kind, kw), growth_kwargs(kind, kw)
            if kind == nat.V11:
                gk["models"] = [(m, dict(p, sigma=0.05)) for m, p in gk["models"]]
            b = EnvBatch(kind, n, device=dev, env_id0=rank * n, auto_reset=True, r=cfg["r"], K=cfg["K"], sigma=cfg["sigma"],
                         C=cfg["C"], x0=cfg["x0"], Tmax=cfg["Tmax"], **gk)
            b.reset()
            ms1, _ = timed(lambda: b.step_many(acts, obs, rew, done), dev)
            ms2, _ = timed(lambda: b.rollout(T, policy="escapement", S=0.5 * float(cfg["K"])), dev)
            lines.append({"config": "f1", "env": env_id, "n_envs_per_gpu": n, "T": T, "n_gpus": world,
                          "k1_env_steps_per_s": world * n * T / (ms1 * 1e-3), "k1_us_per_launch": ms1 * 1e3 / T,
                          "k2_escapement_env_steps_per_s": world * n * T / (ms2 * 1e-3)})

    if rank == 0:
        for ln in lines:
            print(json.dumps(ln), flush=True)
    if world > 1:
        torch.distributed.barrier()
        torch.distributed.destroy_process_group()


if __name__ == "__main__":
    main()
