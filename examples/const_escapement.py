#!/usr/bin/env python
"""Constant-escapement rule on fishing-v1, three ways (the reference's examples/const_escapement.py is the first):

1. the single-environment gym loop, step by step (every step is one launch of the CUDA step kernel);
2. the same rule for 2^16 stochastic replicates with the batched FishingVecEnv (policy evaluated in torch);
3. the same again as ONE fused kernel launch with the policy applied in-kernel (EnvBatch.rollout / rollout_stats).
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

import pandas as pd
import torch

import gym_fishing_b200 as gf

r, K = 0.3, 1
env = gf.make("fishing-v1", r=r, K=K, sigma=0.0)
rows = []
env.reset()
obs = env.state
for t in range(env.Tmax):
    x = env.get_fish_population(obs)
    Q = max(x - K / 2, 0)                  # the escapement rule in model space
    action = env.get_action(Q)             # ... mapped to the [-1, 1] action space
    quota = env.get_quota(action)
    obs, reward, done, info = env.step(action)
    rows.append([t, x, quota, reward, 0])
df = pd.DataFrame(rows, columns=["time", "state", "action", "reward", "rep"])
print("1) gym loop:      total reward", df.reward.sum())

n = 1 << 16
venv = gf.FishingVecEnv("fishing-v1", num_envs=n, r=r, K=K, sigma=0.05, auto_reset=False)
obs = venv.reset()
total = torch.zeros(n, dtype=torch.float64, device=obs.device)
for t in range(100):
    x = (obs[:, 0] + 1) * K
    action = torch.clamp(x - K / 2, min=0.0) / K - 1
    obs, reward, done, _ = venv.step(action)
    total += reward
print(f"2) FishingVecEnv: mean return over {n} replicates (sigma=0.05): {total.mean().item():.4f} +- {total.std().item():.4f}")

stats = gf.rollout_stats("fishing-v1", n, 100, policy="escapement", S=K / 2, r=r, K=K, sigma=0.05, Tmax=100)
print(f"3) fused rollout: mean return over {n} replicates (sigma=0.05): {float(stats['mean']):.4f} +- {float(stats['std']):.4f}"
      f"   extinct: {int(stats['n_done'])}")
