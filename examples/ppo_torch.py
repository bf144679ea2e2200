#!/usr/bin/env python
"""A small PPO loop in plain torch on the batched FishingVecEnv (BASELINE.json config 4: the environment writes
observations / rewards / dones straight into slices of a preallocated rollout buffer; nothing leaves the GPU).

    python examples/ppo_torch.py [--env fishing-v1] [--envs 16384] [--iters 40]

The reference trains stable-baselines3 PPO on ONE Python environment (examples/PPO.py, ~12 minutes for 2e5 steps);
here one iteration collects envs x 64 steps.  Prints the mean episode return next to the MSY and constant-escapement
baselines computed by the fused rollout kernel.
"""
import argparse
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

import torch
from torch import nn

import gym_fishing_b200 as gf


def main(argv=None):
    ap = argparse.ArgumentParser()
    ap.add_argument("--env", default="fishing-v1")
    ap.add_argument("--envs", type=int, default=16384)
    ap.add_argument("--steps", type=int, default=64)
    ap.add_argument("--iters", type=int, default=40)
    ap.add_argument("--sigma", type=float, default=0.05)
    ap.add_argument("--seed", type=int, default=0)
    args = ap.parse_args(argv)
    torch.manual_seed(args.seed)
    dev = torch.device("cuda")
    n, T = args.envs, args.steps
    kw = dict(sigma=args.sigma, Tmax=100)
    venv = gf.FishingVecEnv(args.env, num_envs=n, seed=args.seed, **kw)

    actor = nn.Sequential(nn.Linear(1, 64), nn.Tanh(), nn.Linear(64, 64), nn.Tanh(), nn.Linear(64, 1)).to(dev)
    critic = nn.Sequential(nn.Linear(1, 64), nn.Tanh(), nn.Linear(64, 64), nn.Tanh(), nn.Linear(64, 1)).to(dev)
    with torch.no_grad():                      # start from a cautious policy: quota ~ 0.1 K, small exploration noise
        actor[-1].weight.mul_(0.01)
        actor[-1].bias.fill_(-0.9)
    log_std = nn.Parameter(torch.full((1,), -2.0, device=dev))
    opt = torch.optim.Adam(list(actor.parameters()) + list(critic.parameters()) + [log_std], lr=3e-4)
    gamma, lam, clip, epochs, minibatches = 0.99, 0.95, 0.2, 4, 8

    f64 = dict(dtype=torch.float64, device=dev)
    obs_buf = torch.empty(T + 1, n, **f64)
    rew_buf = torch.empty(T, n, **f64)
    done_buf = torch.empty(T, n, dtype=torch.uint8, device=dev)
    act_buf = torch.empty(T, n, **f64)
    logp_buf = torch.empty(T, n, device=dev)
    val_buf = torch.empty(T + 1, n, device=dev)

    # classical baselines on the same model, evaluated with the fused kernel (one launch each)
    S, q = gf.bmsy(venv.kind, r=venv.cfg["r"], K=venv.cfg["K"], C=venv.cfg["C"])
    for name, pol in (("escapement", dict(policy="escapement", S=S)), ("msy", dict(policy="msy", msy_q=q))):
        st = gf.rollout_stats(args.env, 1 << 16, 100, **pol, **kw)
        print(f"baseline {name:10s}: mean episode return {float(st['mean']):.3f}")

    obs_buf[0].copy_(venv.reset().view(-1))
    ep_returns = []
    running = torch.zeros(n, **f64)            # return of the episode each env is currently in (kept across rollouts)
    t0 = time.time()
    for it in range(args.iters):
        with torch.no_grad():
            for t in range(T):
                o32 = obs_buf[t].float().unsqueeze(1)
                mu = actor(o32).squeeze(1)
                a = mu + log_std.exp() * torch.randn_like(mu)
                logp_buf[t] = -0.5 * ((a - mu) / log_std.exp()) ** 2 - log_std - 0.9189385
                val_buf[t] = critic(o32).squeeze(1)
                act_buf[t].copy_(a)
                # the environment writes this step's results straight into the rollout buffer slices
                venv.step_into(act_buf[t], obs_buf[t + 1], rew_buf[t], done_buf[t])
            val_buf[T] = critic(obs_buf[T].float().unsqueeze(1)).squeeze(1)
            # generalised advantage estimation
            adv = torch.zeros(T, n, device=dev)
            last = torch.zeros(n, device=dev)
            for t in reversed(range(T)):
                nonterm = 1.0 - done_buf[t].float()
                delta = rew_buf[t].float() + gamma * val_buf[t + 1] * nonterm - val_buf[t]
                last = delta + gamma * lam * nonterm * last
                adv[t] = last
            ret = adv + val_buf[:T]
            # episode statistics straight from the buffers: return of the episodes that ended in this rollout
            ended = done_buf.bool()
            tot, cnt = 0.0, 0
            for t in range(T):
                running += rew_buf[t]
                tot += running[ended[t]].sum().item()
                cnt += int(ended[t].sum().item())
                running[ended[t]] = 0.0
        if cnt:
            ep_returns.append(tot / cnt)
        b_obs = obs_buf[:T].reshape(-1, 1).float()
        b_act, b_logp = act_buf.reshape(-1).float(), logp_buf.reshape(-1)
        b_adv, b_ret = adv.reshape(-1), ret.reshape(-1)
        b_adv = (b_adv - b_adv.mean()) / (b_adv.std() + 1e-8)
        N = b_obs.shape[0]
        for _ in range(epochs):
            perm = torch.randperm(N, device=dev)
            for mb in perm.chunk(minibatches):
                mu = actor(b_obs[mb]).squeeze(1)
                logp = -0.5 * ((b_act[mb] - mu) / log_std.exp()) ** 2 - log_std - 0.9189385
                ratio = (logp - b_logp[mb]).exp()
                pg = -torch.min(ratio * b_adv[mb], ratio.clamp(1 - clip, 1 + clip) * b_adv[mb]).mean()
                vf = 0.5 * (critic(b_obs[mb]).squeeze(1) - b_ret[mb]).pow(2).mean()
                loss = pg + 0.5 * vf
                opt.zero_grad(set_to_none=True)
                loss.backward()
                opt.step()
        obs_buf[0].copy_(obs_buf[T])
        if it % 5 == 0 or it == args.iters - 1:
            print(f"iter {it:3d}  env-steps {(it + 1) * n * T:>12,d}  mean return of finished episodes "
                  f"{ep_returns[-1] if ep_returns else float('nan'):.3f}  ({time.time() - t0:.1f}s)", flush=True)
    return ep_returns


if __name__ == "__main__":
    main()
