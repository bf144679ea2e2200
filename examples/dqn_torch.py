#!/usr/bin/env python
"""A small DQN in plain torch on the batched, Discrete-action fishing-v0 (the counterpart of the reference's
tests/test-DQN.py / stable-baselines3 DQN on one Python environment).

    python examples/dqn_torch.py [--envs 4096] [--iters 600]

Every iteration steps all environments once with an epsilon-greedy policy (int64 actions straight into the CUDA
engine), appends the transitions to a replay buffer that lives on the GPU, and takes one gradient step on a batch
sampled from it.  Prints the mean return of the episodes that finished, next to the constant-escapement baseline
computed by the fused rollout kernel.
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
    ap.add_argument("--envs", type=int, default=4096)
    ap.add_argument("--iters", type=int, default=600)
    ap.add_argument("--n-actions", dest="n_actions", type=int, default=25)
    ap.add_argument("--sigma", type=float, default=0.05)
    ap.add_argument("--seed", type=int, default=0)
    args = ap.parse_args(argv)
    torch.manual_seed(args.seed)
    dev = torch.device("cuda")
    n, A = args.envs, args.n_actions
    kw = dict(sigma=args.sigma, Tmax=100, n_actions=A)
    venv = gf.FishingVecEnv("fishing-v0", num_envs=n, seed=args.seed, **kw)
    assert venv.action_space.n == A

    S, _ = gf.bmsy(venv.kind, r=venv.cfg["r"], K=venv.cfg["K"], C=venv.cfg["C"])
    st = gf.rollout_stats("fishing-v0", 1 << 15, 100, policy="escapement", S=S, **kw)
    print(f"baseline escapement (discrete quotas): mean episode return {float(st['mean']):.3f}")

    q = nn.Sequential(nn.Linear(1, 64), nn.ReLU(), nn.Linear(64, 64), nn.ReLU(), nn.Linear(64, A)).to(dev)
    q_target = nn.Sequential(nn.Linear(1, 64), nn.ReLU(), nn.Linear(64, 64), nn.ReLU(), nn.Linear(64, A)).to(dev)
    q_target.load_state_dict(q.state_dict())
    opt = torch.optim.Adam(q.parameters(), lr=1e-3)
    gamma, cap, batch = 0.99, 64 * n, 4 * n
    buf_o = torch.empty(cap, device=dev)
    buf_o2 = torch.empty(cap, device=dev)
    buf_a = torch.empty(cap, dtype=torch.int64, device=dev)
    buf_r = torch.empty(cap, device=dev)
    buf_d = torch.empty(cap, device=dev)
    filled, head = 0, 0

    obs = venv.reset().view(-1)
    running = torch.zeros(n, dtype=torch.float64, device=dev)
    ep_returns = []
    t0 = time.time()
    for it in range(args.iters):
        eps = max(0.05, 1.0 - it / (0.5 * args.iters))
        with torch.no_grad():
            greedy = q(obs.float().unsqueeze(1)).argmax(1)
            explore = torch.rand(n, device=dev) < eps
            a = torch.where(explore, torch.randint(0, A, (n,), device=dev), greedy)
        obs2, rew, done, infos = venv.step(a)
        obs2 = obs2.view(-1)
        # bootstrap from the terminal observation, not from the first observation of the next episode
        nxt = torch.where(done, venv.terminal_obs, obs2)
        sl = slice(head, head + n)
        buf_o[sl], buf_o2[sl], buf_a[sl], buf_r[sl], buf_d[sl] = obs.float(), nxt.float(), a, rew.float(), done.float()
        head = (head + n) % cap
        filled = min(cap, filled + n)
        running += rew
        if bool(done.any()):
            ep_returns.append(float(running[done].mean()))
            running[done] = 0.0
        obs = obs2
        idx = torch.randint(0, filled, (batch,), device=dev)
        with torch.no_grad():
            target = buf_r[idx] + gamma * (1.0 - buf_d[idx]) * q_target(buf_o2[idx].unsqueeze(1)).max(1).values
        pred = q(buf_o[idx].unsqueeze(1)).gather(1, buf_a[idx].unsqueeze(1)).squeeze(1)
        loss = nn.functional.smooth_l1_loss(pred, target)
        opt.zero_grad(set_to_none=True)
        loss.backward()
        opt.step()
        if it % 20 == 0:
            q_target.load_state_dict(q.state_dict())
        if it % 50 == 0 or it == args.iters - 1:
            recent = ep_returns[-20:]
            print(f"iter {it:4d}  env-steps {(it + 1) * n:>10,d}  eps {eps:.2f}  mean return of finished episodes "
                  f"{sum(recent) / len(recent) if recent else float('nan'):.3f}  ({time.time() - t0:.1f}s)", flush=True)
    venv.close()
    return ep_returns


if __name__ == "__main__":
    main()
