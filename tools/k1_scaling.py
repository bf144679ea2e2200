"""K1 lock-step throughput vs batch size (is the kernel DRAM-efficient once state no longer fits in L2?)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from gym_fishing_b200 import _native as nat
from gym_fishing_b200.engine import EnvBatch

dev = torch.device("cuda", 0)
def run(kind, n, T, sigma, **kw):
    acts = torch.rand(T, n, dtype=torch.float64, device=dev) * 0.4 - 1.0
    obs = torch.empty(T, n, dtype=torch.float64, device=dev); rew = torch.empty_like(obs)
    done = torch.empty(T, n, dtype=torch.uint8, device=dev)
    b = EnvBatch(kind, n, device=dev, auto_reset=True, sigma=sigma, Tmax=100, **kw)
    b.reset()
    b.step_many(acts, obs, rew, done); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(3): b.step_many(acts, obs, rew, done)
    e1.record(); torch.cuda.synchronize()
    us = e0.elapsed_time(e1) * 1e3 / (3 * T)
    return us, float(done.float().mean())

for kind, name, bpe in ((nat.V1, "v1", 41), (nat.V4, "v4", 57)):
    for sigma in (0.0, 0.01):
        for lg in (20, 21, 22, 23):
            n = 1 << lg
            T = max(16, (1 << 26) // n)
            us, dr = run(kind, n, T, sigma)
            print(f"{name} sigma={sigma} n=2^{lg}: {us:7.2f} us/launch  {n/us*1e-3:7.2f} Genv-steps/s  {bpe*n/us*1e-3:7.1f} GB/s alg  done-rate {dr:.3f}", flush=True)
