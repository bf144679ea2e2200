import sys, torch
sys.path.insert(0, "/root/repo")
from gym_fishing_b200 import _native as nat
from gym_fishing_b200.engine import EnvBatch
dev = torch.device("cuda", 0)
n, T = 1 << 20, 50
for kind, kw in ((nat.V5, dict(r=0.3, K=1.0, C=0.5)), (nat.V8, dict(r=1.0, K=1.0, M=1.0, theta=3.0, x0=1.5))):
    for lo, hi in ((-1.0, 1.0), (-1.0, -0.8)):
        kw2 = dict(kw); x0 = kw2.pop("x0", 0.75)
        b = EnvBatch(kind, n, device=dev, auto_reset=True, sigma=0.05, Tmax=100, x0=x0, seed=0, **kw2)
        b.reset()
        acts = torch.rand(T, n, dtype=torch.float64, device=dev) * (hi - lo) + lo
        obs = torch.empty(T, n, dtype=torch.float64, device=dev); rew = torch.empty_like(obs)
        done = torch.empty(T, n, dtype=torch.uint8, device=dev)
        b.step_many(acts, obs, rew, done); torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); b.step_many(acts, obs, rew, done); e1.record(); torch.cuda.synchronize()
        print(f"kind {kind} actions [{lo},{hi}]: {e0.elapsed_time(e1)*1e3/T:.2f} us/launch, frac zero stock {(obs[-1] <= -1).float().mean().item():.3f}")
