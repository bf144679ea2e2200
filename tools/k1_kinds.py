import sys, torch
sys.path.insert(0, "/root/repo")
from gym_fishing_b200 import _native as nat
from gym_fishing_b200.engine import EnvBatch
dev = torch.device("cuda", 0)
n, T = 1 << 20, 200
def run(kind, acts, **kw):
    obs = torch.empty(T, n, dtype=torch.float64, device=dev); rew = torch.empty_like(obs)
    done = torch.empty(T, n, dtype=torch.uint8, device=dev)
    b = EnvBatch(kind, n, device=dev, auto_reset=True, Tmax=100, **kw); b.reset()
    for _ in range(3): b.step_many(acts, obs, rew, done)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(3): b.step_many(acts, obs, rew, done)
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) * 1e3 / (3 * T)
f64 = torch.rand(T, n, dtype=torch.float64, device=dev) * 0.4 - 1.0
i64 = torch.randint(0, 30, (T, n), dtype=torch.int64, device=dev)
i32 = i64.to(torch.int32); f32 = f64.to(torch.float32)
for name, kind, acts, kw in (("v1 f64", nat.V1, f64, dict(sigma=0.01)), ("v1 f32 actions", nat.V1, f32, dict(sigma=0.01)),
                             ("v0 i64", nat.V0, i64, dict(sigma=0.01)), ("v0 i32", nat.V0, i32, dict(sigma=0.01)),
                             ("v2 f64", nat.V2, f64, dict(sigma=0.01)), ("v1 K=2", nat.V1, f64, dict(sigma=0.01, K=2.0, x0=1.5)),
                             ("v2 K=2", nat.V2, f64, dict(sigma=0.01, K=2.0, x0=1.5, C=1.0))):
    us = run(kind, acts, **kw)
    print(f"{name:16s} {us:6.2f} us/launch  {n/us*1e-3:6.1f} Genv-steps/s", flush=True)
