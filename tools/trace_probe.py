import sys, torch
sys.path.insert(0, "/root/repo")
from gym_fishing_b200 import _native as nat
from gym_fishing_b200.engine import EnvBatch
dev = torch.device("cuda", 0)
n, T = 1 << 20, 400
acts = torch.rand(T, n, dtype=torch.float64, device=dev) * 2 - 1
for name, kw in (("fast+trace", dict(trace=True)), ("generic+trace", dict(trace=True, trace_action=True)), ("fast no trace", dict())):
    b = EnvBatch(nat.V1, n, device=dev, auto_reset=True, r=0.3, K=1.0, sigma=0.01, x0=0.75, Tmax=100, seed=0)
    b.reset()
    times = []
    for rep in range(5):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        e0.record(); out = b.rollout(T, policy="actions", actions=acts, **kw); e1.record(); torch.cuda.synchronize()
        times.append(round(e0.elapsed_time(e1) / T * 1e3, 2))
        del out
    print(name, "us per step:", times, flush=True)
