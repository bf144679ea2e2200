import sys, time, torch
sys.path.insert(0, "/root/repo")
import gym_fishing_b200 as gf
for n in (1024, 16384, 65536, 262144, 1048576):
    for io in ("torch", "numpy"):
        v = gf.FishingVecEnv("fishing-v1", num_envs=n, device="cuda", io=io, seed=0, sigma=0.01)
        obs = v.reset()
        if io == "torch":
            a = torch.rand(n, 1, dtype=torch.float64, device="cuda") * 2 - 1
        else:
            import numpy as np
            a = np.random.uniform(-1, 1, size=(n, 1))
        for _ in range(200): v.step(a)
        torch.cuda.synchronize(); t0 = time.perf_counter()
        N = 3000
        for _ in range(N): o, r, d, info = v.step(a)
        torch.cuda.synchronize(); dt = time.perf_counter() - t0
        print(f"n={n} io={io}: {dt/N*1e6:.1f} us per step() call, {n*N/dt:.3e} env-steps/s")
        v.close()
env = gf.make("fishing-v1", sigma=0.01)
env.reset()
import numpy as np
t0 = time.perf_counter()
for _ in range(2000): env.step(np.array([-0.9]))
print(f"scalar env.step: {(time.perf_counter()-t0)/2000*1e6:.1f} us per call")
