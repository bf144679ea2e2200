"""Host<->device copy bandwidth of the box (pinned memory), and the e2e host front end at several chunk counts."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch

def bw(nbytes, direction, reps=20):
    h = torch.empty(nbytes, dtype=torch.uint8).pin_memory()
    d = torch.empty(nbytes, dtype=torch.uint8, device="cuda")
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(reps):
        if direction == "h2d": d.copy_(h, non_blocking=True)
        else: h.copy_(d, non_blocking=True)
    torch.cuda.synchronize()
    return nbytes * reps / (time.perf_counter() - t0) / 1e9

def bidir(nbytes, reps=20):
    h1 = torch.empty(nbytes, dtype=torch.uint8).pin_memory(); d1 = torch.empty(nbytes, dtype=torch.uint8, device="cuda")
    h2 = torch.empty(nbytes, dtype=torch.uint8).pin_memory(); d2 = torch.empty(nbytes, dtype=torch.uint8, device="cuda")
    s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(reps):
        with torch.cuda.stream(s1): d1.copy_(h1, non_blocking=True)
        with torch.cuda.stream(s2): h2.copy_(d2, non_blocking=True)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    return nbytes * reps / dt / 1e9

for mb in (1, 4, 16, 64):
    n = mb << 20
    print(f"{mb:3d} MiB  h2d {bw(n,'h2d'):6.1f} GB/s   d2h {bw(n,'d2h'):6.1f} GB/s   bidir(each) {bidir(n):6.1f} GB/s")

import gym_fishing_b200 as gf
from gym_fishing_b200.vec_env import _PinnedArray
n = 1 << 20
for chunks in (1, 2, 3, 4, 6, 8, 12):
    os.environ["FISHING_VEC_CHUNKS"] = str(chunks)
    v = gf.FishingVecEnv("fishing-v1", num_envs=n, io="numpy", sigma=0.01, Tmax=100)
    v.reset()
    pin = _PinnedArray(n, np.float64); pin.array[:] = np.random.default_rng(0).uniform(-1, 1, n)
    for _ in range(5): v.step(pin.array)
    t0 = time.perf_counter()
    for _ in range(100): v.step(pin.array)
    dt = (time.perf_counter() - t0) / 100
    print(f"chunks {chunks:2d}: {dt*1e6:7.1f} us/step  {n/dt:.3e} env-steps/s")
    v.close(); pin.free()
