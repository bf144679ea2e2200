#!/usr/bin/env python
"""e2e rate of FishingVecEnv(io="numpy").step_many (float64 wire, 2^20 envs, pinned buffers) against the number of chunks the
host pipeline cuts the batch into (FISHING_VEC_MANY_CHUNKS is read once per process, hence one subprocess per setting)."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
WORKER = r'''
import sys, time
sys.path.insert(0, %r)
import numpy as np, torch
import gym_fishing_b200 as gf
from gym_fishing_b200.vec_env import _PinnedArray
n, T = 1 << 20, int(sys.argv[1])
v = gf.FishingVecEnv("fishing-v1", num_envs=n, io="numpy", sigma=0.01, Tmax=100)
v.reset()
pa = _PinnedArray(T * n, np.float64); a = pa.array.reshape(T, n); a[:] = np.random.default_rng(0).uniform(-1, 1, (T, n))
outs = tuple(_PinnedArray(T * n, dt) for dt in (np.float64, np.float64, np.uint8))
o = tuple(x.array.reshape(T, n) for x in outs)
v.step_many(a, out=o)
t0 = time.perf_counter(); reps = max(2, 256 // T)
for _ in range(reps): v.step_many(a, out=o)
print(n * T * reps / (time.perf_counter() - t0))
''' % ROOT

for T in (8, 32):
    for chunks in (3, 6, 12, 24, 48):
        env = dict(os.environ, FISHING_VEC_MANY_CHUNKS=str(chunks))
        r = subprocess.run([sys.executable, "-c", WORKER, str(T)], capture_output=True, text=True, env=env)
        print(json.dumps({"T_per_call": T, "chunks": chunks, "env_steps_per_s": float(r.stdout.strip().splitlines()[-1]) if r.returncode == 0 else r.stderr[-300:]}), flush=True)

# the per-step entry point (fishing_vec_step, FISHING_VEC_CHUNKS): pinned actions in, library-owned pinned arrays out
STEP_WORKER = r'''
import sys, time
sys.path.insert(0, %r)
import numpy as np
import gym_fishing_b200 as gf
from gym_fishing_b200.vec_env import _PinnedArray
n = int(sys.argv[1])
v = gf.FishingVecEnv("fishing-v1", num_envs=n, io="numpy", sigma=0.01, Tmax=100)
v.reset()
pa = _PinnedArray(n, np.float64); pa.array[:] = np.random.default_rng(0).uniform(-1, 1, n)
for _ in range(5): v.step(pa.array)
t0 = time.perf_counter(); reps = 100
for _ in range(reps): v.step(pa.array)
print(n * reps / (time.perf_counter() - t0))
''' % ROOT

for n in (1 << 18, 1 << 20):
    for chunks in (1, 2, 3, 4, 6, 8, 12):
        env = dict(os.environ, FISHING_VEC_CHUNKS=str(chunks))
        r = subprocess.run([sys.executable, "-c", STEP_WORKER, str(n)], capture_output=True, text=True, env=env)
        print(json.dumps({"per_step_api_envs": n, "chunks": chunks, "env_steps_per_s": float(r.stdout.strip().splitlines()[-1]) if r.returncode == 0 else r.stderr[-300:]}), flush=True)
