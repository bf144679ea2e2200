#!/usr/bin/env python
"""bench.py -- env-steps/s of the gym_fishing hot path on B200 (BASELINE.json's metric, configs[1]).

Workload ("cfg2"): fishing-v1 (continuous quota), 2^20 environments per GPU in lock-step, r=0.3 K=1 sigma=0.01
x0=0.75 Tmax=100, auto-reset on, float64 actions U(-1,1) from a pre-generated [T, n] tensor, Gaussian process noise
generated in-kernel (Philox), observations / rewards / dones written into a [T, n] rollout buffer.  One bench "step"
is ONE pass of T=1000 lock-steps over the batch = T * n env-steps, i.e. T launches of the single-step kernel K1
(issued from C, replayed as a CUDA graph).  Inputs (8.4 GB of actions) and outputs (18 GB) are far larger than the
126 MB L2, so consecutive bench steps cannot be served from cache.

    python bench.py [--gpus N] [--steps K] [--warmup W]            # this engine
    python bench.py --impl reference [...]                         # the unmodified reference's step() loop, all host cores

Multi-GPU: one process per GPU under torchrun; environments are independent ("replicas only" for the VecEnv
workload), every rank steps its own 2^20-env shard of the global id range, the time is the max
over ranks -> weak scaling.

Prints ONE JSON line (rank 0).  Keys beyond the contract: roofline (dominant kernel = K1, HBM-bound), cpu_baseline,
e2e (host numpy buffers through the C-ABI host front end, copies inside the timed region), fused (the register-resident
K2 variants of the same workload, reported for context), clocks, gpu_launches.
"""
from __future__ import annotations

import argparse
import ctypes
import json
import multiprocessing as mp
import os
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

CFG = dict(r=0.3, K=1.0, sigma=0.01, x0=0.75, Tmax=100)
METRIC = "env-steps/sec"
UNIT = "env-steps/s"
# algorithmic bytes per env-step of K1 in this mode: state in 8 + action in 8 + obs out 8 +
# reward out 8 + done out 1 + years int32 read+write 8
K1_BYTES = 41



_JSON_OUT = None


def emit(line):
    """Write the result line to the real stdout (see main(): file descriptor 1 itself is pointed at stderr)."""
    out = _JSON_OUT or sys.stdout
    out.write(json.dumps(line) + "\n")
    out.flush()


def measured_peak():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:  # noqa: BLE001
        return 6650.0, "fallback (B200_PROFILING.md)"


# ------------------------------------------------------------------------------------------------ clocks sampler
class ClockSampler:
    """Samples SM clock and throttle reasons through NVML while the timed region runs."""

    def __init__(self, index):
        self.samples, self.reasons, self.max_mhz = [], set(), None
        self._stop = threading.Event()
        self._thread = None
        try:
            import pynvml

            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception:  # noqa: BLE001
            self.nv = None

    def _run(self):
        nv = self.nv
        names = {
            "hw_slowdown": getattr(nv, "nvmlClocksEventReasonHwSlowdown", getattr(nv, "nvmlClocksThrottleReasonHwSlowdown", 0x8)),
            "hw_thermal_slowdown": getattr(nv, "nvmlClocksEventReasonHwThermalSlowdown", getattr(nv, "nvmlClocksThrottleReasonHwThermalSlowdown", 0x40)),
            "sw_thermal_slowdown": getattr(nv, "nvmlClocksEventReasonSwThermalSlowdown", getattr(nv, "nvmlClocksThrottleReasonSwThermalSlowdown", 0x20)),
            "sw_power_cap": getattr(nv, "nvmlClocksEventReasonSwPowerCap", getattr(nv, "nvmlClocksThrottleReasonSwPowerCap", 0x4)),
        }
        get = getattr(nv, "nvmlDeviceGetCurrentClocksEventReasons", None) or nv.nvmlDeviceGetCurrentClocksThrottleReasons
        while not self._stop.is_set():
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                mask = get(self.h)
                for k, bit in names.items():
                    if mask & bit:
                        self.reasons.add(k)
            except Exception:  # noqa: BLE001
                pass
            time.sleep(0.01)

    def __enter__(self):
        if self.nv is not None:
            self._thread = threading.Thread(target=self._run, daemon=True)
            self._thread.start()
        return self

    def __exit__(self, *a):
        self._stop.set()
        if self._thread is not None:
            self._thread.join()

    def summary(self):
        s = sorted(self.samples)
        return {"sm_mhz": (s[len(s) // 2] if s else None), "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons),
                "samples": len(s)}


# ------------------------------------------------------------------------------------------------ CPU baselines
def _port_worker(args):
    """`steps` env.step calls of the reference-shaped scalar port (oracle/fishing_oracle.py:ScalarPortEnv)."""
    steps, seed = args
    import numpy as np

    from oracle.fishing_oracle import ScalarPortEnv

    np.random.seed(seed)
    env = ScalarPortEnv("v1", r=CFG["r"], K=1, sigma=CFG["sigma"], init_state=CFG["x0"], Tmax=CFG["Tmax"])
    acts = np.random.uniform(-1, 1, size=(1024, 1))
    env.reset()
    t0 = time.perf_counter()
    for i in range(steps):
        _, _, done, _ = env.step(acts[i & 1023])
        if done:
            env.reset()
    return time.perf_counter() - t0


def cpu_port_throughput(steps_per_proc, procs):
    return cpu_loop_throughput(_port_worker, steps_per_proc, procs)


def cpu_vectorised_throughput(n=1 << 20, T=8):
    """Best-effort CPU line: the numpy-vectorised oracle on the same batch shape (single thread)."""
    import numpy as np

    from oracle import fishing_oracle as fo

    rng = np.random.default_rng(0)
    s = np.full(n, CFG["x0"] / CFG["K"] - 1)
    years = np.zeros(n, dtype=np.int64)
    a = rng.uniform(-1, 1, size=(T, n))
    z = rng.standard_normal((T, n))
    t0 = time.perf_counter()
    for t in range(T):
        s, _, done, years = fo.step_vec("v1", s, a[t], z[t], years, r=CFG["r"], K=CFG["K"], sigma=CFG["sigma"], Tmax=CFG["Tmax"])
        s = np.where(done, CFG["x0"] / CFG["K"] - 1, s)
        years = np.where(done, 0, years)
    return n * T / (time.perf_counter() - t0)


def cpu_c_port_throughput(n=1 << 20, T=16):
    """Best-a-CPU-can-do line: the C restatement (oracle/fishing_oracle.c, -O2, OpenMP over all cores, no FMA)."""
    import numpy as np

    try:
        from oracle.build_c import load

        lib = load()
    except Exception:  # noqa: BLE001
        return None
    rng = np.random.default_rng(0)
    a = rng.uniform(-1, 1, size=(T, n))
    z = rng.standard_normal((T, n))
    s = np.full(n, CFG["x0"] / CFG["K"] - 1)
    years = np.zeros(n, dtype=np.int32)
    rew, done = np.zeros(n), np.zeros(n, dtype=np.uint8)
    args = (1, n, T, CFG["r"], CFG["K"], CFG["sigma"], 0.5, 100.0, CFG["Tmax"], CFG["x0"] / CFG["K"] - 1, a.ctypes.data, z.ctypes.data,
            s.ctypes.data, years.ctypes.data, rew.ctypes.data, done.ctypes.data)
    lib.oracle_rollout(*args)
    best = 0.0
    for _ in range(3):
        t0 = time.perf_counter()
        lib.oracle_rollout(*args)
        best = max(best, n * T / (time.perf_counter() - t0))
    return best


REF_DIR = os.path.join(ROOT, "baseline", "_ref")


def reference_available():
    return os.path.isdir(os.path.join(REF_DIR, "gym_fishing"))


def _ref_worker(args):
    """`steps` env.step calls of the UNMODIFIED reference (baseline/_ref/gym_fishing, installed by baseline/install_ref.py)
    under the gym / matplotlib stand-ins of oracle/refstub (neither package is in the image): the loop a user of
    boettiger-lab/gym_fishing runs -- gym.make("fishing-v1", ...), env.step(action), env.reset() on done."""
    steps, seed = args
    for p in (os.path.join(ROOT, "oracle", "refstub"), REF_DIR):
        if p not in sys.path:
            sys.path.insert(0, p)
    import gym
    import gym_fishing  # noqa: F401  (registers the ids)
    import numpy as np

    np.random.seed(seed)
    env = gym.make("fishing-v1", r=CFG["r"], K=1, sigma=CFG["sigma"], init_state=CFG["x0"], Tmax=CFG["Tmax"])
    acts = np.random.uniform(-1, 1, size=(1024, 1))
    env.reset()
    t0 = time.perf_counter()
    for i in range(steps):
        _, _, done, _ = env.step(acts[i & 1023])
        if done:
            env.reset()
    return time.perf_counter() - t0


def _ref_simulate_worker(args):
    """cfg 1 on the reference: fishing-v0, escapement policy, env.simulate(model, reps) with Tmax = 1000 (shared_env.py:29-54);
    returns (seconds, steps)."""
    reps, seed = args
    for p in (os.path.join(ROOT, "oracle", "refstub"), REF_DIR):
        if p not in sys.path:
            sys.path.insert(0, p)
    import gym
    import gym_fishing  # noqa: F401
    import numpy as np
    from gym_fishing.models.policies import escapement

    np.random.seed(seed)
    env = gym.make("fishing-v0", sigma=0.05, Tmax=1000)
    model = escapement(env)
    env.simulate(model, reps=1)
    t0 = time.perf_counter()
    df = env.simulate(model, reps=reps)
    return time.perf_counter() - t0, len(df)


def cfg1_leg(dev):
    """BASELINE.json config 1 through the scalar drop-in API: make("fishing-v0"), escapement policy, 1000-step episodes.
    (i) the caller's own loop over model.predict / env.step (one kernel launch + one synchronisation per step),
    (ii) env.simulate(model, reps) -- all repetitions as one fused rollout with the reference's use of numpy's generator,
    (iii) the unmodified reference's env.simulate on one core, when baseline/_ref is there."""
    import numpy as np

    import gym_fishing_b200 as gf

    np.random.seed(0)
    env = gf.make("fishing-v0", sigma=0.05, Tmax=1000, device=dev)
    model = gf.escapement(env)
    obs = env.reset()
    for _ in range(50):
        obs, _, _, _ = env.step(model.predict(obs)[0])
    obs = env.reset()
    t0 = time.perf_counter()
    for _ in range(1000):
        a, _ = model.predict(obs)
        obs, r, d, _ = env.step(a)
    step_rate = 1000 / (time.perf_counter() - t0)
    env.simulate(model, reps=2)
    reps = 200
    t0 = time.perf_counter()
    df = env.simulate(model, reps=reps)
    sim_rate = len(df) / (time.perf_counter() - t0)
    out = {"workload": "cfg1: fishing-v0 (n_actions=100, sigma=0.05), constant-escapement policy, 1000-step episodes through make()",
           "step_loop": {"value": step_rate, "unit": UNIT, "what": "1000 x (model.predict, env.step): one launch + one synchronisation per step"},
           "simulate": {"value": sim_rate, "unit": UNIT, "reps": reps,
                        "what": "env.simulate(escapement(env), reps): one fused rollout for all repetitions, noise drawn from numpy's "
                                "global generator in the reference's order (gym_fishing_b200/shared_env.py:simulate_on_device)"}}
    if reference_available():
        sec, steps = _ref_simulate_worker((20, 0))
        out["reference_simulate"] = {"value": steps / sec, "unit": UNIT, "reps": 20, "cores": 1,
                                     "what": "the unmodified reference's env.simulate(escapement(env), reps) (baseline/_ref)"}
    return out


def cpu_loop_throughput(worker, steps_per_proc, procs):
    if procs == 1:
        return steps_per_proc / worker((steps_per_proc, 0))
    ctx = mp.get_context("spawn")
    with ctx.Pool(procs) as pool:
        pool.map(worker, [(1000, i) for i in range(procs)])  # warm the workers (imports)
        t0 = time.perf_counter()
        pool.map(worker, [(steps_per_proc, i) for i in range(procs)])
        wall = time.perf_counter() - t0
    return steps_per_proc * procs / wall


def run_reference(args):
    """--impl reference: the reference's own CPU implementation of the path, on every host core.  When baseline/_ref holds
    the unmodified reference package (baseline/install_ref.py; it travels with the tree) THAT is what runs
    (cpu_baseline.kind = "reference"); otherwise the oracle's reference-shaped scalar port (kind = "port": the same
    numpy-scalar operations per step, validated bit for bit against the reference in tests/test_oracle_golden.py)."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    per_proc = args.ref_steps_per_core
    have_ref = reference_available() and not args.ref_port
    worker = _ref_worker if have_ref else _port_worker
    vals = []
    for _ in range(args.warmup):
        cpu_loop_throughput(worker, max(1000, per_proc // 10), cores)
    for _ in range(args.steps):
        vals.append(cpu_loop_throughput(worker, per_proc, cores))
    v = sum(vals) / len(vals)
    what = ("the unmodified reference (baseline/_ref/gym_fishing, gym.make('fishing-v1') under the oracle/refstub gym stand-in)"
            if have_ref else "oracle.ScalarPortEnv (" + ("--ref-port" if args.ref_port else "baseline/_ref is absent on this machine") + ")")
    sample = (f"{per_proc} env.step calls per process x {cores} processes per bench step of {what}, fishing-v1 sigma=0.01, "
              "reset() on done")
    line = {
        "impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * per_proc * cores / v, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": workload_config(args, int(os.environ.get("WORLD_SIZE", "1"))),      # the same dict as this run's GPU arm prints
        "cpu_baseline": {"value": v, "unit": UNIT, "cores": cores, "kind": "reference" if have_ref else "port", "sample": sample},
        "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    emit(line)


def workload_config(args, world):
    return {"workload": "cfg2: fishing-v1, 2^20 envs/GPU lock-step, sigma=0.01, random-action tensor [T,n] f64, in-kernel Philox noise, "
                        "auto-reset, rollout-buffer outputs",
            "envs_per_gpu": args.envs, "T_lockstep_per_bench_step": args.T, "env": "fishing-v1", "sigma": CFG["sigma"],
            "r": CFG["r"], "K": CFG["K"], "x0": CFG["x0"], "Tmax": CFG["Tmax"], "parallelism": f"replicas x{world}",
            "l2": "inputs+outputs per bench step (26 GB) >> 126 MB L2; no flush needed",
            "launch": "stream" if getattr(args, "no_graph", False) else "cuda-graph replay of T K1 launches (programmatic dependent launch between them)"}


# ------------------------------------------------------------------------------------------------ GPU arm
def run_gpu(args):
    import torch

    from gym_fishing_b200 import _native as nat
    from gym_fishing_b200.distributed import barrier_max_ms, init_from_env
    from gym_fishing_b200.engine import EnvBatch
    from gym_fishing_b200.vec_env import FishingVecEnv

    rank, world, local = init_from_env()
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the engine has no CPU fallback")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    import torch.distributed as dist

    n, T = args.envs, args.T
    lib = nat.lib()
    g = torch.Generator(device=dev).manual_seed(1234 + rank)
    actions = torch.rand((T, n), dtype=torch.float64, device=dev, generator=g) * 2 - 1
    obs = torch.empty((T, n), dtype=torch.float64, device=dev)
    rew = torch.empty((T, n), dtype=torch.float64, device=dev)
    done = torch.empty((T, n), dtype=torch.uint8, device=dev)
    batch = EnvBatch(nat.V1, n, device=dev, env_id0=rank * n, seed=0, auto_reset=True, **CFG)
    batch.reset()

    # capture T launches of K1 into one CUDA graph on a side stream (the kernels go to whatever stream is current);
    # --no-graph (or a failed capture) issues the same launches directly from fishing_step_many's C loop instead
    stream = torch.cuda.Stream(device=dev)
    graph = None
    with torch.cuda.stream(stream):
        batch.step_many(actions, obs, rew, done)  # warm-up outside capture (module load etc.)
        stream.synchronize()
        l0 = lib.fishing_launch_count()
        if not args.no_graph:
            try:
                graph = torch.cuda.CUDAGraph()
                with torch.cuda.graph(graph, stream=stream):
                    batch.step_many(actions, obs, rew, done)
            except Exception as exc:  # noqa: BLE001
                print(f"bench: CUDA-graph capture failed ({exc}); falling back to direct launches", file=sys.stderr)
                graph = None
                torch.cuda.synchronize()
        if graph is None:
            l0 = lib.fishing_launch_count()
            batch.step_many(actions, obs, rew, done)
        launches_per_step = lib.fishing_launch_count() - l0
    torch.cuda.synchronize()

    def one_bench_step():
        if graph is not None:
            graph.replay()
        else:
            batch.step_many(actions, obs, rew, done)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    with torch.cuda.stream(stream):
        for _ in range(args.warmup):
            one_bench_step()
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        with ClockSampler(local) as clocks:
            e0.record(stream)
            for _ in range(args.steps):
                one_bench_step()
            e1.record(stream)
            barrier()
        ms = e0.elapsed_time(e1)
    ms = barrier_max_ms(ms, dev)
    value = world * n * T * args.steps / (ms * 1e-3)
    k1_us = ms * 1e3 / (args.steps * launches_per_step)

    # ---- the same kernel at the north-star's 2^21 envs per GPU and at 2^23 (state far beyond the L2: DRAM-resident)
    def k1_point(n_envs, T_pt):
        g2 = torch.Generator(device=dev).manual_seed(99 + rank)
        a2 = torch.rand((T_pt, n_envs), dtype=torch.float64, device=dev, generator=g2) * 2 - 1
        o2 = torch.empty((T_pt, n_envs), dtype=torch.float64, device=dev)
        r2 = torch.empty((T_pt, n_envs), dtype=torch.float64, device=dev)
        d2 = torch.empty((T_pt, n_envs), dtype=torch.uint8, device=dev)
        bp = EnvBatch(nat.V1, n_envs, device=dev, env_id0=rank * n_envs, seed=0, auto_reset=True, **CFG)
        bp.reset()
        with torch.cuda.stream(stream):
            bp.step_many(a2, o2, r2, d2)
            stream.synchronize()
            p0, p1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            p0.record(stream)
            for _ in range(3):
                bp.step_many(a2, o2, r2, d2)
            p1.record(stream)
            stream.synchronize()
        us = barrier_max_ms(p0.elapsed_time(p1), dev) * 1e3 / (3 * T_pt)
        gbs = K1_BYTES * n_envs / (us * 1e-6) / 1e9
        return {"envs_per_gpu": n_envs, "avg_launch_us": us, "achieved": gbs, "unit": "GB/s", "value": world * n_envs / (us * 1e-6),
                "launches": "issued from C (fishing_step_many), no graph"}

    k1_points = []
    if not args.no_configs:
        peak_now, _ = measured_peak()
        for n_pt, T_pt in ((1 << 21, 200), (1 << 23, 60)):
            pt = k1_point(n_pt, T_pt)
            pt["frac"] = pt["achieved"] / peak_now
            k1_points.append(pt)
        torch.cuda.empty_cache()

    # sanity: the timed work really happened (dones fire, rewards are finite and non-trivial)
    frac_done = float(done[-1].float().mean().item())
    mean_rew = float(rew[-1].mean().item())
    assert 0.0 < frac_done < 1.0 and mean_rew > 0.0 and bool(torch.isfinite(obs[-1]).all())

    # ---- fused variants of the same workload (context): K2 with the action tensor, trace on / off
    fused = {}
    with torch.cuda.stream(stream):
        for name, kw in (("actions_tensor_with_trace", dict(trace=True)), ("actions_tensor_returns_only", dict())):
            b2 = EnvBatch(nat.V1, n, device=dev, env_id0=rank * n, seed=0, auto_reset=True, **CFG)
            b2.reset()
            out = b2.rollout(T, policy="actions", actions=actions, **kw)
            stream.synchronize()
            f0, f1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            reps = 3
            f0.record(stream)
            for _ in range(reps):
                out = None      # release the previous call's trace tensors first: no allocator growth inside the timed region
                out = b2.rollout(T, policy="actions", actions=actions, **kw)
            f1.record(stream)
            stream.synchronize()
            fms = barrier_max_ms(f0.elapsed_time(f1) / reps, dev)
            fused[name] = {"value": world * n * T / (fms * 1e-3), "unit": UNIT, "ms_per_step": fms}
            del out, b2
        # FP64-pipe peak: a DFMA loop measured here, on this GPU, now
        sm_count = torch.cuda.get_device_properties(dev).multi_processor_count
        scratch = torch.zeros(256, dtype=torch.float64, device=dev)
        p_blocks, p_iters = sm_count * 8, 8192
        sp = ctypes.c_void_p(stream.cuda_stream)
        nat.check(nat.lib().fishing_probe_fp64(scratch.data_ptr(), p_blocks, 64, sp), "fishing_probe_fp64")
        stream.synchronize()
        best = float("inf")
        for _ in range(3):
            f0, f1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            f0.record(stream)
            nat.check(nat.lib().fishing_probe_fp64(scratch.data_ptr(), p_blocks, p_iters, sp), "fishing_probe_fp64")
            f1.record(stream)
            stream.synchronize()
            best = min(best, f0.elapsed_time(f1))
        peak_thread_ops = p_blocks * 256 * p_iters * 8 / (best * 1e-3)          # DFMA thread-instructions per second

        def fused_policy_leg(kind, name, mangled, S, alg_ops, auto_reset, freeze, **cfg):
            """K2 with the in-kernel escapement policy + Philox noise, and its roofline: the instruction mix of the loop is
            read from the binary that just ran (tools/sass_loop.py), the SM clock from NVML during the timed launches."""
            b3 = EnvBatch(kind, n, device=dev, env_id0=rank * n, seed=0, auto_reset=auto_reset, **cfg)
            b3.reset()
            b3.rollout(T, policy="escapement", S=S, freeze_done=freeze)
            stream.synchronize()
            f0, f1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            with ClockSampler(local) as ck:
                f0.record(stream)
                for _ in range(5):
                    b3.reset()
                    b3.rollout(T, policy="escapement", S=S, freeze_done=freeze)
                f1.record(stream)
                stream.synchronize()
            fms = barrier_max_ms(f0.elapsed_time(f1) / 5, dev)
            per_gpu = n * T / (fms * 1e-3)
            fused[name] = {"value": world * per_gpu, "unit": UNIT, "ms_per_step": fms}
            try:
                sys.path.insert(0, os.path.join(ROOT, "tools"))
                from sass_loop import loop_mix

                mix = list(loop_mix(nat.LIB_PATH, mangled).values())[0]
                steps = max(1, round(mix["pipes"].get("imad.wide", 16) / 17))      # 16 wide multiplies + 1-2 per unrolled step
                fp64 = mix["pipes"].get("fp64", 0) / steps / 2                      # per environment-step (a thread owns a pair)
                alu = mix["pipes"].get("alu", 0) / steps / 2
                mhz = ck.summary()["sm_mhz"] or 1965
                pair_steps_per_smsp = (n / 2 / 32) * T / (sm_count * 4)
                cyc = mhz * 1e6 * fms * 1e-3 / pair_steps_per_smsp
                wide = mix["pipes"].get("imad.wide", 0) / steps                    # per pair-step
                t_fp64 = 2.03 * 2 * fp64 + 3.9 * wide      # cycles per pair-step the FP64 pipe is taken (see `why`)
                t_alu = 2.04 * 2 * alu                     # cycles per pair-step on the ALU pipe, which can run beside it
                port = max(t_fp64, t_alu)
                fused[name]["roofline"] = {
                    "kernel": mangled, "bound": "FP64 pipe time: FP64 instructions + the Philox wide multiplies that do not overlap them",
                    "why": "tools/issue_probe.cu on B200 (profiles/issue_probe_r02.txt), cycles per warp instruction and sub-partition: DFMA / "
                           "DADD / DMUL 2.03, LOP3 2.04, IMAD.WIDE 5.1; 2 DFMA + LOP3 4.05 and DFMA + 2 LOP3 4.13 (the lighter pipe hides), "
                           "DFMA + LOP3 2.59; but DFMA + IMAD.WIDE 6.10 and 8 DFMA + 2 IMAD.WIDE 24.0 = 8 x 2.03 + 2 x 3.9: a wide multiply "
                           "costs 3.9 cycles ON TOP of the FP64 work.  peak = the time max(2.03 FP64 + 3.9 IMAD.WIDE, 2.04 ALU) would take",
                    "instr_per_pair_step": mix["n"] / steps, "fp64_instr_per_env_step": fp64, "alu_instr_per_env_step": alu,
                    "imad_wide_per_pair_step": wide, "fp64_pipe_cycles_per_pair_step": t_fp64, "alu_pipe_cycles_per_pair_step": t_alu,
                    "measured_cycles_per_pair_step": cyc, "sm_mhz": mhz,
                    "achieved": per_gpu, "peak": per_gpu * cyc / port, "unit": UNIT, "frac": port / cyc,
                    "mix_probe": {"cycles_per_fp64_instr": 3.41, "frac": 3.41 * 2 * fp64 / cyc,
                                  "what": "a dependency-free loop of 8 DFMA + 8 LOP3 + 2 IMAD.WIDE (the logistic kernel's ratio of FP64 : "
                                          "ALU : wide multiplies, tools/issue_probe.cu) needs 27.3 cycles = 3.41 per FP64 instruction, i.e. "
                                          "0.88 of its own pipe-time bound; frac = that loop's time for this kernel's FP64 count / measured"},
                    "fp64_pipe": {"peak": peak_thread_ops, "unit": "fp64 thread-instr/s", "peak_source": "fishing_probe_fp64 (DFMA loop, this run)",
                                  "pipe_frac": fp64 * per_gpu / peak_thread_ops, "alg_fp64_ops_per_env_step": alg_ops,
                                  "useful_frac": alg_ops * per_gpu / peak_thread_ops,
                                  "note": "the north-star's 'FP64-pipe utilisation': executed FP64 instructions over the DFMA peak; it FELL "
                                          "from 0.58 (round 1) because the Box-Muller now needs 28 instead of 58 FP64 instructions per pair of "
                                          "draws, while the throughput rose -- the pipe cannot be filled next to the ALU work of Philox and "
                                          "of the dynamics' selects, see `why`"}}
            except Exception as exc:  # noqa: BLE001  (no cuobjdump on this machine)
                fused[name]["roofline"] = {"unavailable": repr(exc)}
            del b3

        fused_policy_leg(nat.V1, "escapement_in_kernel", "rollout_fast_kernelILi0ELi0ELi1ELb1ELi1ELb0ELb1EE",
                         0.49960005283355713,      # float32 BMSY of the default logistic model (tests/golden/kat.json kat1.S)
                         29, True, False, **CFG)   # fishing-v1 step 22 + continuous escapement policy 6 + return accumulation 1
        fused_policy_leg(nat.V1, "escapement_in_kernel_plain", "rollout_fast_kernelILi0ELi0ELi1ELb1ELi0ELb0ELb0EE",
                         0.49960005283355713, 29, False, False, **CFG)   # the same without auto-reset (the reference's plain loop)
        fused_policy_leg(nat.V2, "tipping_escapement_in_kernel", "rollout_fast_kernelILi1ELi0ELi1ELb1ELi2ELb0ELb0EE",
                         float.fromhex("0x1.9374bcp-1"), 31, False, True, r=0.3, K=1.0, C=0.5, sigma=0.05, x0=0.75, Tmax=1000)
    del obs, rew, done
    torch.cuda.empty_cache()

    # ---- BASELINE.json configs 3 and 5 at FULL size, sharded over the ranks of this run (strong scaling: the global problem is
    # the same for every N, so the statistics -- and the checksums below -- must be identical for N = 1, 2, 4, 8).  Each timed
    # call contains everything: reset (K3), the fused rollout (K2), the per-GPU reduction (K4) and the ONE collective.
    configs = {}
    if not args.no_configs:
        from gym_fishing_b200.distributed import allreduce_stats
        from gym_fishing_b200.rollouts import escapement_sweep, rollout_stats

        def timed_call(fn, reps=3):
            out = fn()                      # warm-up: allocator blocks, NCCL communicator
            barrier()
            c0, c1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            c0.record()
            for _ in range(reps):
                out = fn()
            c1.record()
            barrier()
            return barrier_max_ms(c0.elapsed_time(c1) / reps, dev), out

        S2 = float.fromhex("0x1.9374bcp-1")    # float32 BMSY of fishing-v2 (tests/golden/kat.json kat3b.S)
        n3, T3 = args.cfg3_envs, 1000
        ms3, r3 = timed_call(lambda: rollout_stats("fishing-v2", n3, T3, policy="escapement", S=S2, device=dev, seed=0,
                                                   r=0.3, K=1.0, C=0.5, sigma=0.05, init_state=0.75, Tmax=1000))
        configs["cfg3"] = {
            "workload": f"fishing-v2 tipping point, {n3} envs total over {world} GPU(s), fused {T3}-step constant-escapement rollout, "
                        "sigma=0.05, freeze on done; reset + K2 + K4 + all-gather of the [256,3] statistics inside the timed call",
            "value": n3 * T3 / (ms3 * 1e-3), "unit": UNIT, "ms": ms3, "scaling": "strong",
            "checksum": {"sum": float(r3["sum"]).hex(), "sumsq": float(r3["sumsq"]).hex(), "n_done": float(r3["n_done"])},
            "mean_return": float(r3["mean"]), "std_return": float(r3["std"])}
        L5, R5, T5 = 1024, args.cfg5_replicates, 1000
        ms5, r5 = timed_call(lambda: escapement_sweep("fishing-v1", levels=L5, replicates=R5, T=T5, device=dev, seed=0,
                                                      r=0.3, K=1.0, sigma=0.1, init_state=0.75, Tmax=1000))
        # the collective alone: the [levels, 3] block of this rank through the same deterministic all-gather + tree
        blk = torch.stack([r5["sum"], r5["sumsq"], r5["n_done"]], dim=1).contiguous()
        allreduce_stats(blk)
        barrier()
        c0, c1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        c0.record()
        for _ in range(10):
            allreduce_stats(blk)
        c1.record()
        barrier()
        coll_ms = barrier_max_ms(c0.elapsed_time(c1) / 10, dev)
        best = int(torch.argmax(r5["mean"]).item())
        configs["cfg5"] = {
            "workload": f"escapement sweep: {L5} levels x {R5} replicates x {T5} steps of fishing-v1 (sigma=0.1, common random numbers), "
                        f"replicates sharded over {world} GPU(s); reset + sweep kernel + K4 + all-gather of the [{L5},3] statistics inside the timed call",
            "value": L5 * R5 * T5 / (ms5 * 1e-3), "unit": UNIT, "ms": ms5, "scaling": "strong",
            "collective_ms": coll_ms if world > 1 else 0.0, "collective_share": (coll_ms / ms5) if world > 1 else 0.0,
            "checksum": {"sum_of_sums": float(r5["sum"].sum()).hex(), "sumsq_17": float(r5["sumsq"][17]).hex(),
                         "mean_200": float(r5["mean"][200]).hex(), "n_done": float(r5["n_done"].sum())},
            "best_level": best, "best_S": float(r5["S"][best]), "best_mean_return": float(r5["mean"][best])}
        del r3, r5, blk
        torch.cuda.empty_cache()

    # ---- what the link alone can carry: the copies of the e2e section without any kernel, both directions in flight
    # (tools/pcie_probe.py holds the standalone version, profiles/pcie_probe_r02.txt its records at N = 1 and N = 8)
    def link_rate(h2d_b, d2h_b, n_env=n * 8, reps=6):
        h_in = torch.empty(int(n_env * h2d_b), dtype=torch.uint8).pin_memory()
        d_in = torch.empty_like(h_in, device=dev)
        d_out = torch.empty(int(n_env * d2h_b), dtype=torch.uint8, device=dev)
        h_out = torch.empty(d_out.numel(), dtype=torch.uint8).pin_memory()
        s1, s2 = torch.cuda.Stream(device=dev), torch.cuda.Stream(device=dev)
        best = 0.0
        for _ in range(3):
            barrier()
            t0 = time.perf_counter()
            for _ in range(reps):
                with torch.cuda.stream(s1):
                    d_in.copy_(h_in, non_blocking=True)
                with torch.cuda.stream(s2):
                    h_out.copy_(d_out, non_blocking=True)
            torch.cuda.synchronize()
            dt = barrier_max_ms((time.perf_counter() - t0) * 1e3, dev) * 1e-3
            best = max(best, world * n_env * reps / dt)
        return best

    # ---- e2e: numpy host buffers through the C-ABI host front end (H2D of actions, D2H of obs/reward/done every lock-step)
    e2e_T = min(T, args.e2e_T)
    # host side of the copies: run on (and allocate the pinned buffers from) the cores next to this GPU; undone before the
    # CPU baseline below, which must see every core of the box
    from gym_fishing_b200.distributed import bind_cpu_to_gpu

    affinity0 = os.sched_getaffinity(0)
    bound_cpus = 0 if args.no_cpu_bind else bind_cpu_to_gpu(local)
    venv = FishingVecEnv("fishing-v1", num_envs=n, device=dev, io="numpy", seed=0, auto_reset=True, env_id_offset=rank * n,
                         r=CFG["r"], K=1, sigma=CFG["sigma"], init_state=CFG["x0"], Tmax=CFG["Tmax"])
    venv.reset()
    import numpy as np

    from gym_fishing_b200.vec_env import _PinnedArray

    pinned = _PinnedArray(n * args.e2e_chunk, np.float64)
    h_act = pinned.array.reshape(args.e2e_chunk, n)
    h_act[:] = np.random.default_rng(rank).uniform(-1, 1, size=(args.e2e_chunk, n))
    for t in range(3):
        venv.step(h_act[t % 8])
    barrier()
    t0 = time.perf_counter()
    acc = 0.0
    per_step_T = min(e2e_T, 50)
    for t in range(per_step_T):
        o, r_, d_, _ = venv.step(h_act[t % 8])
        acc += float(r_[0])
    torch.cuda.synchronize()
    per_step_ms = barrier_max_ms((time.perf_counter() - t0) * 1e3, dev)
    e2e_per_step = world * n * per_step_T / (per_step_ms * 1e-3)
    # whole-rollout host call: actions [T, n] in pinned host memory -> obs / reward / done [T, n] in pinned host memory
    chunk_T = args.e2e_chunk
    p_obs, p_rew, p_done = _PinnedArray(chunk_T * n, np.float64), _PinnedArray(chunk_T * n, np.float64), _PinnedArray(chunk_T * n, np.uint8)
    outs = (p_obs.array.reshape(chunk_T, n), p_rew.array.reshape(chunk_T, n), p_done.array.reshape(chunk_T, n))
    venv.step_many(h_act, out=outs)
    barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_T // chunk_T):
        o, r_, d_ = venv.step_many(h_act, out=outs)
        acc += float(r_[-1, 0])
    torch.cuda.synchronize()
    e2e_ms = (time.perf_counter() - t0) * 1e3
    e2e_ms = barrier_max_ms(e2e_ms, dev)
    e2e_T = (e2e_T // chunk_T) * chunk_T
    e2e_value = world * n * e2e_T / (e2e_ms * 1e-3)
    for pa in (p_obs, p_rew, p_done):
        pa.free()
    venv.close()
    # the same through the stable-baselines3 dtypes (float32 actions in; float32 obs + float32 reward + bool done out): 13
    # instead of 25 bytes per env-step over PCIe.  Reported next to the float64 (reference-dtype) figure, which stays `value`.
    venv32 = FishingVecEnv("fishing-v1", num_envs=n, device=dev, io="numpy", seed=0, auto_reset=True, env_id_offset=rank * n,
                           sb3_dtypes=True, r=CFG["r"], K=1, sigma=CFG["sigma"], init_state=CFG["x0"], Tmax=CFG["Tmax"])
    venv32.reset()
    p_act32 = _PinnedArray(chunk_T * n, np.float32)
    a32 = p_act32.array.reshape(chunk_T, n)
    a32[:] = h_act.astype(np.float32)
    q_obs, q_rew, q_done = _PinnedArray(chunk_T * n, np.float32), _PinnedArray(chunk_T * n, np.float32), _PinnedArray(chunk_T * n, np.uint8)
    outs32 = (q_obs.array.reshape(chunk_T, n), q_rew.array.reshape(chunk_T, n), q_done.array.reshape(chunk_T, n))
    venv32.step_many(a32, out=outs32)
    barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_T // chunk_T):
        venv32.step_many(a32, out=outs32)
    torch.cuda.synchronize()
    e2e32_ms = barrier_max_ms((time.perf_counter() - t0) * 1e3, dev)
    e2e32_value = world * n * e2e_T / (e2e32_ms * 1e-3)
    venv32.close()
    # float32 + bit-packed dones: 4 B in, 8.125 B out per env-step
    venvp = FishingVecEnv("fishing-v1", num_envs=n, device=dev, io="numpy", seed=0, auto_reset=True, env_id_offset=rank * n,
                          sb3_dtypes=True, packed_done=True, r=CFG["r"], K=1, sigma=CFG["sigma"], init_state=CFG["x0"], Tmax=CFG["Tmax"])
    venvp.reset()
    q_bits = _PinnedArray(chunk_T * ((n + 7) // 8), np.uint8)
    outsp = (outs32[0], outs32[1], q_bits.array.reshape(chunk_T, (n + 7) // 8))
    venvp.step_many(a32, out=outsp)
    barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_T // chunk_T):
        venvp.step_many(a32, out=outsp)
    torch.cuda.synchronize()
    e2ep_ms = barrier_max_ms((time.perf_counter() - t0) * 1e3, dev)
    e2ep_value = world * n * e2e_T / (e2ep_ms * 1e-3)
    venvp.close()
    for pa in (p_act32, q_obs, q_rew, q_done, q_bits):
        pa.free()
    pinned.free()
    link_f64, link_f32, link_packed = link_rate(8, 17), link_rate(4, 9), link_rate(4, 8.125)
    os.sched_setaffinity(0, affinity0)

    if rank == 0:
        peak, peak_src = measured_peak()
        achieved = K1_BYTES * n / (k1_us * 1e-6) / 1e9
        traffic, traffic_src = None, None      # NOT measured in this run: the ncu figure recorded under profiles/
        try:
            with open(os.path.join(ROOT, "profiles", "k1_traffic.json")) as f:
                tj = json.load(f)
            traffic, traffic_src = tj.get("dram_bytes_per_launch"), "profiles/k1_traffic.json: " + tj.get("source", "")
        except Exception:  # noqa: BLE001
            pass
        cfg1 = cfg1_leg(dev) if (world == 1 and not args.no_cpu_baseline) else None
        cpu = None
        if world == 1 and not args.no_cpu_baseline:
            have_ref = reference_available()
            v1 = cpu_loop_throughput(_ref_worker if have_ref else _port_worker, args.cpu_steps, 1)
            cpu = {"value": v1, "unit": UNIT, "cores": 1, "kind": "reference" if have_ref else "port",
                   "sample": f"{args.cpu_steps} env.step calls of " + ("the unmodified reference (baseline/_ref/gym_fishing under the oracle/refstub gym stand-in)"
                                                                       if have_ref else "the reference-shaped scalar port (oracle.ScalarPortEnv)") + ", fishing-v1 sigma=0.01, one core",
                   "scalar_port_1core": cpu_port_throughput(min(args.cpu_steps, 300_000), 1) if have_ref else None,
                   "numpy_vectorised_1core": cpu_vectorised_throughput(),
                   "c_openmp_all_cores": cpu_c_port_throughput(), "host_cores": os.cpu_count(),
                   "note": "value = the reference's own step() loop on one core; scalar_port / numpy-vectorised / C-OpenMP are the oracle's "
                           "restatements (pre-generated noise for the latter two), for context"}
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic", "config": workload_config(args, world),
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": traffic, "traffic_source": traffic_src,
                         "kernel": "step_kernel_vec2<LOGISTIC,P_KONE,F64>", "bytes_per_env_step": K1_BYTES,
                         "avg_launch_us": k1_us, "peak_source": peak_src,
                         "note": "at 2^20 envs the state / years arrays (12 MB) live in the 126 MB L2, so the algorithmic-byte rate can pass "
                                 "the measured DRAM copy rate; a compute-free replica of the same traffic (tools/k1_traffic_probe.cu, "
                                 "profiles/k1_traffic_probe_r02.txt) needs 5.5 us per launch",
                         "other_sizes": k1_points},
            "cpu_baseline": cpu,
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": n * 8, "d2h_bytes_per_step": n * 17,
                    "step_definition": "one lock-step of the n-env batch per GPU (actions f64 in; obs f64 + reward f64 + done u8 out)",
                    "lock_steps_timed": e2e_T, "host_cores_bound_to_gpu_numa_node": bound_cpus,
                    "api": f"FishingVecEnv(io='numpy').step_many -> fishing_vec_step_many (actions/obs/reward/done in pinned host memory, {args.e2e_chunk} lock-steps per call: one strided H2D and three strided D2H copies per twelfth of the batch)",
                    "per_step_api": {"value": e2e_per_step, "unit": UNIT, "api": "FishingVecEnv(io='numpy').step -> fishing_vec_step, one call per lock-step"},
                    "link_only": {"value": link_f64, "unit": UNIT,
                                  "what": "the same 8 B in / 17 B out per env-step as bare pinned-memory copies, both directions in flight, no kernels"},
                    "frac_of_pcie": e2e_value / link_f64,
                    "sb3_dtypes": {"value": e2e32_value, "unit": UNIT, "h2d_bytes_per_step": n * 4, "d2h_bytes_per_step": n * 9,
                                   "link_only": link_f32, "frac_of_pcie": e2e32_value / link_f32,
                                   "api": "FishingVecEnv(io='numpy', sb3_dtypes=True).step_many -> fishing_vec_step_many_f32 (float32 actions / obs / rewards, the dtypes a stable-baselines3 VecEnv exchanges)"},
                    "sb3_dtypes_packed_done": {"value": e2ep_value, "unit": UNIT, "h2d_bytes_per_step": n * 4, "d2h_bytes_per_step": n * 8 + (n + 7) // 8,
                                               "link_only": link_packed, "frac_of_pcie": e2ep_value / link_packed,
                                               "api": "FishingVecEnv(io='numpy', sb3_dtypes=True, packed_done=True).step_many -> fishing_vec_step_many_wire (dones cross PCIe bit-packed; unpacking on the host is not timed)"}},
            "fused": fused,
            "configs": configs,
            "cfg1": cfg1,
            "gpu_launches": int(launches_per_step * args.steps),
            "clocks": clocks.summary(),
            "check": {"frac_done_last_step": frac_done, "mean_reward_last_step": mean_rew},
        }
        emit(line)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--envs", type=int, default=1 << 20)
    ap.add_argument("--T", type=int, default=1000)
    ap.add_argument("--e2e-T", dest="e2e_T", type=int, default=256)
    ap.add_argument("--e2e-chunk", dest="e2e_chunk", type=int, default=32, help="lock-steps per fishing_vec_step_many call")
    ap.add_argument("--cpu-steps", dest="cpu_steps", type=int, default=1_000_000)
    ap.add_argument("--ref-steps-per-core", dest="ref_steps_per_core", type=int, default=100_000)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-configs", action="store_true", help="skip the cfg3 / cfg5 legs")
    ap.add_argument("--cfg3-envs", dest="cfg3_envs", type=int, default=1 << 24)
    ap.add_argument("--cfg5-replicates", dest="cfg5_replicates", type=int, default=16384)
    ap.add_argument("--ref-port", action="store_true", help="reference arm: time the oracle's scalar port even when baseline/_ref exists")
    ap.add_argument("--no-cpu-bind", action="store_true", help="do not pin the e2e section to the GPU-local cores")
    ap.add_argument("--no-graph", action="store_true", help="issue the K1 launches directly instead of replaying a CUDA graph")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "b200" else args.warmup
    # stdout carries the ONE JSON line and nothing else: libraries that write to file descriptor 1 themselves (NCCL prints
    # "NCCL version ..." there when NCCL_DEBUG=VERSION is set on the box) are pointed at stderr for the whole run, and the
    # line goes out through a private duplicate of the original stdout.
    sys.stdout.flush()
    json_fd = os.dup(1)
    os.dup2(2, 1)
    global _JSON_OUT
    _JSON_OUT = os.fdopen(json_fd, "w")
    if args.impl == "reference":
        run_reference(args)
    else:
        run_gpu(args)


if __name__ == "__main__":
    main()
