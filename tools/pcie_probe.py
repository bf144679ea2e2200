#!/usr/bin/env python
"""Raw host<->device copy rates of the box next to what the e2e host front end moves (pinned memory, copy engines only, no
kernels).  Run alone (`python tools/pcie_probe.py`) or with all ranks copying at once
(`python -m torch.distributed.run --nproc-per-node 8 tools/pcie_probe.py`): every figure is then the per-rank rate while
all N GPUs stream concurrently, and the aggregate is printed by rank 0.  Results are kept under profiles/pcie_probe_*.txt.

The last block is the one the bench line's `e2e.frac_of_pcie` refers to: an H2D copy of 8 B per environment overlapped with a
D2H copy of 17 B per environment (the float64 wire format of FishingVecEnv(io="numpy")), and 4 B : 9 B / 4 B : 8.125 B for
the float32 / float32 + packed-dones formats -- the rate at which the link alone could carry env-steps."""
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from gym_fishing_b200.distributed import init_from_env  # noqa: E402


def mixed_rate(n_env, h2d_bytes_per_env, d2h_bytes_per_env, reps=10, barrier=None):
    """env-steps/s when ONLY the copies of `reps` lock-steps are done: both directions in flight at once."""
    h_in = torch.empty(int(n_env * h2d_bytes_per_env), dtype=torch.uint8).pin_memory()
    d_in = torch.empty_like(h_in, device="cuda")
    d_out = torch.empty(int(n_env * d2h_bytes_per_env), dtype=torch.uint8, device="cuda")
    h_out = torch.empty(d_out.numel(), dtype=torch.uint8).pin_memory()
    s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
    best = 0.0
    for _ in range(3):
        if barrier:
            barrier()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(reps):
            with torch.cuda.stream(s1):
                d_in.copy_(h_in, non_blocking=True)
            with torch.cuda.stream(s2):
                h_out.copy_(d_out, non_blocking=True)
        torch.cuda.synchronize()
        best = max(best, n_env * reps / (time.perf_counter() - t0))
    return best


def one_way(nbytes, direction, reps=10, barrier=None):
    h = torch.empty(nbytes, dtype=torch.uint8).pin_memory()
    d = torch.empty(nbytes, dtype=torch.uint8, device="cuda")
    best = 0.0
    for _ in range(3):
        if barrier:
            barrier()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(reps):
            if direction == "h2d":
                d.copy_(h, non_blocking=True)
            else:
                h.copy_(d, non_blocking=True)
        torch.cuda.synchronize()
        best = max(best, nbytes * reps / (time.perf_counter() - t0) / 1e9)
    return best


def main():
    rank, world, local = init_from_env()
    torch.cuda.set_device(local)
    import torch.distributed as dist

    barrier = dist.barrier if world > 1 else None
    n = 1 << 20
    res = {"world": world, "rank": rank}
    for mb in (8, 64):
        res[f"h2d_{mb}MiB_GBs"] = one_way(mb << 20, "h2d", barrier=barrier)
        res[f"d2h_{mb}MiB_GBs"] = one_way(mb << 20, "d2h", barrier=barrier)
    res["f64_wire_8B_in_17B_out_env_steps_per_s"] = mixed_rate(n * 8, 8, 17, barrier=barrier)
    res["f32_wire_4B_in_9B_out_env_steps_per_s"] = mixed_rate(n * 8, 4, 9, barrier=barrier)
    res["f32_packed_wire_4B_in_8.125B_out_env_steps_per_s"] = mixed_rate(n * 8, 4, 8.125, barrier=barrier)
    if world > 1:
        keys = [k for k in res if k not in ("world", "rank")]
        t = torch.tensor([res[k] for k in keys], dtype=torch.float64, device="cuda")
        tmin = t.clone()
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        dist.all_reduce(tmin, op=dist.ReduceOp.MIN)
        if rank == 0:
            print(json.dumps({"world": world, "aggregate": dict(zip(keys, t.tolist())), "slowest_rank": dict(zip(keys, tmin.tolist()))}))
        dist.barrier()
        dist.destroy_process_group()
    else:
        print(json.dumps(res))


if __name__ == "__main__":
    main()
